"""Differential fuzz of the CUDA path against the oracle on a B200: random workload shapes, flag sets,
reference block counts and reads-index modes for a fixed number of seconds.  Every case is announced
before it runs (a fatal error inside the library ends the process: the last announced case is the culprit)
and its verdict appended to gpurun_out/gpu_fuzz.jsonl.

    python tools/gpu_fuzz.py --seconds 60 --seed 1
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def draw(rng):
    cfg = str(rng.choice(["C1", "C1", "C5", "C3", "C2"]))
    scale = {"C1": 0.12, "C5": 0.12, "C3": 0.005, "C2": 0.02}[cfg] * float(rng.uniform(0.3, 2.5))
    kw = {}
    k = int(rng.choice([12, 14, 16, 20, 20, 20, 24, 32]))
    if k != 20: kw["kmer"] = k
    if rng.random() < 0.5: kw["do_b"] = 1
    if rng.random() < 0.5: kw["profile"] = 1
    if rng.random() < 0.5: kw["best_tie"] = float(rng.choice([0.7, 0.8, 0.9, 0.95]))
    if rng.random() < 0.4: kw["ave_corr"] = float(rng.choice([0.7, 0.75, 0.8, 0.9]))
    if rng.random() < 0.4: kw["spacing"] = int(rng.choice([50, 75, 120, 126, 200]))
    if rng.random() < 0.3: kw["suppress"] = int(rng.choice([3, 5, 10, 20]))
    if rng.random() < 0.3: kw["mem_limit"] = int(rng.choice([0, 1, 4])) << 30
    nblocks = int(rng.choice([1, 1, 2, 3]))
    mode = str(rng.choice(["auto", "always", "off"]))
    return cfg, scale, int(rng.integers(100, 1000000)), kw, nblocks, mode


def run_case(api, orc, cfg, scale, seed, kw, nblocks, mode):
    from conftest import base_freq, make_case
    from damapper_b200 import dazzdb
    contigs, rb, rl, rd, rf, rc = make_case(cfg, scale, seed)
    freq = base_freq(contigs)
    g = np.concatenate(contigs)
    if nblocks == 1:
        parts = [contigs]
    else:                                              # the reference DB cut into blocks of whole contigs
        cut = [g.size * i // nblocks for i in range(nblocks + 1)]
        parts = [[g[cut[i]:cut[i + 1]]] for i in range(nblocks)]
    fwd, pairs, first = [], [], 0
    for p in parts:
        f, c = dazzdb.load_block(p), dazzdb.load_block(dazzdb.revcomp_contigs(p))
        fwd.append(api.HostBlock(*f, tfirst=first))
        pairs.append((orc.HostBlock(*f, tfirst=first), orc.HostBlock(*c, tfirst=first)))
        first += len(p)
    whole = dazzdb.load_block([c for p in parts for c in p])
    o = orc.map_block(orc.HostBlock(*rd), pairs, orc.HostBlock(*whole), freq=freq, **kw)
    out = api.map_block(api.HostBlock(*rd), fwd, api.HostBlock(*whole), freq=freq, want_candidates=True,
                        reads_filter=mode, **kw)
    oc, ojc, oj = o["candidates"]
    gc, gjc, gj = out["candidates"]
    res = {"candidates": oc.tobytes() == gc.tobytes(),
           "jumps": bool((ojc == gjc).all()) and oj.tobytes() == gj.tobytes(),
           "M": out["a"] == o["a"], "R": out["b"] == o["b"], "prof": out["prof"] == o["prof"],
           "stats": all(out["stats"][k] == o["stats"][k] for k in ("nalign", "nwaves", "ncells")),
           "trace_check": out["stats"]["trace_fails"] == 0}
    stat_pair = {k: (int(out["stats"][k]), int(o["stats"][k])) for k in ("nalign", "nwaves", "ncells")}
    info = {"stats_gpu_oracle": stat_pair, "reads": int(len(rl)), "records": int(out["anrec"] + out["bnrec"]), "bytes": len(out["a"]) + len(out["b"]),
            "overflow_jobs": int(out["stats"].get("overflow_jobs", 0)), "deferred": bool(out["deferred"]),
            "limit": int(out["limit"])}
    return res, info


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=60.)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--replay", default=None, help="a gpu_fuzz.jsonl: re-run its failed cases first")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "gpu_fuzz.jsonl"))
    args = ap.parse_args()
    from damapper_b200 import api
    from oracle import oracle as orc
    orc.build()
    api.init(0)
    rng = np.random.default_rng(args.seed)
    t_end = time.time() + args.seconds
    n = bad = 0
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    replay = []
    if args.replay and os.path.exists(args.replay):
        for l in open(args.replay):
            d = json.loads(l)
            if not d.get("ok", True):
                replay.append((d["cfg"], d["scale"], d["seed"], d["kw"], d["ref_blocks"], d["reads_index"]))
    with open(args.out, "a") as log:
        while replay or time.time() < t_end:
            cfg, scale, seed, kw, nblocks, mode = replay.pop(0) if replay else draw(rng)
            case = {"cfg": cfg, "scale": round(scale, 5), "seed": seed, "kw": kw, "ref_blocks": nblocks, "reads_index": mode}
            print("case", json.dumps(case), flush=True)
            t0 = time.time()
            res, info = run_case(api, orc, cfg, scale, seed, kw, nblocks, mode)
            ok = all(res.values())
            n += 1
            bad += 0 if ok else 1
            line = dict(case, ok=ok, seconds=round(time.time() - t0, 2), **info)
            if not ok:
                line["differs"] = [k for k, v in res.items() if not v]
            log.write(json.dumps(line) + "\n"); log.flush()
            print("  ->", "ok" if ok else "DIFF %s" % line["differs"], info, flush=True)
    print("gpu_fuzz: %d cases, %d mismatches" % (n, bad))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
