"""Dev / evidence: BASELINE.json configs at FULL size through the two command lines.

For every named config the synthetic DBs are written to a scratch directory, the unmodified
reference (oracle/_ref/damapper -T<n>) and the product's host driver (damapper_b200/damapper)
map the same blocks with the same flags, and the per-thread .las record streams (M files, R
files with -C) and the -p track are compared byte for byte, block by block.  Both runs are
whole-process wall clocks (exec -> exit: DB load, .las writes, LAsort/LAcat stubbed), SURVEY.md
section 8(d)(i).  One JSON line per config goes to stdout and to gpurun_out/cli_full.jsonl.

    python tools/gpu_cli_full.py C2 C5 C3 C4          # env SCALE=<f> shrinks every config
"""
import glob
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from damapper_b200 import dazzdb, las, synth  # noqa: E402
from oracle import run_ref  # noqa: E402

EXE = os.environ.get("DAMCLI_EXE", os.path.join(ROOT, "damapper_b200", "damapper"))
FLAGS = {"C1": [], "C2": [], "C3": ["-n.95", "-p"], "C4": [], "C5": ["-C", "-p"]}
BLOCKS = {"C1": 1, "C2": 1, "C3": 2, "C4": 8, "C5": 1}
CHUNK = 12500                                            # reads generated at a time


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def build_dbs(cfg, scale, wd, seed=7):
    c = synth.CONFIGS[cfg]
    G = max(20000, int(c["genome"] * scale))
    R = max(8, int(c["nreads"] * scale))
    genome = synth.make_repeat_genome(G, seed=seed) if c["kind"] == "repeat" else synth.make_genome(G, seed=seed)
    nc = c["contigs"]
    cuts = [0] + [int(G * (i + 1) / nc) for i in range(nc)]
    if nc == 2:
        cuts = [0, int(G * 0.6), G]
    w = dazzdb.StreamDBWriter(os.path.join(wd, "ref.dam"), is_dam=True)
    w.append(genome, np.diff(cuts))
    w.close()
    w = dazzdb.StreamDBWriter(os.path.join(wd, "reads.db"))
    done = 0
    bases = 0
    while done < R:
        n = min(CHUNK, R - done)
        if c["kind"] == "chimeric":
            b, rl, _ = synth.make_chimeric_reads(genome, n, seed=seed + 1 + done)
        else:
            b, rl, _ = synth.make_reads(genome, n, seed=seed + 1 + done, contig_bounds=np.array(cuts))
        w.append(b, rl)
        bases += int(rl.sum())
        done += n
    nb = BLOCKS[cfg] if R >= 8 * BLOCKS[cfg] else 1
    w.close(nblocks=nb)
    return G, R, bases, nb


def run_cli(exe, wd, tag, flags, threads, blocks, env_extra=None, timeout=3000):
    keep = os.path.join(wd, "keep_" + tag)
    shutil.rmtree(keep, ignore_errors=True)
    os.makedirs(keep)
    tmp = os.path.join(wd, "tmp_" + tag)
    os.makedirs(tmp, exist_ok=True)
    for f in glob.glob(os.path.join(wd, ".reads*.prof.*")):
        os.remove(f)
    env = dict(os.environ)
    env["PATH"] = os.path.join(run_ref.REF_DIR, "bin") + os.pathsep + env.get("PATH", "")
    env["DAMAPPER_KEEP_DIR"] = keep
    env.update(env_extra or {})
    args = ["reads.%d" % (i + 1) for i in range(blocks)] if blocks > 1 else ["reads"]
    cmd = [exe, "-T%d" % threads, "-P" + tmp] + flags + ["ref.dam"] + args
    t0 = time.perf_counter()
    p = subprocess.run(cmd, cwd=wd, env=env, capture_output=True, text=True, timeout=timeout)
    wall = time.perf_counter() - t0
    if p.returncode != 0:
        raise RuntimeError("%s failed (%d)\n%s\n%s" % (" ".join(cmd), p.returncode, p.stdout[-2000:], p.stderr[-2000:]))
    profs = {}
    for f in glob.glob(os.path.join(wd, ".reads*.prof.data")) + glob.glob(os.path.join(wd, ".reads*.prof.anno")):
        dst = os.path.join(keep, os.path.basename(f))
        shutil.move(f, dst)
        profs[os.path.basename(f)] = dst
    for line in p.stderr.splitlines():
        if line.startswith("[timing]"):
            log("   ", tag, line)
    return wall, keep, profs, p.stdout


def streams(keep, blocks):
    out = {}
    for b in range(blocks):
        root = "reads.%d" % (b + 1) if blocks > 1 else "reads"
        m = run_ref._thread_sorted(glob.glob(os.path.join(keep, root + ".ref.M[0-9]*.las")))
        r = run_ref._thread_sorted(glob.glob(os.path.join(keep, "ref." + root + ".R[0-9]*.las")))
        out[root] = (las.canonical_stream(m) if m else b"", las.canonical_stream(r) if r else b"")
    return out


def one(cfg, scale, threads):
    wd = tempfile.mkdtemp(prefix="damcli_" + cfg + "_", dir=os.environ.get("SCRATCH", "/tmp"))
    res = {"config": cfg, "scale": scale, "threads": threads}
    try:
        t0 = time.time()
        G, R, bases, nb = build_dbs(cfg, scale, wd)
        res.update(ref_bp=G, reads=R, read_bases=bases, read_blocks=nb, gen_s=round(time.time() - t0, 1))
        log(cfg, "generated", res)
        flags = ["-M%d" % int(os.environ.get("MEMGB", "32"))] + FLAGS[cfg]
        res["flags"] = flags
        gw, gkeep, gprof, gout = run_cli(EXE, wd, "gpu", flags, threads, nb)
        res["gpu_wall_s"] = round(gw, 3)
        log(cfg, "product CLI %.2fs" % gw)
        gw2, gkeep, gprof, gout = run_cli(EXE, wd, "gpu", flags, threads, nb,        # page cache warm for both arms
                                          env_extra={"DAMGPU_TIMING": "1"} if os.environ.get("CLI_TIMING") else None)
        res["gpu_wall_s_2nd"] = round(gw2, 3)
        log(cfg, "product CLI (2nd) %.2fs" % gw2)
        gs = streams(gkeep, nb)
        gp = {k: open(v, "rb").read() for k, v in gprof.items()}
        rw, rkeep, rprof, rout = run_cli(run_ref.REF_BIN, wd, "ref", flags, threads, nb)
        res["ref_wall_s"] = round(rw, 3)
        log(cfg, "reference CLI %.2fs" % rw)
        rs = streams(rkeep, nb)
        rp = {k: open(v, "rb").read() for k, v in rprof.items()}
        ok = True
        nrec = 0
        for root in rs:
            a_ok = rs[root][0] == gs[root][0]
            b_ok = rs[root][1] == gs[root][1]
            ok &= a_ok and b_ok
            nrec += sum(1 for _ in las.stream_records(rs[root][0], 100)) if len(rs[root][0]) < (1 << 27) else -1
            if not (a_ok and b_ok):
                log(cfg, root, "MISMATCH M", a_ok, len(rs[root][0]), len(gs[root][0]), "R", b_ok, len(rs[root][1]), len(gs[root][1]))
        p_ok = (rp == gp)
        if not p_ok:
            # The reference's -p counters are not deterministic at high -T on repeat-rich input (two runs of
            # the unmodified reference differ from each other; -T1/-T2/-T4 agree): settle it with a -T1 run.
            w1, k1, p1, _ = run_cli(run_ref.REF_BIN, wd, "ref1", flags, 1, nb)
            r1 = {k: open(v, "rb").read() for k, v in p1.items()}
            res.update(ref_prof_differs_between_T=bool(r1 != rp), ref_T1_wall_s=round(w1, 3))
            p_ok = (r1 == gp)
        res.update(parity_las=bool(ok), parity_prof=bool(p_ok), prof_files=len(rp), m_bytes=sum(len(v[0]) for v in rs.values()),
                   r_bytes=sum(len(v[1]) for v in rs.values()), m_records=nrec,
                   gpu_bases_per_s=round(bases / min(gw, gw2)), ref_bases_per_s=round(bases / rw))
    except Exception as e:  # keep going with the next config
        res["error"] = str(e)[-1500:]
    finally:
        shutil.rmtree(wd, ignore_errors=True)
    return res


if __name__ == "__main__":
    cfgs = [a for a in sys.argv[1:] if a in synth.CONFIGS] or ["C2"]
    scale = float(os.environ.get("SCALE", "1"))
    ncpu = os.cpu_count() or 1
    threads = 1
    while threads * 2 <= min(ncpu, 64):
        threads *= 2
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    allok = True
    for cfg in cfgs:
        r = one(cfg, scale, threads)
        line = json.dumps(r)
        print(line, flush=True)
        with open(os.path.join(ROOT, "gpurun_out", "cli_full.jsonl"), "a") as f:
            f.write(line + "\n")
        allok &= bool(r.get("parity_las")) and bool(r.get("parity_prof"))
    print("ALL OK" if allok else "MISMATCH/ERROR")
