"""Differential fuzz of the oracle restatement against the unmodified reference (oracle/_ref/damapper),
CPU only: random flag sets, seeds and workload shapes at sizes that run in seconds.  Prints one line per
case and a summary; a mismatch is a finding about the ORACLE (test infrastructure), to be fixed before
any GPU parity claim that rests on it.

    python tools/oracle_fuzz.py [ncases] [seed] [masks]
"""
import os
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from conftest import base_freq, make_case
    from damapper_b200 import dazzdb, las
    from oracle import oracle as orc, run_ref
    ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    masked = len(sys.argv) > 3 and sys.argv[3] == "masks"
    bad = 0
    for it in range(ncases):
        cfg = rng.choice(["C1", "C1", "C5", "C3"])
        scale = {"C1": 0.02, "C5": 0.04, "C3": 0.002}[cfg] * float(rng.uniform(0.5, 1.5))
        seed = int(rng.integers(100, 100000))
        flags, kw = [], {}
        k = int(rng.choice([14, 16, 18, 20, 20, 24, 28, 32]))
        if k != 20:
            flags.append("-k%d" % k); kw["kmer"] = k
        if rng.random() < 0.5:
            flags.append("-C"); kw["do_b"] = 1
        if rng.random() < 0.5:
            flags.append("-p"); kw["profile"] = 1
        if rng.random() < 0.5:
            n = float(rng.choice([0.7, 0.8, 0.9, 0.95]))
            flags.append("-n%g" % n); kw["best_tie"] = n
        if rng.random() < 0.4:
            e = float(rng.choice([0.7, 0.75, 0.8, 0.9]))
            flags.append("-e%g" % e); kw["ave_corr"] = e
        if rng.random() < 0.4:
            s = int(rng.choice([50, 80, 120, 126, 200]))
            flags.append("-s%d" % s); kw["spacing"] = s
        if rng.random() < 0.3:
            t = int(rng.choice([3, 5, 10, 20]))
            flags.append("-t%d" % t); kw["suppress"] = t
        if rng.random() < 0.3:
            m = int(rng.choice([0, 1, 4]))
            flags.append("-M%d" % m); kw["mem_limit"] = m << 30
        else:
            flags.append("-M16"); kw["mem_limit"] = 16 << 30
        contigs, rb, rl, rd, rf, rc = make_case(cfg, scale, seed)
        masks = masked and rng.random() < 0.6               # -mdust -mtan: random interval tracks on both DBs
        rm = gd = gc = None
        wd = tempfile.mkdtemp(prefix="orc_fuzz_")
        try:
            if masks:
                from oracle import make_golden as mg
                gd, rm = mg.write_mask_case(wd, contigs, rb, rl, seed)
                gc = dazzdb.mirror_masks(gd[0], gd[1], rf[2])
                flags += ["-mdust", "-mtan"]
            else:
                dazzdb.write_db(os.path.join(wd, "ref.dam"), contigs, is_dam=True)
                dazzdb.write_db(os.path.join(wd, "reads.db"), (rb, rl))
            # -p on repeat-rich input: the reference's threaded run is not deterministic
            r = run_ref.run_damapper(wd, "ref.dam", "reads.db", flags=flags, threads=1 if "-p" in flags else 2)
            ref_a = las.canonical_stream(r["m_files"])
            ref_b = las.canonical_stream(r["r_files"]) if r["r_files"] else b""
            ref_p = open(r["prof_data"], "rb").read() if r["prof_data"] else b""
        finally:
            shutil.rmtree(wd, ignore_errors=True)
        if masks:
            out = orc.map_block(orc.HostBlock(*rd, mask=rm), [(orc.HostBlock(*rf, mask=gd), orc.HostBlock(*rc, mask=gc))],
                                orc.HostBlock(*rf), freq=base_freq(contigs), **kw)
        else:
            out = orc.map_block(orc.HostBlock(*rd), [(orc.HostBlock(*rf), orc.HostBlock(*rc))],
                                orc.HostBlock(*rf), freq=base_freq(contigs), **kw)
        ok = (out["a"] == ref_a, out["b"] == ref_b, out["prof"] == ref_p)
        if not all(ok):
            bad += 1
        print("%s %s scale=%.4f seed=%d %s -> M %s R %s prof %s (%d + %d bytes)" % (
            "ok  " if all(ok) else "DIFF", cfg, scale, seed, " ".join(flags), ok[0], ok[1], ok[2],
            len(ref_a), len(ref_b)), flush=True)
    print("%d cases, %d mismatches" % (ncases, bad))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
