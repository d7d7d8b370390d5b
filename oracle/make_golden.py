"""TEST INFRASTRUCTURE ONLY: generate tests/golden/*.npz by running the compiled, UNMODIFIED
reference (oracle/_ref/damapper, built by oracle/Makefile from /root/reference).

Run here (where /root/reference exists):  python oracle/make_golden.py
Each fixture stores the generator parameters (inputs are re-created deterministically from
them by damapper_b200.synth), a sha256 of the generated inputs, and the reference's outputs:
canonical M and R record streams (padding bytes 36-39 zeroed) and the -p track.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from damapper_b200 import synth, dazzdb, las          # noqa: E402
from oracle import run_ref                            # noqa: E402

# name -> (config, scale, seed, reference flags, oracle/product keyword arguments)
CASES = {
    "c1_default": ("C1", 0.02, 3, (), {}),
    "c5_cover_profile": ("C5", 0.05, 13, ("-C", "-p"), dict(do_b=1, profile=1)),
    "c3_repeat_n95": ("C3", 0.002, 15, ("-p", "-n.95"), dict(profile=1, best_tie=0.95)),
    "c1_k16_s50_t20_e80": ("C1", 0.03, 18, ("-k16", "-s50", "-t20", "-e.8", "-C"),
                           dict(kmer=16, spacing=50, suppress=20, ave_corr=0.8, do_b=1)),
    "c1_k24_s200": ("C1", 0.03, 19, ("-k24", "-s200", "-C"), dict(kmer=24, spacing=200, do_b=1)),
}


# -m: name -> (config, scale, seed, flags, kwargs); the tracks are dazzdb.random_masks(seed..):
# "dust" on the reference and on the reads, "tan" on the reads only (so both get merged there)
MASK_CASES = {
    "c1_masks": ("C1", 0.03, 71, ("-C", "-mdust", "-mtan"), dict(do_b=1)),
}


def mask_tracks(contigs, rl, seed):
    """(reference dust, reads dust, reads tan) for a mask case."""
    glen = np.array([c.size for c in contigs], dtype=np.int64)
    return (dazzdb.random_masks(glen, seed=seed + 1, max_intervals=40, max_len=2000),
            dazzdb.random_masks(rl, seed=seed + 2, max_intervals=3, max_len=600),
            dazzdb.random_masks(rl, seed=seed + 3, max_intervals=2, max_len=300))


def write_mask_case(wd, contigs, rb, rl, seed):
    ref = dazzdb.write_db(os.path.join(wd, "ref.dam"), contigs, is_dam=True)
    rds = dazzdb.write_db(os.path.join(wd, "reads.db"), (rb, rl))
    gd, rdust, rtan = mask_tracks(contigs, rl, seed)
    dazzdb.write_mask_track(ref, "dust", *gd)
    dazzdb.write_mask_track(rds, "dust", *rdust)
    dazzdb.write_mask_track(rds, "tan", *rtan)
    return gd, dazzdb.union_masks(rdust, rtan)


def input_digest(contigs, rb, rl) -> str:
    h = hashlib.sha256()
    for c in contigs:
        h.update(np.ascontiguousarray(c).tobytes())
    h.update(np.ascontiguousarray(rb).tobytes())
    h.update(np.ascontiguousarray(rl, dtype=np.int64).tobytes())
    return h.hexdigest()


def make_inputs(case):
    cfg, scale, seed, _, _ = CASES[case]
    return synth.make_config(cfg, scale=scale, seed=seed)


def main():
    assert run_ref.have_ref(), "build oracle/_ref first: make -C oracle ref"
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, (cfg, scale, seed, flags, kw) in CASES.items():
        contigs, rb, rl = synth.make_config(cfg, scale=scale, seed=seed)
        wd = tempfile.mkdtemp(prefix="golden_")
        try:
            dazzdb.write_db(os.path.join(wd, "ref.dam"), contigs, is_dam=True)
            dazzdb.write_db(os.path.join(wd, "reads.db"), (rb, rl))
            r = run_ref.run_damapper(wd, "ref.dam", "reads.db", flags=flags, threads=4)
            a = las.canonical_stream(r["m_files"])
            b = las.canonical_stream(r["r_files"]) if r["r_files"] else b""
            prof = open(r["prof_data"], "rb").read() if r["prof_data"] else b""
        finally:
            shutil.rmtree(wd, ignore_errors=True)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"),
                            a=np.frombuffer(a, dtype=np.uint8), b=np.frombuffer(b, dtype=np.uint8),
                            prof=np.frombuffer(prof, dtype=np.uint8),
                            digest=np.array(input_digest(contigs, rb, rl)),
                            flags=np.array(" ".join(flags)))
        print(name, "records bytes", len(a), len(b), "prof", len(prof))
    for name, (cfg, scale, seed, flags, kw) in MASK_CASES.items():
        contigs, rb, rl = synth.make_config(cfg, scale=scale, seed=seed)
        wd = tempfile.mkdtemp(prefix="golden_")
        try:
            write_mask_case(wd, contigs, rb, rl, seed)
            r = run_ref.run_damapper(wd, "ref.dam", "reads.db", flags=flags, threads=4)
            a = las.canonical_stream(r["m_files"])
            b = las.canonical_stream(r["r_files"]) if r["r_files"] else b""
        finally:
            shutil.rmtree(wd, ignore_errors=True)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"),
                            a=np.frombuffer(a, dtype=np.uint8), b=np.frombuffer(b, dtype=np.uint8),
                            prof=np.zeros(0, dtype=np.uint8),
                            digest=np.array(input_digest(contigs, rb, rl)),
                            flags=np.array(" ".join(flags)))
        print(name, "records bytes", len(a), len(b))


if __name__ == "__main__":
    main()
