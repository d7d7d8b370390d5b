"""Dev / evidence: the host driver with -G<n> on n real devices against -G1 (same per-thread .las streams),
wall clocks of both; run under `gpurun --gpus 2` (or more).  N=<n> picks the device count, BLOCKS the number
of reads blocks, SCALE the C2 scale of every block."""
import glob, json, os, subprocess, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from damapper_b200 import dazzdb, las, synth
from oracle import run_ref
n = int(os.environ.get("N", "2")); blocks = int(os.environ.get("BLOCKS", "4")); scale = float(os.environ.get("SCALE", "0.25"))
exe = os.path.join(ROOT, "damapper_b200", "damapper")
contigs, rb, rl = synth.make_config("C2", scale=1.0, seed=7)
genome = np.concatenate(contigs); cuts = np.array([0, contigs[0].size, genome.size])
wd = tempfile.mkdtemp(prefix="gmulti_")
dazzdb.write_db(os.path.join(wd, "ref.dam"), contigs, is_dam=True)
w = dazzdb.StreamDBWriter(os.path.join(wd, "reads.db")); bounds = [0]; bases = 0
for b in range(blocks):
    bb, bl, _ = synth.make_reads(genome, int(13800 * scale), seed=100 + b, contig_bounds=cuts)
    w.append(bb, bl); bounds.append(bounds[-1] + len(bl)); bases += int(bl.sum())
w.close(block_bounds=bounds)
def run(tag, flags, env_extra):
    keep = os.path.join(wd, "keep_" + tag); os.makedirs(keep); tmp = os.path.join(wd, "tmp_" + tag); os.makedirs(tmp)
    env = dict(os.environ, DAMAPPER_KEEP_DIR=keep, **env_extra)
    env["PATH"] = os.path.join(run_ref.REF_DIR, "bin") + os.pathsep + env["PATH"]
    t0 = time.time()
    p = subprocess.run([exe, "-T4", "-P" + tmp, "-M32"] + flags + ["ref.dam"] + ["reads.%d" % (b + 1) for b in range(blocks)],
                       cwd=wd, env=env, capture_output=True, text=True)
    dt = time.time() - t0
    assert p.returncode == 0, p.stderr
    out = {}
    for b in range(1, blocks + 1):
        out[b] = las.canonical_stream(run_ref._thread_sorted(glob.glob(os.path.join(keep, "reads.%d.ref.M[0-9]*.las" % b))))
    return out, dt
one, t1 = run("g1", [], {})
one, t1b = run("g1b", [], {})
many, tn = run("gn", ["-G%d" % n], {})
many, tnb = run("gnb", ["-G%d" % n], {})
print(json.dumps({"devices": n, "reads_blocks": blocks, "read_bases": bases, "same_las_streams": one == many,
                  "records_bytes": sum(len(v) for v in one.values()),
                  "wall_s_G1": [round(t1, 3), round(t1b, 3)], "wall_s_Gn": [round(tn, 3), round(tnb, 3)]}))
