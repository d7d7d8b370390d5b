"""Dev: whole-process phases of the host driver on the C2 workload (DAMGPU_TIMING=1).
RUNS=<n> runs per variant; VARIANTS="name:ENV=val,ENV=val;..." extra environment variants."""
import os, sys, subprocess, tempfile, time, shutil
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from damapper_b200 import synth, dazzdb
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
exe = os.path.join(ROOT, "damapper_b200", "damapper")
contigs, rb, rl = synth.make_config("C2", scale=float(os.environ.get("SCALE", "1.0")), seed=7)
wd = tempfile.mkdtemp(prefix="clit_")
dazzdb.write_db(os.path.join(wd, "ref.dam"), contigs, is_dam=True)
dazzdb.write_db(os.path.join(wd, "reads.db"), (rb, rl))
os.makedirs(os.path.join(wd, "tmp"))
variants = [("default", {})]
for v in filter(None, os.environ.get("VARIANTS", "").split(";")):
    name, kv = v.split(":")
    variants.append((name, dict(x.split("=") for x in kv.split(","))))
runs = int(os.environ.get("RUNS", "3"))
for name, extra in variants:
    env = dict(os.environ); env["DAMGPU_TIMING"] = "1"; env["DAMGPU_BUILTIN_SORT"] = "1"; env.update(extra)
    walls = []
    for it in range(runs):
        t0 = time.time()
        p = subprocess.run([exe, "-T16", "-M32", "-P" + os.path.join(wd, "tmp"), "ref.dam", "reads.db"], cwd=wd, env=env,
                           capture_output=True, text=True)
        walls.append(time.time() - t0)
    print("%s: rc %d walls %s" % (name, p.returncode, " ".join("%.3f" % w for w in walls)))
    print(p.stderr[-3500:])
shutil.rmtree(wd, ignore_errors=True)
