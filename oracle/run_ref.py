"""TEST INFRASTRUCTURE ONLY: run the compiled, unmodified reference (oracle/_ref/damapper).

Used by the golden-vector generator, by CPU tests that validate the oracle restatement and by
bench.py's cpu_baseline / --impl reference legs.  Nothing in the product path imports this.

The reference deletes its per-thread outputs unless LAsort/LAcat/LAmerge succeed
(damapper.c:543-554,882-911); the stubs in oracle/_ref/bin copy the per-thread files into
$DAMAPPER_KEEP_DIR first.
"""
from __future__ import annotations

import glob
import os
import re
import subprocess
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
REF_BIN = os.path.join(REF_DIR, "damapper")


def have_ref() -> bool:
    return os.access(REF_BIN, os.X_OK) and os.access(os.path.join(REF_DIR, "bin", "LAsort"), os.X_OK)


def _thread_sorted(files):
    def key(p):
        m = re.search(r"\.[MR](\d+)\.las$", p)
        return int(m.group(1))
    return sorted(files, key=key)


def run_damapper(workdir: str, ref: str, reads: str, flags=(), threads: int = 4, timeout=3600, exe=None, env_extra=None):
    """Run `damapper <flags> -T<threads> ref reads` in `workdir` (exe: another binary that keeps the
    reference command line: oracle/_ref/damapper_timed, oracle/_ref/damapper_gpu, damapper_b200/damapper).

    Returns dict(wall_s, core_s, m_files, r_files, prof_anno, prof_data, stdout)."""
    keep = os.path.join(workdir, "keep")
    os.makedirs(keep, exist_ok=True)
    for f in glob.glob(os.path.join(keep, "*.las")):
        os.remove(f)
    env = dict(os.environ)
    env["PATH"] = os.path.join(REF_DIR, "bin") + os.pathsep + env.get("PATH", "")
    env["DAMAPPER_KEEP_DIR"] = keep
    sortdir = os.path.join(workdir, "tmp")
    os.makedirs(sortdir, exist_ok=True)
    if env_extra:
        env.update(env_extra)
    cmd = [exe or REF_BIN, "-T%d" % threads, "-P" + sortdir] + list(flags) + [ref, reads]
    t0 = time.perf_counter()
    p = subprocess.run(cmd, cwd=workdir, env=env, capture_output=True, text=True, timeout=timeout)
    wall = time.perf_counter() - t0
    if p.returncode != 0:
        raise RuntimeError("reference damapper failed: %s\n%s\n%s" % (cmd, p.stdout, p.stderr))
    m_files = _thread_sorted(glob.glob(os.path.join(keep, "*.M[0-9]*.las")))
    r_files = _thread_sorted(glob.glob(os.path.join(keep, "*.R[0-9]*.las")))
    rroot = os.path.basename(reads)
    for ext in (".db", ".dam"):
        if rroot.endswith(ext):
            rroot = rroot[: -len(ext)]
    anno = os.path.join(workdir, "." + rroot + ".prof.anno")
    data = os.path.join(workdir, "." + rroot + ".prof.data")
    core = re.search(r"\[core\] ([0-9.]+) s", p.stderr or "")
    return dict(wall_s=wall, core_s=float(core.group(1)) if core else None, m_files=m_files, r_files=r_files,
                prof_anno=anno if os.path.exists(anno) else None,
                prof_data=data if os.path.exists(data) else None,
                stdout=p.stdout, stderr=p.stderr)
