"""The host driver's own logic on CPU: damapper_b200/host/damapper.c + dazz_db.c + las_post.c compiled against
oracle/mock_damgpu.c (a test double of libdamgpu.so built on the oracle -- test infrastructure, see its header)
and run next to the unmodified reference binary on the same databases.  What is under test is everything the
driver does around the library: the reads-block loop and the resident reference blocks (damapper.c:825-914 of
the reference), packed block loading, mask tracks (whole-DB and per block), -G worker processes and their exit
codes, the sort directory, the LAsort/LAcat/LAmerge commands or the built-in stand-in, Clean_Exit.  The same
cases run against the real library under -m gpu (tests/test_cli_gpu.py)."""
import glob
import os
import subprocess

import pytest

from conftest import ROOT
from test_cli_gpu import _run_blocks, _three_block_case


def _have_ref():
    from oracle import run_ref
    return run_ref.have_ref()


pytestmark = pytest.mark.skipif(not _have_ref(), reason="oracle/_ref (compiled reference + stubs) not present")


@pytest.fixture(scope="module")
def mock_exe(tmp_path_factory, oracle_mod):
    exe = str(tmp_path_factory.mktemp("mockdrv") / "damapper")
    host = os.path.join(ROOT, "damapper_b200", "host")
    orc = os.path.join(ROOT, "oracle")
    subprocess.check_call(["gcc", "-O2", "-Wall", "-Wextra", "-Wno-unused-result"] +
                          os.environ.get("DAMGPU_TEST_CFLAGS", "").split() + ["-o", exe,
                           os.path.join(host, "damapper.c"), os.path.join(host, "dazz_db.c"),
                           os.path.join(host, "las_post.c"), os.path.join(orc, "mock_damgpu.c"),
                           "-L" + orc, "-loracle", "-Wl,-rpath," + orc, "-lm", "-lpthread"])
    return exe


def test_block_loop_against_the_reference(mock_exe, tmp_path):
    """Three reads blocks in one invocation (the reference indices stay resident from the second block on) against
    three invocations of the reference: M and R record streams and the -p track of every block."""
    from oracle import run_ref
    wd = str(tmp_path)
    _three_block_case(wd)
    got, p = _run_blocks(wd, "mock", mock_exe, ["-C", "-p"])
    want, _ = _run_blocks(wd, "ref", run_ref.REF_BIN, ["-C", "-p"])
    assert sum(len(v[0]) for v in want.values()) > 10000 and all(len(v[2]) > 0 for v in want.values())
    assert got == want
    # the reference DB in two blocks: streamed block by block, the second reads block finds them resident
    _three_block_case(wd, nblocks_ref=2)
    got2, _ = _run_blocks(wd, "mock2", mock_exe, ["-C"], blocks=(1, 2))
    want2, _ = _run_blocks(wd, "ref2", run_ref.REF_BIN, ["-C"], blocks=(1, 2))
    assert got2 == want2
    # DAMGPU_REF_CACHE=0: nothing kept, every reads block loads the reference again (as the reference does)
    got3, _ = _run_blocks(wd, "mock3", mock_exe, ["-C"], {"DAMGPU_REF_CACHE": "0"}, blocks=(1, 2))
    assert got3 == want2


def test_worker_processes(mock_exe, tmp_path):
    """-G<n>: n worker processes, reads blocks round robin; more workers than blocks; a failing worker makes the
    command fail; a device list that is too short is refused before anything is forked."""
    wd = str(tmp_path)
    _three_block_case(wd)
    one, _ = _run_blocks(wd, "g1", mock_exe, ["-C", "-p"])
    two, _ = _run_blocks(wd, "g2", mock_exe, ["-C", "-p", "-G2"], {"DAMGPU_DEVICES": "0,1"})
    many, _ = _run_blocks(wd, "g8", mock_exe, ["-C", "-p", "-G8"], {"DAMGPU_DEVICES": "0,0,0"})
    assert one == two == many
    q = subprocess.run([mock_exe, "-G3", "ref.dam", "reads.1", "reads.2", "reads.3"], cwd=wd,
                       env=dict(os.environ, DAMGPU_DEVICES="0"), capture_output=True, text=True)
    assert q.returncode == 1 and "only 1 devices" in q.stderr
    os.makedirs(os.path.join(wd, "tmpf"))
    q = subprocess.run([mock_exe, "-G2", "-P" + os.path.join(wd, "tmpf"), "ref.dam", "reads.1", "reads.9"], cwd=wd,
                       env=dict(os.environ, DAMGPU_BUILTIN_SORT="1", DAMGPU_DEVICES="0,1"), capture_output=True, text=True)
    assert q.returncode == 1, "a worker that cannot open its block must fail the command"
    assert glob.glob(os.path.join(wd, "tmpf", "damapper.*")) == [], "Clean_Exit removes every worker's sort directory"


def test_mask_tracks_whole_db_and_per_block(mock_exe, tmp_path):
    """-m: whole-DB track against the reference; the same intervals stored per block (what DBdust leaves when it
    is run on blocks); a track no block has is reported as never used."""
    from damapper_b200 import dazzdb
    from oracle import run_ref
    wd = str(tmp_path)
    contigs, rb, rl, cuts = _three_block_case(wd, seed=62)
    off, pts = dazzdb.random_masks(rl, seed=5, max_intervals=4, max_len=600)
    dazzdb.write_mask_track(os.path.join(wd, "reads.db"), "dust", off, pts)
    whole, _ = _run_blocks(wd, "whole", mock_exe, ["-C", "-mdust"])
    ref, _ = _run_blocks(wd, "ref", run_ref.REF_BIN, ["-C", "-mdust"])
    assert whole == ref
    nomask, p = _run_blocks(wd, "nomask", mock_exe, ["-C", "-mnone"])
    assert nomask != whole and "Track none given but never used" in p.stdout
    for f in glob.glob(os.path.join(wd, ".reads.dust.*")):
        os.remove(f)
    for b, (a, e) in enumerate(zip(cuts[:-1], cuts[1:]), start=1):
        dazzdb.write_mask_track(os.path.join(wd, "reads.%d" % b), "dust", off[a:e + 1] - off[a], pts[off[a]:off[e]])
    per_block, p = _run_blocks(wd, "blocks", mock_exe, ["-C", "-mdust"])
    assert "never used" not in p.stdout
    assert per_block == whole


def test_builtin_sort_leaves_the_final_files(mock_exe, tmp_path):
    """No LAsort on PATH: the driver sorts and merges the per-thread files itself and leaves
    <reads>.<ref>.las / <ref>.<reads>.las (damapper.c:893-911), every chain of the reference's stream in them."""
    from damapper_b200 import las
    from oracle import run_ref
    wd = str(tmp_path)
    _three_block_case(wd)
    want, _ = _run_blocks(wd, "ref", run_ref.REF_BIN, ["-C"], blocks=(2,))
    os.makedirs(os.path.join(wd, "tmpb"))
    q = subprocess.run([mock_exe, "-T4", "-C", "-M16", "-P" + os.path.join(wd, "tmpb"), "ref.dam", "reads.2"], cwd=wd,
                       env=dict(os.environ, DAMGPU_BUILTIN_SORT="1"), capture_output=True, text=True)
    assert q.returncode == 0, q.stderr
    assert glob.glob(os.path.join(wd, "tmpb", "damapper.*")) == [], "the sort directory is removed at the end"
    for name, fam in (("reads.2.ref.las", 0), ("ref.reads.2.las", 1)):
        ts, recs = las.read_las(os.path.join(wd, name))
        assert ts == 100
        blob = lambda r: las.REC.pack(r["tlen"], r["diffs"], r["abpos"], r["bbpos"], r["aepos"], r["bepos"],
                                      r["flags"], r["aread"], r["bread"]) + r["trace"].tobytes()
        ref_recs = las.stream_records(want[2][fam], 100)
        assert len(recs) == len(ref_recs) >= 30
        assert sorted(blob(r) for r in recs) == sorted(blob(r) for r in ref_recs), name


def test_reads_shorter_than_k_and_plain_db(mock_exe, tmp_path):
    """damapper.c:403-410: a block holding a read shorter than k stops the run with the reference's message; an
    unsplit reads DB (no block suffix) is mapped whole."""
    import numpy as np
    from damapper_b200 import dazzdb, synth
    from oracle import run_ref
    wd = str(tmp_path)
    contigs, rb, rl = synth.make_config("C1", scale=0.03, seed=67)
    dazzdb.write_db(os.path.join(wd, "ref.dam"), contigs, is_dam=True)
    dazzdb.write_db(os.path.join(wd, "reads.db"), (rb, rl))
    r = run_ref.run_damapper(wd, "ref.dam", "reads.db", flags=("-M16",), threads=4, exe=mock_exe)
    want = run_ref.run_damapper(wd, "ref.dam", "reads.db", flags=("-M16",), threads=4)
    from damapper_b200 import las
    assert las.canonical_stream(r["m_files"]) == las.canonical_stream(want["m_files"])
    short = np.concatenate([rb, np.zeros(12, dtype=np.uint8)])
    dazzdb.write_db(os.path.join(wd, "bad.db"), (short, np.concatenate([rl, [12]]).astype(rl.dtype)))
    os.makedirs(os.path.join(wd, "tmps"))
    cmd = ["-T4", "-P" + os.path.join(wd, "tmps"), "ref.dam", "bad.db"]
    q = subprocess.run([mock_exe] + cmd, cwd=wd, capture_output=True, text=True)
    w = subprocess.run([run_ref.REF_BIN] + cmd, cwd=wd, capture_output=True, text=True)
    assert q.returncode == w.returncode == 1
    assert "contains reads < 20bp long" in q.stderr and "contains reads < 20bp long" in w.stderr
    assert glob.glob(os.path.join(wd, "tmps", "damapper.*")) == []
