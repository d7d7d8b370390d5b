"""CPU test of the host driver's built-in LAsort/LAcat/LAmerge stand-in (damapper_b200/host/las_post.c,
SURVEY 8(f)1): chains stay whole, the -a key orders them, nothing is lost.  The real DALIGNER programs
are not in the reference tree, so this pins the documented behaviour, not their output."""
import ctypes as C
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import ROOT

GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def post(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("laspost") / "las_post.so")
    subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC"] + os.environ.get("DAMGPU_TEST_CFLAGS", "").split() +
                          ["-o", so, os.path.join(ROOT, "damapper_b200", "host", "las_post.c")])
    L = C.CDLL(so)
    L.las_sort_cat.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int]
    L.las_sort_cat.restype = C.c_int
    L.on_path.argtypes = [C.c_char_p]
    L.on_path.restype = C.c_int
    return L


def _chains(recs):
    out = []
    for r in recs:
        if r["flags"] & 0x8 and out:
            out[-1].append(r)
        else:
            out.append([r])
    return out


def _blob(r, tb):
    from damapper_b200 import las
    return las.REC.pack(r["tlen"], r["diffs"], r["abpos"], r["bbpos"], r["aepos"], r["bepos"], r["flags"],
                        r["aread"], r["bread"]) + r["trace"].astype(np.uint8 if tb == 1 else np.uint16).tobytes()


@pytest.mark.parametrize("case,fam,map_order", [("c5_cover_profile", "a", 1), ("c5_cover_profile", "b", 1),
                                                ("c3_repeat_n95", "a", 0), ("c1_k24_s200", "a", 1)])
def test_builtin_sort_keeps_chains_and_orders_them(post, tmp_path, case, fam, map_order):
    from damapper_b200 import las
    from oracle.make_golden import CASES
    g = np.load(os.path.join(GOLDEN, case + ".npz"))
    spacing = CASES[case][4].get("spacing", 100)
    tb = 1 if spacing <= 125 else 2
    recs = las.stream_records(g[fam].tobytes(), spacing)
    if not recs:
        pytest.skip("fixture has no records of this family")
    ch = _chains(recs)
    nfiles = 4
    cut = [len(ch) * i // nfiles for i in range(nfiles + 1)]
    for i in range(nfiles):
        part = [r for c in ch[cut[i]:cut[i + 1]] for r in c]
        with open(tmp_path / ("x.y.M%d.las" % (i + 1)), "wb") as f:
            f.write(struct.pack("<qi", len(part), spacing))
            for r in part:
                f.write(_blob(r, tb))
    out = str(tmp_path / "x.y.las")
    assert post.las_sort_cat(str(tmp_path / "x.y.M").encode(), nfiles, out.encode(), map_order, 0) == 0
    ts, got = las.read_las(out)
    assert ts == spacing and len(got) == len(recs)
    gch = _chains(got)
    key = lambda c: tuple(_blob(r, tb) for r in c)
    assert sorted(map(key, gch)) == sorted(map(key, ch))
    if map_order:
        heads = [(c[0]["aread"], c[0]["abpos"]) for c in gch]
    else:
        heads = [(c[0]["aread"], c[0]["bread"], c[0]["flags"] & 1, c[0]["abpos"]) for c in gch]
    assert heads == sorted(heads)
    assert las.check_trace_points(got, ts) == 0


def test_builtin_sort_errors_and_path_probe(post, tmp_path):
    assert post.las_sort_cat(str(tmp_path / "none.M").encode(), 2, str(tmp_path / "o.las").encode(), 1, 0) == 1
    with open(tmp_path / "t.M1.las", "wb") as f:
        f.write(struct.pack("<qi", 3, 100))               # claims 3 records, holds none
    assert post.las_sort_cat(str(tmp_path / "t.M").encode(), 1, str(tmp_path / "o.las").encode(), 1, 0) == 1
    with open(tmp_path / "e.M1.las", "wb") as f:
        f.write(struct.pack("<qi", 0, 100))               # an empty block is fine
    assert post.las_sort_cat(str(tmp_path / "e.M").encode(), 1, str(tmp_path / "e.las").encode(), 1, 0) == 0
    assert open(tmp_path / "e.las", "rb").read() == struct.pack("<qi", 0, 100)
    assert post.on_path(b"sh") == 1 and post.on_path(b"no-such-program-xyz") == 0


def test_builtin_merge_over_reads_blocks(post, tmp_path):
    """The R family of several reads blocks (damapper.c:903-911 runs LAsort + LAmerge per reads block; merging
    the per-block ref.reads.<b>.las files is one more LAmerge, as HPC.damapper's scripts do): two blocks'
    worth of per-thread files are merged block by block, the two results merged again.  Nothing is lost, chains
    stay whole, the -a key orders the final file, and merging an already merged file changes nothing."""
    import shutil
    from damapper_b200 import las
    from oracle.make_golden import CASES
    g = np.load(os.path.join(GOLDEN, "c5_cover_profile.npz"))
    spacing = CASES["c5_cover_profile"][4].get("spacing", 100)
    recs = las.stream_records(g["b"].tobytes(), spacing)
    ch = _chains(recs)
    assert len(ch) > 100
    # two "reads blocks": chains dealt alternately (both hold every contig), each written as 4 per-thread files
    blocks = [ch[0::2], ch[1::2]]
    for b, part in enumerate(blocks, start=1):
        for i in range(4):
            sub = [r for c in part[len(part) * i // 4:len(part) * (i + 1) // 4] for r in c]
            with open(tmp_path / ("ref.reads%d.R%d.las" % (b, i + 1)), "wb") as f:
                f.write(struct.pack("<qi", len(sub), spacing))
                for r in sub:
                    f.write(_blob(r, 1))
        out = str(tmp_path / ("blk.R%d.las" % b))
        assert post.las_sort_cat(str(tmp_path / ("ref.reads%d.R" % b)).encode(), 4, out.encode(), 1, 0) == 0
    final = str(tmp_path / "ref.reads.las")
    assert post.las_sort_cat(str(tmp_path / "blk.R").encode(), 2, final.encode(), 1, 0) == 0
    ts, got = las.read_las(final)
    gch = _chains(got)
    key = lambda c: tuple(_blob(r, 1) for r in c)
    assert ts == spacing and len(got) == len(recs)
    assert sorted(map(key, gch)) == sorted(map(key, ch))
    heads = [(c[0]["aread"], c[0]["abpos"]) for c in gch]
    assert heads == sorted(heads)
    # idempotence: a merged file merged again is the same file
    shutil.copy(final, tmp_path / "again.R1.las")
    again = str(tmp_path / "again.las")
    assert post.las_sort_cat(str(tmp_path / "again.R").encode(), 1, again.encode(), 1, 0) == 0
    assert open(again, "rb").read() == open(final, "rb").read()
