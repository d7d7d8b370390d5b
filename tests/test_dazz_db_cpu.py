"""CPU tests of the host driver's DAZZ_DB reader (damapper_b200/host/dazz_db.c): the block image it hands
to the library must be the one Load_All_Reads builds (reference DB.c:1389-1441) -- bases with 4-terminators,
boff, rlen, tfirst, totlen, maxlen -- from the plain and from the 2-bit packed load, for whole DBs and for
blocks of a split DB; mask tracks (whole-DB files, several tracks merged, damapper.c:352-399) and the in-place
complement (damapper.c:433-469) as well.  No GPU involved: the file is compiled on its own with gcc."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT


class DazzBlock(C.Structure):
    _fields_ = [("root", C.c_char_p), ("pwd", C.c_char_p), ("isdam", C.c_int), ("nblocks", C.c_int),
                ("part", C.c_int), ("freq", C.c_float * 4), ("cutoff", C.c_int), ("all", C.c_int),
                ("nreads", C.c_int), ("tfirst", C.c_int), ("maxlen", C.c_int), ("totlen", C.c_int64),
                ("raw", C.POINTER(C.c_uint8)), ("packed", C.POINTER(C.c_uint8)), ("poff", C.POINTER(C.c_int64)),
                ("packed_bytes", C.c_int64), ("boff", C.POINTER(C.c_int64)), ("rlen", C.POINTER(C.c_int32)),
                ("path_len", C.c_int64), ("ufirst", C.c_int), ("ulast", C.c_int), ("db_ureads", C.c_int),
                ("db_treads", C.c_int), ("kept", C.POINTER(C.c_uint8)), ("mask_off", C.POINTER(C.c_int64)),
                ("mask_pts", C.POINTER(C.c_int32))]


@pytest.fixture(scope="module")
def dz(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("dazz") / "dazz_db.so")
    # DAMGPU_TEST_CFLAGS="-g -fsanitize=address,undefined" (with libasan preloaded) runs the reader sanitized
    subprocess.check_call(["gcc", "-O2", "-Wall", "-shared", "-fPIC"] + os.environ.get("DAMGPU_TEST_CFLAGS", "").split() +
                          ["-o", so, os.path.join(ROOT, "damapper_b200", "host", "dazz_db.c")])
    L = C.CDLL(so)
    for f in ("dazz_open", "dazz_load", "dazz_load_packed"):
        getattr(L, f).argtypes = [C.c_char_p, C.POINTER(DazzBlock)]
        getattr(L, f).restype = C.c_int
    L.dazz_add_mask.argtypes = [C.POINTER(DazzBlock), C.c_char_p, C.c_char_p]
    L.dazz_add_mask.restype = C.c_int
    L.dazz_complement.argtypes = [C.POINTER(DazzBlock)]
    L.dazz_close.argtypes = [C.POINTER(DazzBlock)]
    return L


def _arr(ptr, n, dt):
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dt, copy=True) if n else np.zeros(0, dt)


def _image(b):
    """(bases incl. the leading 4, boff, rlen) of a plainly loaded block."""
    n = b.nreads
    boff = _arr(b.boff, n + 1, np.int64)
    return _arr(b.raw, int(boff[n]) + 1, np.uint8), boff, _arr(b.rlen, n, np.int32)


def _unpack(b):
    n = b.nreads
    rlen = _arr(b.rlen, n, np.int32)
    poff = _arr(b.poff, n, np.int64)
    pk = _arr(b.packed, int(b.packed_bytes), np.uint8)
    out = []
    for i in range(n):
        q = pk[poff[i]: poff[i] + (rlen[i] + 3) // 4]
        bases = np.stack([(q >> 6) & 3, (q >> 4) & 3, (q >> 2) & 3, q & 3], axis=1).reshape(-1)[: rlen[i]]
        out.append(bases.astype(np.uint8))
    return out


def _reads(seed, n=23):
    rng = np.random.default_rng(seed)
    rl = rng.integers(30, 900, size=n).astype(np.int32)      # lengths of every residue mod 4
    rb = rng.integers(0, 4, size=int(rl.sum()), dtype=np.uint8)
    return rb, rl


def test_whole_db_image_plain_and_packed(dz, tmp_path):
    from damapper_b200 import dazzdb
    rb, rl = _reads(1)
    dazzdb.write_db(str(tmp_path / "reads.db"), (rb, rl))
    want_bases, want_boff, want_rlen = dazzdb.load_block((rb, rl))
    b = DazzBlock()
    assert dz.dazz_load(str(tmp_path / "reads.db").encode(), C.byref(b)) == 0
    bases, boff, rlen = _image(b)
    n = len(rl)
    assert b.nreads == n and b.tfirst == 0 and b.part == 0 and b.isdam == 0
    assert b.totlen == int(rl.sum()) and b.maxlen == int(rl.max())
    assert (rlen == want_rlen).all() and (boff == want_boff[: n + 1]).all()
    assert bases.tobytes() == want_bases[: len(bases)].tobytes()
    assert bases[0] == 4 and all(bases[1 + boff[i] + rlen[i]] == 4 for i in range(n))
    dz.dazz_close(C.byref(b))

    p = DazzBlock()
    assert dz.dazz_load_packed(str(tmp_path / "reads").encode(), C.byref(p)) == 0     # the extension is optional
    assert not p.raw and p.nreads == n
    assert (_arr(p.boff, n + 1, np.int64) == want_boff[: n + 1]).all()
    off = np.concatenate([[0], np.cumsum(rl)])
    for i, r in enumerate(_unpack(p)):
        assert r.tobytes() == rb[off[i]: off[i + 1]].tobytes()
    dz.dazz_close(C.byref(p))


def test_blocks_of_a_split_dam(dz, tmp_path):
    from damapper_b200 import dazzdb
    rng = np.random.default_rng(2)
    contigs = [rng.integers(0, 4, size=int(s), dtype=np.uint8) for s in (700, 1201, 50, 333, 2048, 91, 640)]
    dazzdb.write_db(str(tmp_path / "ref.dam"), contigs, is_dam=True, nblocks=3)
    hdr = DazzBlock()
    assert dz.dazz_open(str(tmp_path / "ref.dam").encode(), C.byref(hdr)) == 0
    assert hdr.isdam == 1 and hdr.nblocks == 3 and hdr.root == b"ref"
    bounds = [int(7 * i / 3) for i in range(4)]
    for k in range(1, 4):
        b = DazzBlock()
        assert dz.dazz_load(str(tmp_path / ("ref.%d" % k)).encode(), C.byref(b)) == 0
        part = contigs[bounds[k - 1]: bounds[k]]
        want_bases, want_boff, want_rlen = dazzdb.load_block(part)
        bases, boff, rlen = _image(b)
        assert b.part == k and b.tfirst == bounds[k - 1] and b.nreads == len(part)
        assert b.totlen == sum(c.size for c in part) and b.maxlen == max(c.size for c in part)
        assert (rlen == want_rlen).all() and bases.tobytes() == want_bases[: len(bases)].tobytes()
        dz.dazz_close(C.byref(b))
    whole = DazzBlock()
    assert dz.dazz_load(str(tmp_path / "ref.dam").encode(), C.byref(whole)) == 0
    assert whole.nreads == 7 and whole.part == 0 and whole.tfirst == 0
    dz.dazz_close(C.byref(whole))


def test_complement_in_place(dz, tmp_path):
    from damapper_b200 import dazzdb
    rng = np.random.default_rng(3)
    contigs = [rng.integers(0, 4, size=int(s), dtype=np.uint8) for s in (101, 64, 1, 999)]
    dazzdb.write_db(str(tmp_path / "g.dam"), contigs, is_dam=True)
    b = DazzBlock()
    assert dz.dazz_load(str(tmp_path / "g.dam").encode(), C.byref(b)) == 0
    dz.dazz_complement(C.byref(b))
    bases, boff, rlen = _image(b)
    want = dazzdb.load_block(dazzdb.revcomp_contigs(contigs))[0]
    assert bases.tobytes() == want[: len(bases)].tobytes()
    dz.dazz_complement(C.byref(b))
    assert _image(b)[0].tobytes() == dazzdb.load_block(contigs)[0][: len(bases)].tobytes()
    dz.dazz_close(C.byref(b))


def test_mask_tracks_are_merged(dz, tmp_path):
    """-mdust -mtan: the union of both tracks (damapper.c:181-343 merge_tracks), a track the DB does not have is
    reported as unused (0), offsets count ints (damapper.c:385-387)."""
    from damapper_b200 import dazzdb
    rb, rl = _reads(4, n=40)
    stub = dazzdb.write_db(str(tmp_path / "reads.db"), (rb, rl))
    dust = dazzdb.random_masks(rl, seed=5, max_intervals=3, max_len=200)
    tan = dazzdb.random_masks(rl, seed=6, max_intervals=2, max_len=120)
    dazzdb.write_mask_track(stub, "dust", *dust)
    dazzdb.write_mask_track(stub, "tan", *tan)
    b = DazzBlock()
    assert dz.dazz_load_packed(stub.encode(), C.byref(b)) == 0
    assert dz.dazz_add_mask(C.byref(b), b"dust", b"damapper") == 1
    off = _arr(b.mask_off, b.nreads + 1, np.int64)
    assert (off == np.asarray(dust[0])).all()
    assert (_arr(b.mask_pts, int(off[-1]), np.int32) == np.asarray(dust[1])).all()
    assert dz.dazz_add_mask(C.byref(b), b"nosuch", b"damapper") == 0
    assert dz.dazz_add_mask(C.byref(b), b"tan", b"damapper") == 1
    uoff, upts = dazzdb.union_masks(dust, tan)
    off = _arr(b.mask_off, b.nreads + 1, np.int64)
    pts = _arr(b.mask_pts, int(off[-1]), np.int32)
    # the union may keep zero-length gaps between abutting intervals: compare covered bases
    for i in range(b.nreads):
        def cover(o, p):
            m = np.zeros(int(rl[i]) + 1, dtype=bool)
            for j in range(int(o[i]), int(o[i + 1]), 2):
                m[p[j]: p[j + 1]] = True
            return m
        assert (cover(off, pts) == cover(np.asarray(uoff), np.asarray(upts))).all(), i
    dz.dazz_close(C.byref(b))


def test_missing_and_damaged_files(dz, tmp_path, capfd):
    from damapper_b200 import dazzdb
    b = DazzBlock()
    assert dz.dazz_load(str(tmp_path / "nothing.db").encode(), C.byref(b)) != 0
    rb, rl = _reads(7, n=5)
    dazzdb.write_db(str(tmp_path / "r.db"), (rb, rl))
    os.remove(tmp_path / ".r.bps")
    assert dz.dazz_load(str(tmp_path / "r.db").encode(), C.byref(b)) != 0
    dazzdb.write_db(str(tmp_path / "s.db"), (rb, rl))
    with open(tmp_path / ".s.idx", "r+b") as f:
        f.truncate(150)                                   # header + part of the first read record
    assert dz.dazz_load(str(tmp_path / "s.db").encode(), C.byref(b)) != 0
    assert dz.dazz_load(str(tmp_path / "r.9").encode(), C.byref(b)) != 0               # no such block
    capfd.readouterr()
