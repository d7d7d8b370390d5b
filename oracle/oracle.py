"""TEST INFRASTRUCTURE ONLY: ctypes binding of the C oracle (oracle/liboracle.so).

Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs only.  The product path (damapper_b200) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")

KMER_DT = np.dtype([("code", "<u8"), ("rpos", "<i4"), ("read", "<i4")])
SEED_DT = np.dtype([("diag", "<i4"), ("apos", "<i4"), ("bread", "<i4"), ("aread", "<i4")])
CAND_DT = np.dtype([(n, "<i4") for n in ("read", "score", "length", "bread", "comp", "afirst",
                                         "alast", "bfirst", "blast")])


class Block(C.Structure):
    _fields_ = [("bases", C.c_void_p), ("boff", C.c_void_p), ("rlen", C.c_void_p),
                ("nreads", C.c_int32), ("tfirst", C.c_int32), ("maxlen", C.c_int32),
                ("totlen", C.c_int64), ("sizeof_db", C.c_int64),
                ("mask_off", C.c_void_p), ("mask_pts", C.c_void_p)]


class Params(C.Structure):
    _fields_ = [("kmer", C.c_int32), ("suppress", C.c_int32), ("spacing", C.c_int32),
                ("profile", C.c_int32), ("ave_corr", C.c_double), ("best_tie", C.c_double),
                ("freq", C.c_float * 4), ("mem_limit", C.c_uint64),
                ("do_a", C.c_int32), ("do_b", C.c_int32)]


def build(force: bool = False) -> str:
    srcs = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".c", ".h"))
            and f != "ref_tap.c"]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB)
                                               for s in srcs):
        subprocess.check_call(["make", "-C", HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        L.orc_sort_kmers.restype = C.c_void_p
        L.orc_sort_kmers.argtypes = [C.POINTER(Block), C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.orc_merge_join.restype = C.c_void_p
        L.orc_merge_join.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_uint64,
                                     C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int,
                                     C.POINTER(C.c_int64), C.POINTER(C.c_int), C.c_void_p]
        L.orc_align_spec.argtypes = [C.c_double, C.POINTER(C.c_float), C.POINTER(C.c_int),
                                     C.c_void_p, C.c_void_p]
        L.orc_mapper_new.restype = C.c_void_p
        L.orc_mapper_new.argtypes = [C.POINTER(Params), C.POINTER(Block)]
        L.orc_mapper_free.argtypes = [C.c_void_p]
        L.orc_match_filter.argtypes = [C.c_void_p, C.POINTER(Block), C.c_int, C.c_int]
        L.orc_chain_seeds.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int]
        L.orc_num_candidates.restype = C.c_int64
        L.orc_num_candidates.argtypes = [C.c_void_p]
        L.orc_get_candidates.restype = C.c_int64
        L.orc_get_candidates.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        L.orc_get_cover.restype = C.c_int64
        L.orc_get_cover.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        L.orc_report.argtypes = [C.c_void_p, C.POINTER(Block)] + [C.c_void_p] * 8
        L.orc_report_stats.argtypes = [C.c_void_p] + [C.POINTER(C.c_int64)] * 4
        L.orc_local_alignment.restype = C.c_int
        L.orc_local_alignment.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                          C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_int]
        L.free = C.CDLL(None).free
        L.free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


class HostBlock:
    """A loaded DB block (Load_All_Reads image) kept alive for ctypes calls."""

    def __init__(self, bases: np.ndarray, boff: np.ndarray, rlen: np.ndarray, tfirst: int = 0,
                 path_len: int = 16, mask=None):
        # `bases` carries the leading 4 at index 0 (damapper_b200.dazzdb.load_block)
        self.bases = np.ascontiguousarray(bases, dtype=np.uint8)
        self.boff = np.ascontiguousarray(boff, dtype=np.int64)
        self.rlen = np.ascontiguousarray(rlen, dtype=np.int32)
        self.nreads = int(self.rlen.size)
        self.tfirst = tfirst
        self.maxlen = int(self.rlen.max()) if self.nreads else 0
        self.totlen = int(self.rlen.sum())
        # sizeof_DB (DB.c:1044-1051): sizeof(DAZZ_DB)=112, sizeof(DAZZ_READ)=40
        self.sizeof_db = 112 + 40 * (self.nreads + 2) + path_len + 1 + (self.totlen + self.nreads + 4)
        self.mask_off = self.mask_pts = None
        if mask is not None:
            self.mask_off = np.ascontiguousarray(mask[0], dtype=np.int64)
            self.mask_pts = np.ascontiguousarray(np.concatenate([mask[1], [0]]), dtype=np.int32)
        self.c = Block(self.bases.ctypes.data + 1, self.boff.ctypes.data, self.rlen.ctypes.data,
                       self.nreads, tfirst, self.maxlen, self.totlen, self.sizeof_db,
                       self.mask_off.ctypes.data if mask is not None else None,
                       self.mask_pts.ctypes.data if mask is not None else None)


def sort_kmers(blk: HostBlock, kmer: int, suppress: int = 0) -> np.ndarray:
    n = C.c_int(0)
    p = lib().orc_sort_kmers(C.byref(blk.c), kmer, suppress, C.byref(n))
    if not p:
        return np.zeros(0, dtype=KMER_DT)
    out = np.frombuffer((C.c_char * ((n.value + 2) * 16)).from_address(p), dtype=KMER_DT).copy()
    lib().free(p)
    return out      # includes the two sentinels


def merge_join(aidx: np.ndarray, bidx: np.ndarray, mem_limit: int, asize: int, bsize: int,
               amaxlen: int, anreads: int, bnreads: int):
    """aidx = reads index, bidx = reference index (both with sentinels).
    Returns (seeds incl. sentinel, nhits, limit, histogram)."""
    nh = C.c_int64(0)
    lim = C.c_int(0)
    histo = np.zeros(10000, dtype=np.int64)
    p = lib().orc_merge_join(aidx.ctypes.data, len(aidx) - 2, bidx.ctypes.data, len(bidx) - 2,
                             mem_limit, asize, bsize, amaxlen, anreads, bnreads,
                             C.byref(nh), C.byref(lim), histo.ctypes.data)
    if not p:
        return np.zeros(0, dtype=SEED_DT), 0, 0, histo
    out = np.frombuffer((C.c_char * ((nh.value + 1) * 16)).from_address(p), dtype=SEED_DT).copy()
    lib().free(p)
    return out, nh.value, lim.value, histo


def align_spec(ave_corr: float, freq):
    f = (C.c_float * 4)(*freq)
    ap = C.c_int(0)
    score = np.zeros(32768, dtype=np.int16)
    table = np.zeros(32768, dtype=np.int16)
    lib().orc_align_spec(ave_corr, f, C.byref(ap), score.ctypes.data, table.ctypes.data)
    return ap.value, score, table


def local_alignment(aseq: np.ndarray, bseq: np.ndarray, acomp: int, dg: int, ad: int,
                    spacing: int, ave_path: int, score, table):
    """aseq/bseq: uint8 arrays bracketed by 4 at both ends (index 0 and -1)."""
    ap = (C.c_int * 6)()
    bp = (C.c_int * 6)()
    tcap = 2 * (max(len(aseq), len(bseq)) // spacing + 4) * 2
    at = np.zeros(tcap, dtype=np.uint16)
    bt = np.zeros(tcap, dtype=np.uint16)
    rc = lib().orc_local_alignment(aseq.ctypes.data + 1, len(aseq) - 2, bseq.ctypes.data + 1,
                                   len(bseq) - 2, acomp, dg, ad, spacing, ave_path,
                                   score.ctypes.data, table.ctypes.data, ap, at.ctypes.data,
                                   bp, bt.ctypes.data, tcap)
    assert rc == 0
    return list(ap), at[:ap[5]].copy(), list(bp), bt[:bp[5]].copy()


class Mapper:
    def __init__(self, reads: HostBlock, kmer=20, suppress=0, spacing=100, profile=0,
                 ave_corr=0.85, best_tie=1.0, freq=(.25, .25, .25, .25), mem_limit=64 << 30,
                 do_a=1, do_b=0):
        self.reads = reads
        self.par = Params(kmer, suppress, spacing, profile, ave_corr, best_tie,
                          (C.c_float * 4)(*freq), mem_limit, do_a, do_b)
        self.h = lib().orc_mapper_new(C.byref(self.par), C.byref(reads.c))

    def close(self):
        if self.h:
            lib().orc_mapper_free(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def match_filter(self, ref: HostBlock, comp: int, start: int):
        lib().orc_match_filter(self.h, C.byref(ref.c), comp, start)

    def chain_seeds(self, seeds: np.ndarray, nhits: int, bstart: int, comp: int, start: int):
        lib().orc_chain_seeds(self.h, seeds.ctypes.data, nhits, bstart, comp, start)

    def candidates(self):
        n = lib().orc_num_candidates(self.h)
        out = np.zeros(n, dtype=CAND_DT)
        jcnt = np.zeros(n, dtype=np.int32)
        nj = lib().orc_get_candidates(self.h, out.ctypes.data, jcnt.ctypes.data, None, 0)
        jumps = np.zeros((max(nj, 1), 2), dtype=np.int32)
        lib().orc_get_candidates(self.h, out.ctypes.data, jcnt.ctypes.data, jumps.ctypes.data, nj)
        return out, jcnt, jumps[:nj]

    def cover(self):
        n = lib().orc_get_cover(self.h, None, 0)
        out = np.zeros(n, dtype=np.int16)
        lib().orc_get_cover(self.h, out.ctypes.data, n)
        return out

    def report(self, wholeref: HostBlock):
        ab, bb, pf = C.c_void_p(), C.c_void_p(), C.c_void_p()
        al, an, bl, bn, pl = (C.c_int64() for _ in range(5))
        lib().orc_report(self.h, C.byref(wholeref.c), C.byref(ab), C.byref(al), C.byref(an),
                         C.byref(bb), C.byref(bl), C.byref(bn), C.byref(pf), C.byref(pl))
        get = lambda p, n: C.string_at(p, n.value) if n.value else b""
        return dict(a=get(ab, al), anrec=an.value, b=get(bb, bl), bnrec=bn.value, prof=get(pf, pl))

    def stats(self):
        v = [C.c_int64() for _ in range(4)]
        lib().orc_report_stats(self.h, *[C.byref(x) for x in v])
        return dict(nalign=v[0].value, nwaves=v[1].value, ncells=v[2].value, h2=v[3].value)


def map_block(reads: HostBlock, ref_blocks, wholeref: HostBlock, **kw):
    """Whole damapper flow for one reads block (damapper.c:825-879): for every reference
    block, forward then complement Match_Filter, then Reporter.  `ref_blocks` is a list of
    (forward HostBlock, complemented HostBlock)."""
    m = Mapper(reads, **kw)
    for k, (fwd, rev) in enumerate(ref_blocks):
        m.match_filter(fwd, 0, 1 if k == 0 else 0)
        m.match_filter(rev, 1, 0)
    out = m.report(wholeref)
    out["stats"] = m.stats()
    out["candidates"] = m.candidates()
    m.close()
    return out
