"""ctypes binding of libdamgpu.so (include/libdamgpu.h) -- the host-side mirror of the
reference's map.h interface (Set_Filter_Params / Sort_Kmers / Match_Filter / Reporter).

There is no CPU fallback: importing works without a GPU (so that the symbol table can be
checked), every compute call requires a B200 and fails loudly otherwise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DAMGPU_LIB") or os.path.join(HERE, "libdamgpu.so")   # DAMGPU_LIB: build variants (dev)

KMER_DT = np.dtype([("code", "<u8"), ("rpos", "<i4"), ("read", "<i4")])
SEED_DT = np.dtype([("diag", "<i4"), ("apos", "<i4"), ("bread", "<i4"), ("aread", "<i4")])
CAND_DT = np.dtype([(n, "<i4") for n in ("read", "score", "length", "bread", "comp", "afirst",
                                         "alast", "bfirst", "blast")])


class CBlock(C.Structure):
    _fields_ = [("bases", C.c_void_p), ("boff", C.c_void_p), ("rlen", C.c_void_p),
                ("nreads", C.c_int32), ("tfirst", C.c_int32), ("maxlen", C.c_int32),
                ("totlen", C.c_int64), ("sizeof_db", C.c_int64),
                ("mask_off", C.c_void_p), ("mask_pts", C.c_void_p),
                ("packed", C.c_void_p), ("poff", C.c_void_p), ("packed_bytes", C.c_int64)]


class COptions(C.Structure):
    _fields_ = [("verbose", C.c_int32), ("profile", C.c_int32), ("spacing", C.c_int32),
                ("best_tie", C.c_double), ("sort_path", C.c_char_p), ("mem_limit", C.c_uint64),
                ("mem_physical", C.c_uint64)]


class CAlignSpec(C.Structure):
    _fields_ = [("ave_corr", C.c_double), ("trace_space", C.c_int32), ("freq", C.c_float * 4)]


# every symbol include/libdamgpu.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "damgpu_init": (C.c_int, [C.c_int]),
    "damgpu_set_options": (None, [C.POINTER(COptions)]),
    "damgpu_set_fatal": (None, [_P]),
    "damgpu_last_error": (C.c_char_p, []),
    "damgpu_launch_count": (C.c_uint64, []),
    "damgpu_device_memory": (C.c_int, [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "damgpu_time_kernels": (None, [C.c_int]),
    "damgpu_set_align_tier": (None, [C.c_int, C.c_int]),
    "damgpu_radix_totals": (None, [C.POINTER(C.c_double), C.c_int]),
    "damgpu_last_sort_times": (None, [C.POINTER(C.c_float)]),
    "damgpu_last_join_times": (None, [C.POINTER(C.c_float)]),
    "damgpu_Set_Filter_Params": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "damgpu_Sort_Kmers": (_P, [C.POINTER(CBlock), C.POINTER(C.c_int)]),
    "damgpu_block_upload": (_P, [C.POINTER(CBlock)]),
    "damgpu_block_upload_packed": (_P, [C.POINTER(CBlock), C.c_void_p, C.c_void_p, C.c_int64]),
    "damgpu_block_free": (None, [_P]),
    "damgpu_block_complement": (None, [_P]),
    "damgpu_block_download_bases": (None, [_P, _P]),
    "damgpu_index_build": (_P, [_P]),
    "damgpu_index_build_deferred": (_P, [_P]),
    "damgpu_index_is_deferred": (C.c_int, [_P]),
    "damgpu_set_reads_filter": (None, [C.c_int, C.c_int]),
    "damgpu_last_filter_times": (None, [C.POINTER(C.c_float)]),
    "damgpu_index_len": (C.c_int, [_P]),
    "damgpu_index_download": (None, [_P, _P]),
    "damgpu_index_free": (None, [_P]),
    "damgpu_index_device_ptr": (_P, [_P]),
    "damgpu_index_export": (None, [_P, _P]),
    "damgpu_index_import": (_P, [_P, C.c_int]),
    "damgpu_seeds_build": (_P, [_P, _P, _P, _P]),
    "damgpu_seeds_count": (C.c_int64, [_P]),
    "damgpu_seeds_limit": (C.c_int, [_P]),
    "damgpu_seeds_histogram": (None, [_P, _P]),
    "damgpu_seeds_download": (None, [_P, _P]),
    "damgpu_seeds_free": (None, [_P]),
    "damgpu_Match_Filter": (None, [C.POINTER(CBlock), C.POINTER(CBlock), _P, C.c_int, _P, C.c_int,
                                   C.c_int, C.c_int]),
    "damgpu_Reporter": (None, [C.c_char_p, C.POINTER(CBlock), C.c_char_p, C.POINTER(CBlock),
                               C.POINTER(CAlignSpec), C.c_int]),
    "damgpu_mapper_new": (_P, [_P, _P]),
    "damgpu_mapper_free": (None, [_P]),
    "damgpu_mapper_match": (None, [_P, _P, _P, C.c_int, C.c_int]),
    "damgpu_mapper_chain": (None, [_P, _P, C.c_int, C.c_int, C.c_int]),
    "damgpu_mapper_last_hits": (C.c_int64, [_P]),
    "damgpu_mapper_last_limit": (C.c_int, [_P]),
    "damgpu_mapper_num_candidates": (C.c_int64, [_P]),
    "damgpu_mapper_get_candidates": (C.c_int64, [_P, _P, _P, _P, C.c_int64]),
    "damgpu_mapper_get_cover": (C.c_int64, [_P, _P, C.c_int64]),
    "damgpu_mapper_report": (_P, [_P, _P, C.POINTER(CAlignSpec), C.c_int]),
    "damgpu_report_free": (None, [_P]),
    "damgpu_report_bytes": (C.c_int64, [_P, C.c_int]),
    "damgpu_report_records": (C.c_int64, [_P, C.c_int]),
    "damgpu_report_copy": (None, [_P, C.c_int, _P]),
    "damgpu_report_stats": (None, [_P, C.POINTER(C.c_int64)]),
    "damgpu_report_write_las": (C.c_int, [_P, C.c_int, C.c_char_p, C.c_char_p, C.c_char_p,
                                          C.c_int, C.c_int]),
    "damgpu_report_write_profile": (C.c_int, [_P, C.POINTER(CBlock), C.c_char_p, C.c_char_p,
                                              C.c_int]),
}

_lib = None


def load():
    """Load libdamgpu.so and bind every declared symbol.  Raises if the library is missing:
    the product path has no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libdamgpu.so is not built: run `python -c 'import __graft_entry__ "
                               "as g; g.build()'` (or make -C damapper_b200/csrc)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


_inited = None


def init(device: int = -1):
    """damgpu_init once per device (the library itself returns at once when it is set up already)."""
    global _inited
    L = load()
    if _inited is not None and (device < 0 or device == _inited):
        return L
    if L.damgpu_init(device) != 0:
        raise RuntimeError("libdamgpu: no usable CUDA device: %s" % L.damgpu_last_error().decode())
    _inited = device if device >= 0 else 0
    return L


class HostBlock:
    """Host image of a loaded DB block (what Load_All_Reads leaves in DAZZ_DB)."""

    def __init__(self, bases: np.ndarray, boff: np.ndarray, rlen: np.ndarray, tfirst: int = 0,
                 path_len: int = 16, mask=None):
        self.bases = np.ascontiguousarray(bases, dtype=np.uint8)     # leading 4 at index 0
        self.boff = np.ascontiguousarray(boff, dtype=np.int64)
        self.rlen = np.ascontiguousarray(rlen, dtype=np.int32)
        self.nreads = int(self.rlen.size)
        self.tfirst = tfirst
        self.maxlen = int(self.rlen.max()) if self.nreads else 0
        self.totlen = int(self.rlen.sum())
        self.sizeof_db = (112 + 40 * (self.nreads + 2) + path_len + 1
                          + (self.totlen + self.nreads + 4))       # sizeof_DB, DB.c:1044-1051
        self.mask_off = self.mask_pts = None
        if mask is not None:                                # (offsets in ints [nreads+1], points)
            self.mask_off = np.ascontiguousarray(mask[0], dtype=np.int64)
            self.mask_pts = np.ascontiguousarray(np.concatenate([mask[1], [0]]), dtype=np.int32)
        self.c = CBlock(self.bases.ctypes.data + 1, self.boff.ctypes.data, self.rlen.ctypes.data,
                        self.nreads, tfirst, self.maxlen, self.totlen, self.sizeof_db,
                        self.mask_off.ctypes.data if mask is not None else None,
                        self.mask_pts.ctypes.data if mask is not None else None, None, None, 0)

    def attach_packed(self, packed: np.ndarray, poff: np.ndarray):
        """Give the block its .bps image (pack_bps): layer 1 then uploads 2 bits per base."""
        self.packed = packed if packed.dtype == np.uint8 and packed.flags.c_contiguous else np.ascontiguousarray(packed, dtype=np.uint8)
        self.poff = np.ascontiguousarray(poff, dtype=np.int64)
        self.c.packed = self.packed.ctypes.data
        self.c.poff = self.poff.ctypes.data
        self.c.packed_bytes = int(self.packed.size)
        return self


def set_filter_params(kmer: int = 20, suppress: int = 0, nthreads: int = 4) -> int:
    return load().damgpu_Set_Filter_Params(kmer, suppress, nthreads)


_opts_keep = None


def set_options(verbose=0, profile=0, spacing=100, best_tie=1.0, sort_path="/tmp",
                mem_limit=64 << 30, mem_physical=64 << 30):
    global _opts_keep
    o = COptions(verbose, profile, spacing, best_tie, sort_path.encode(), mem_limit, mem_physical)
    _opts_keep = o
    load().damgpu_set_options(C.byref(o))


def pack_bps(hb: "HostBlock"):
    """The block as the .bps file holds it: (packed bytes, per-read byte offsets); four bases per
    byte, first base in the two top bits (reference DB.c:319-340)."""
    chunks, poff, o = [], np.zeros(hb.nreads, dtype=np.int64), 0
    for i in range(hb.nreads):
        s = hb.bases[1 + int(hb.boff[i]): 1 + int(hb.boff[i]) + int(hb.rlen[i])]
        pad = (-s.size) % 4
        q = np.concatenate([s, np.zeros(pad, dtype=np.uint8)]).reshape(-1, 4)
        chunks.append(((q[:, 0] << 6) | (q[:, 1] << 4) | (q[:, 2] << 2) | q[:, 3]).astype(np.uint8))
        poff[i] = o
        o += chunks[-1].size
    packed = np.concatenate(chunks) if chunks else np.zeros(0, dtype=np.uint8)
    return np.ascontiguousarray(packed), poff


class DeviceBlock:
    def __init__(self, hb: HostBlock, packed: bool = False):
        self.host = hb
        if packed:
            pk, poff = pack_bps(hb)
            self.h = init().damgpu_block_upload_packed(C.byref(hb.c), pk.ctypes.data, poff.ctypes.data, pk.size)
        else:
            self.h = init().damgpu_block_upload(C.byref(hb.c))

    @classmethod
    def from_host(cls, hb: HostBlock):
        """Upload through the block view itself: 2 bits per base when the block carries its .bps image
        (HostBlock.attach_packed), one byte per base otherwise."""
        self = cls.__new__(cls)
        self.host = hb
        if getattr(hb, "packed", None) is not None:
            self.h = init().damgpu_block_upload_packed(C.byref(hb.c), hb.packed.ctypes.data, hb.poff.ctypes.data,
                                                       int(hb.packed.size))
        else:
            self.h = init().damgpu_block_upload(C.byref(hb.c))
        return self

    def complement(self):
        load().damgpu_block_complement(self.h)

    def download_bases(self) -> np.ndarray:
        out = np.zeros_like(self.host.bases)
        load().damgpu_block_download_bases(self.h, out.ctypes.data + 1)
        return out

    def free(self):
        if self.h:
            load().damgpu_block_free(self.h)
            self.h = None

    def __del__(self):
        self.free()


class Index:
    """Sort_Kmers of a resident block.  deferred=True is the reads-side form: the list is built
    (filtered by the reference block's codes) by the first Match_Filter that uses it; the block must
    stay alive as long as the index."""

    def __init__(self, blk: DeviceBlock = None, handle=None, deferred: bool = False):
        if handle is not None:
            self.h = handle
        elif deferred:
            self.h = load().damgpu_index_build_deferred(blk.h)
            self._blk = blk
        else:
            self.h = load().damgpu_index_build(blk.h)

    @property
    def is_deferred(self) -> bool:
        return bool(load().damgpu_index_is_deferred(self.h))

    def __len__(self):
        return load().damgpu_index_len(self.h)

    def download(self) -> np.ndarray:
        n = len(self)
        if n == 0:
            return np.zeros(0, dtype=KMER_DT)
        out = np.zeros(n + 2, dtype=KMER_DT)
        load().damgpu_index_download(self.h, out.ctypes.data)
        return out

    def free(self):
        if self.h:
            load().damgpu_index_free(self.h)
            self.h = None

    def __del__(self):
        self.free()


class Seeds:
    def __init__(self, ridx: Index, reads: DeviceBlock, gidx: Index, ref: DeviceBlock):
        self.h = load().damgpu_seeds_build(ridx.h, reads.h, gidx.h, ref.h)

    @property
    def count(self):
        return load().damgpu_seeds_count(self.h)

    @property
    def limit(self):
        return load().damgpu_seeds_limit(self.h)

    def histogram(self):
        out = np.zeros(10000, dtype=np.int64)
        load().damgpu_seeds_histogram(self.h, out.ctypes.data)
        return out

    def download(self) -> np.ndarray:
        n = self.count
        out = np.zeros(n + 1, dtype=SEED_DT)
        if n > 0:
            load().damgpu_seeds_download(self.h, out.ctypes.data)
        return out

    def free(self):
        if self.h:
            load().damgpu_seeds_free(self.h)
            self.h = None

    def __del__(self):
        self.free()


def last_sort_times():
    v = (C.c_float * 3)()
    load().damgpu_last_sort_times(v)
    return dict(extract_ms=v[0], sort_ms=v[1], npass=int(v[2]))


def last_join_times():
    v = (C.c_float * 4)()
    load().damgpu_last_join_times(v)
    return dict(lut_ms=v[0], match_ms=v[1], alen=int(v[2]), blen=int(v[3]))


def last_filter_times():
    """Of the last filtered build of a deferred reads index (kernel timing on): ms of the reference
    bitmap, of the filtered extraction, of the radix passes over the survivors; survivors."""
    v = (C.c_float * 4)()
    load().damgpu_last_filter_times(v)
    return dict(bitmap_ms=v[0], extract_ms=v[1], sort_ms=v[2], survivors=int(v[3]))


def set_reads_filter(mode: str = "auto", log2_bits: int = 0):
    """off: a deferred reads index is simply sorted; auto: filtered when the block is large;
    always: filtered whatever the size (parity tests)."""
    load().damgpu_set_reads_filter({"off": 0, "auto": 1, "always": 2}[mode], log2_bits)


class Report:
    def __init__(self, handle):
        self.h = handle

    def _bytes(self, fam):
        n = load().damgpu_report_bytes(self.h, fam)
        buf = np.zeros(max(n, 1), dtype=np.uint8)
        load().damgpu_report_copy(self.h, fam, buf.ctypes.data)
        return buf[:n].tobytes()

    @property
    def a(self):
        return self._bytes(0)

    @property
    def b(self):
        return self._bytes(1)

    @property
    def prof(self):
        return self._bytes(2)

    def records(self, fam=0):
        return load().damgpu_report_records(self.h, fam)

    def stats(self):
        v = (C.c_int64 * 8)()
        load().damgpu_report_stats(self.h, v)
        return dict(nalign=v[0], nwaves=v[1], ncells=v[2], h2=v[3], overflow_jobs=v[4],
                    empty_band=v[5], align_ms=v[6] / 1000.0, trace_fails=v[7])

    def write_las(self, fam, directory, aname, bname, nfiles, tspace):
        rc = load().damgpu_report_write_las(self.h, fam, directory.encode(), aname.encode(),
                                            bname.encode(), nfiles, tspace)
        if rc != 0:
            raise IOError("cannot write .las files in %s" % directory)

    def free(self):
        if self.h:
            load().damgpu_report_free(self.h)
            self.h = None

    def __del__(self):
        self.free()


class Mapper:
    """Per-reads-block state between Match_Filter calls and Reporter (map.c:2885)."""

    def __init__(self, reads: DeviceBlock, ridx: Index):
        self.reads, self.ridx = reads, ridx
        self.h = load().damgpu_mapper_new(reads.h, ridx.h)

    def match(self, ref: DeviceBlock, gidx: Index, comp: int, start: int):
        load().damgpu_mapper_match(self.h, ref.h, gidx.h, comp, start)

    def chain(self, seeds: Seeds, bstart: int, comp: int, start: int):
        load().damgpu_mapper_chain(self.h, seeds.h, bstart, comp, start)

    @property
    def last_hits(self):
        return load().damgpu_mapper_last_hits(self.h)

    def candidates(self):
        L = load()
        n = L.damgpu_mapper_num_candidates(self.h)
        out = np.zeros(n, dtype=CAND_DT)
        jcnt = np.zeros(n, dtype=np.int32)
        nj = L.damgpu_mapper_get_candidates(self.h, out.ctypes.data, jcnt.ctypes.data, None, 0)
        jumps = np.zeros((max(nj, 1), 2), dtype=np.int32)
        L.damgpu_mapper_get_candidates(self.h, out.ctypes.data, jcnt.ctypes.data, jumps.ctypes.data, nj)
        return out, jcnt, jumps[:nj]

    def cover(self):
        n = load().damgpu_mapper_get_cover(self.h, None, 0)
        out = np.zeros(n, dtype=np.int16)
        load().damgpu_mapper_get_cover(self.h, out.ctypes.data, n)
        return out

    def report(self, wholeref: DeviceBlock, ave_corr=0.85, spacing=100,
               freq=(.25, .25, .25, .25), mflag=1) -> Report:
        spec = CAlignSpec(ave_corr, spacing, (C.c_float * 4)(*freq))
        return Report(load().damgpu_mapper_report(self.h, wholeref.h, C.byref(spec), mflag))

    def free(self):
        if self.h:
            load().damgpu_mapper_free(self.h)
            self.h = None

    def __del__(self):
        self.free()


def map_block(reads: HostBlock, ref_blocks, wholeref: HostBlock, kmer=20, suppress=0, spacing=100,
              profile=0, ave_corr=0.85, best_tie=1.0, freq=(.25, .25, .25, .25),
              mem_limit=64 << 30, do_a=1, do_b=0, nthreads=4, want_candidates=False,
              reads_filter="auto"):
    """The damapper flow for one reads block (damapper.c:825-879) on the GPU: index the reads,
    then for every reference block Match_Filter forward and complemented (the block is
    complemented on the device), then Reporter against the whole reference.
    `ref_blocks` is a list of forward HostBlocks.  Returns a dict of record streams, the -p track and counters."""
    init()
    set_filter_params(kmer, suppress, nthreads)
    set_options(profile=profile, spacing=spacing, best_tie=best_tie, mem_limit=mem_limit)
    set_reads_filter(reads_filter)
    dr = DeviceBlock(reads)
    ir = Index(dr, deferred=True)         # the reads list is built per reference block (kmer_filter.cu)
    m = Mapper(dr, ir)
    for k, fwd in enumerate(ref_blocks):
        dg = DeviceBlock(fwd)
        ig = Index(dg)
        m.match(dg, ig, 0, 1 if k == 0 else 0)
        ig.free()
        dg.complement()
        ig = Index(dg)
        m.match(dg, ig, 1, 0)
        limit = load().damgpu_mapper_last_limit(m.h)
        deferred = ir.is_deferred
        ig.free()
        dg.free()
    cands = m.candidates() if want_candidates else None
    cover = m.cover() if (want_candidates and profile) else None
    dw = DeviceBlock(wholeref)
    rep = m.report(dw, ave_corr, spacing, freq, (1 if do_a else 0) | (2 if do_b else 0))
    out = dict(a=rep.a, anrec=rep.records(0), b=rep.b, bnrec=rep.records(1), prof=rep.prof,
               stats=rep.stats(), candidates=cands, cover=cover, limit=limit, deferred=deferred)
    rep.free(); dw.free(); m.free(); ir.free(); dr.free()
    set_reads_filter("auto")
    return out
