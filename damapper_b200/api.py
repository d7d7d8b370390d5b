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
LIB_PATH = os.path.join(HERE, "libdamgpu.so")

KMER_DT = np.dtype([("code", "<u8"), ("rpos", "<i4"), ("read", "<i4")])
SEED_DT = np.dtype([("diag", "<i4"), ("apos", "<i4"), ("bread", "<i4"), ("aread", "<i4")])
CAND_DT = np.dtype([(n, "<i4") for n in ("read", "score", "length", "bread", "comp", "afirst",
                                         "alast", "bfirst", "blast")])


class CBlock(C.Structure):
    _fields_ = [("bases", C.c_void_p), ("boff", C.c_void_p), ("rlen", C.c_void_p),
                ("nreads", C.c_int32), ("tfirst", C.c_int32), ("maxlen", C.c_int32),
                ("totlen", C.c_int64), ("sizeof_db", C.c_int64)]


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
    "damgpu_time_kernels": (None, [C.c_int]),
    "damgpu_last_sort_times": (None, [C.POINTER(C.c_float)]),
    "damgpu_Set_Filter_Params": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "damgpu_Sort_Kmers": (_P, [C.POINTER(CBlock), C.POINTER(C.c_int)]),
    "damgpu_block_upload": (_P, [C.POINTER(CBlock)]),
    "damgpu_block_free": (None, [_P]),
    "damgpu_block_complement": (None, [_P]),
    "damgpu_block_download_bases": (None, [_P, _P]),
    "damgpu_index_build": (_P, [_P]),
    "damgpu_index_len": (C.c_int, [_P]),
    "damgpu_index_download": (None, [_P, _P]),
    "damgpu_index_free": (None, [_P]),
    "damgpu_index_device_ptr": (_P, [_P]),
    "damgpu_index_adopt": (_P, [_P, C.c_int]),
    "damgpu_seeds_build": (_P, [_P, _P, _P, _P]),
    "damgpu_seeds_count": (C.c_int64, [_P]),
    "damgpu_seeds_limit": (C.c_int, [_P]),
    "damgpu_seeds_histogram": (None, [_P, _P]),
    "damgpu_seeds_download": (None, [_P, _P]),
    "damgpu_seeds_free": (None, [_P]),
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


def init(device: int = -1):
    L = load()
    if L.damgpu_init(device) != 0:
        raise RuntimeError("libdamgpu: no usable CUDA device: %s" % L.damgpu_last_error().decode())
    return L


class HostBlock:
    """Host image of a loaded DB block (what Load_All_Reads leaves in DAZZ_DB)."""

    def __init__(self, bases: np.ndarray, boff: np.ndarray, rlen: np.ndarray, tfirst: int = 0,
                 path_len: int = 16):
        self.bases = np.ascontiguousarray(bases, dtype=np.uint8)     # leading 4 at index 0
        self.boff = np.ascontiguousarray(boff, dtype=np.int64)
        self.rlen = np.ascontiguousarray(rlen, dtype=np.int32)
        self.nreads = int(self.rlen.size)
        self.tfirst = tfirst
        self.maxlen = int(self.rlen.max()) if self.nreads else 0
        self.totlen = int(self.rlen.sum())
        self.sizeof_db = (112 + 40 * (self.nreads + 2) + path_len + 1
                          + (self.totlen + self.nreads + 4))       # sizeof_DB, DB.c:1044-1051
        self.c = CBlock(self.bases.ctypes.data + 1, self.boff.ctypes.data, self.rlen.ctypes.data,
                        self.nreads, tfirst, self.maxlen, self.totlen, self.sizeof_db)


def set_filter_params(kmer: int = 20, suppress: int = 0, nthreads: int = 4) -> int:
    return load().damgpu_Set_Filter_Params(kmer, suppress, nthreads)


_opts_keep = None


def set_options(verbose=0, profile=0, spacing=100, best_tie=1.0, sort_path="/tmp",
                mem_limit=64 << 30, mem_physical=64 << 30):
    global _opts_keep
    o = COptions(verbose, profile, spacing, best_tie, sort_path.encode(), mem_limit, mem_physical)
    _opts_keep = o
    load().damgpu_set_options(C.byref(o))


class DeviceBlock:
    def __init__(self, hb: HostBlock):
        self.host = hb
        self.h = init().damgpu_block_upload(C.byref(hb.c))

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
    def __init__(self, blk: DeviceBlock = None, handle=None):
        self.h = handle if handle is not None else load().damgpu_index_build(blk.h)

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
