"""Minimal DAZZ_DB writer/reader (stub + .idx + .bps) for synthetic workloads.

The DAZZ_DB tools (fasta2DB, fasta2DAM, DBsplit) are not part of the reference tree,
so databases are written directly in the on-disk layout the reference reads:
  stub      reference DB.h:431-435 (DB_NFILE/DB_FDATA/DB_NBLOCK/DB_PARAMS/DB_BDATA)
  .<root>.idx   raw DAZZ_DB (112 B) + DAZZ_READ (40 B) records, reference DB.h:285-295,390-420
  .<root>.bps   2 bits/base, first base in the top bits, reference DB.c:319-363
and the in-memory block form of Load_All_Reads (reference DB.c:1389-1441).
"""
from __future__ import annotations

import os
import struct

import numpy as np

DB_HDR = struct.Struct("<iiii4fi4xqiiiii4xqi4xqqq")   # 112 bytes
DB_READ = struct.Struct("<iii4xqqi4x")                # 40 bytes
DB_BEST = 0x800
DB_ALL = 0x1
assert DB_HDR.size == 112 and DB_READ.size == 40


def pack_bps(seq: np.ndarray) -> bytes:
    """2-bit pack one read; (len+3)>>2 bytes, first base in the two most significant bits."""
    n = seq.size
    pad = (-n) % 4
    s = np.concatenate([seq.astype(np.uint8), np.zeros(pad, dtype=np.uint8)]).reshape(-1, 4)
    return ((s[:, 0] << 6) | (s[:, 1] << 4) | (s[:, 2] << 2) | s[:, 3]).astype(np.uint8).tobytes()


def write_db(path: str, seqs, is_dam: bool = False, nblocks: int = 1, block_bounds=None) -> str:
    """Write a DB/DAM holding `seqs` (list of uint8 arrays 0..3, or (bases, rlen) tuple).

    `nblocks` splits the reads evenly by count into blocks (or pass explicit
    `block_bounds` = list of first-read indices, len nblocks+1).  cutoff=0/all=1 so
    Trim_DB is a no-op (reference DB.c:918).  Returns the stub path."""
    if isinstance(seqs, tuple):
        bases, rlen = seqs
        rlen = np.asarray(rlen, dtype=np.int64)
        offs = np.concatenate([[0], np.cumsum(rlen)])
        get = lambda i: bases[offs[i]:offs[i + 1]]
        n = rlen.size
    else:
        n = len(seqs)
        rlen = np.array([s.size for s in seqs], dtype=np.int64)
        get = lambda i: seqs[i]
    d, base = os.path.split(path)
    d = d or "."
    root = base
    for ext in (".dam", ".db"):
        if root.endswith(ext):
            root = root[: -len(ext)]
    ext = ".dam" if is_dam else ".db"
    stub = os.path.join(d, root + ext)
    if block_bounds is None:
        block_bounds = [int(n * i / nblocks) for i in range(nblocks + 1)]
    nblocks = len(block_bounds) - 1

    # .bps (vectorised when all reads are the packed tuple form)
    bps_path = os.path.join(d, "." + root + ".bps")
    boff = np.zeros(n, dtype=np.int64)
    with open(bps_path, "wb") as f:
        o = 0
        for i in range(n):
            b = pack_bps(get(i))
            boff[i] = o
            f.write(b)
            o += len(b)

    allb = np.concatenate([get(i) for i in range(n)]) if n else np.zeros(0, np.uint8)
    cnt = np.bincount(allb, minlength=4).astype(np.float64)
    freq = cnt / max(1.0, cnt.sum())
    totlen = int(rlen.sum())
    maxlen = int(rlen.max()) if n else 0
    with open(os.path.join(d, "." + root + ".idx"), "wb") as f:
        f.write(DB_HDR.pack(n, n, 0, DB_ALL, float(freq[0]), float(freq[1]), float(freq[2]),
                            float(freq[3]), maxlen, totlen, n, 0, 0, 0, 0, 0, 0, 0, 0, 0))
        for i in range(n):
            f.write(DB_READ.pack(i if is_dam else i, int(rlen[i]), 0, int(boff[i]), 0, DB_BEST))

    with open(stub, "w") as f:
        f.write("files = %9d\n" % 1)
        f.write("  %9d %s %s\n" % (n, root, root))
        f.write("blocks = %9d\n" % nblocks)
        f.write("size = %11d cutoff = %9d all = %1d\n" % (max(1, totlen // 1000000 + 1), 0, 1))
        for b in block_bounds:
            f.write(" %9d %9d\n" % (b, b))
    return stub


def load_block(seqs) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """In-memory block image as produced by Load_All_Reads (reference DB.c:1389-1441).

    Returns (bases, boff, rlen): `bases` is one byte per base with a 4 before the first
    read and after every read (bases[0] is the leading 4; read i starts at bases[1+boff[i]]),
    `boff` has nreads+1 entries."""
    if isinstance(seqs, tuple):
        b, rlen = seqs
        rlen = np.asarray(rlen, dtype=np.int64)
        n = rlen.size
        boff = np.concatenate([[0], np.cumsum(rlen + 1)]).astype(np.int64)
        out = np.full(int(boff[-1]) + 4, 4, dtype=np.uint8)
        src_off = np.concatenate([[0], np.cumsum(rlen)])
        idx = np.arange(int(rlen.sum()), dtype=np.int64)
        rid = np.repeat(np.arange(n, dtype=np.int64), rlen)
        out[1 + boff[rid] + (idx - src_off[rid])] = b
        return out, boff, rlen.astype(np.int32)
    n = len(seqs)
    rlen = np.array([s.size for s in seqs], dtype=np.int64)
    boff = np.concatenate([[0], np.cumsum(rlen + 1)]).astype(np.int64)
    out = np.full(int(boff[-1]) + 4, 4, dtype=np.uint8)
    for i, s in enumerate(seqs):
        out[1 + boff[i]: 1 + boff[i] + s.size] = s
    return out, boff, rlen.astype(np.int32)


def revcomp_contigs(contigs):
    """complement_DB (reference damapper.c:433-525): reverse-complement every contig in place."""
    return [(3 - c[::-1]).astype(np.uint8) for c in contigs]


def random_masks(rlen, seed: int = 1, max_intervals: int = 3, max_len: int = 400):
    """A merged mask track for reads of lengths `rlen`: (offsets [n+1] counting ints, points) with
    sorted, disjoint [begin, end) intervals per read, some touching the read ends (the form
    block->tracks has after damapper.c:381-399)."""
    rng = np.random.default_rng(seed)
    off, pts = [0], []
    for n in np.asarray(rlen, dtype=np.int64):
        k = int(rng.integers(0, max_intervals + 1))
        cuts = np.sort(rng.choice(np.arange(0, int(n) + 1), size=min(2 * k, int(n) + 1), replace=False))
        iv = []
        for j in range(0, cuts.size - 1, 2):
            b, e = int(cuts[j]), int(min(cuts[j + 1], cuts[j] + max_len))
            if e > b:
                iv += [b, e]
        if k and rng.random() < 0.2 and iv:
            iv[0] = 0                                   # starts at the first base
        if k and rng.random() < 0.2 and iv:
            iv[-1] = int(n)                             # ends at the last base
        pts += iv
        off.append(len(pts))
    return np.asarray(off, dtype=np.int64), np.asarray(pts, dtype=np.int32)


def mirror_masks(off, pts, rlen):
    """Mask track of the complemented block (complement_DB, reference damapper.c:471-522): per read
    the point list is reversed and every point x becomes rlen - x."""
    out = np.array(pts, dtype=np.int32, copy=True)
    for i, n in enumerate(np.asarray(rlen, dtype=np.int64)):
        b, e = int(off[i]), int(off[i + 1])
        out[b:e] = (int(n) - np.asarray(pts[b:e], dtype=np.int64))[::-1]
    return np.asarray(off, dtype=np.int64), out


def write_mask_track(stub: str, name: str, off, pts) -> None:
    """Write mask track `name` of the DB whose stub is `stub` (whole-DB track over the untrimmed
    reads): .<root>.<name>.anno = int32 nreads, int32 0, int64 byte offsets [nreads+1];
    .<root>.<name>.data = int32 interval end points (reference DB.c:1649-1702,1804-1990)."""
    d, base = os.path.split(stub)
    root = base
    for ext in (".dam", ".db"):
        if root.endswith(ext):
            root = root[: -len(ext)]
    off = np.asarray(off, dtype=np.int64)
    with open(os.path.join(d or ".", ".%s.%s.anno" % (root, name)), "wb") as f:
        f.write(np.array([off.size - 1, 0], dtype=np.int32).tobytes())
        f.write((off * 4).astype(np.int64).tobytes())
    with open(os.path.join(d or ".", ".%s.%s.data" % (root, name)), "wb") as f:
        f.write(np.asarray(pts, dtype=np.int32).tobytes())


def union_masks(a, b):
    """Union of two mask tracks (offsets, points) read by read, as sorted disjoint intervals
    (what merge_tracks, reference damapper.c:253-343, yields up to zero-length gaps)."""
    off, pts = [0], []
    for i in range(len(a[0]) - 1):
        iv = [tuple(x) for x in np.asarray(a[1][a[0][i]:a[0][i + 1]]).reshape(-1, 2)]
        iv += [tuple(x) for x in np.asarray(b[1][b[0][i]:b[0][i + 1]]).reshape(-1, 2)]
        iv.sort()
        cur = None
        for s0, e0 in iv:
            if cur is not None and s0 <= cur[1]:
                cur[1] = max(cur[1], e0)
            else:
                if cur is not None:
                    pts += cur
                cur = [int(s0), int(e0)]
        if cur is not None:
            pts += cur
        off.append(len(pts))
    return np.asarray(off, dtype=np.int64), np.asarray(pts, dtype=np.int32)


class StreamDBWriter:
    """write_db for data that does not fit in memory at once: reads are appended in chunks
    ((bases, rlen) tuples), the stub / .idx header are written by close().  Same files as
    write_db (reference DB.h:285-295,390-435, DB.c:319-363)."""

    def __init__(self, path: str, is_dam: bool = False):
        d, base = os.path.split(path)
        self.d = d or "."
        root = base
        for ext in (".dam", ".db"):
            if root.endswith(ext):
                root = root[: -len(ext)]
        self.root = root
        self.is_dam = is_dam
        self.stub = os.path.join(self.d, root + (".dam" if is_dam else ".db"))
        self.bps = open(os.path.join(self.d, "." + root + ".bps"), "wb")
        self.idx_path = os.path.join(self.d, "." + root + ".idx")
        self.idx = open(self.idx_path, "wb")
        self.idx.write(b"\0" * DB_HDR.size)
        self.n = 0
        self.boff = 0
        self.totlen = 0
        self.maxlen = 0
        self.cnt = np.zeros(4, dtype=np.float64)

    def append(self, bases: np.ndarray, rlen) -> None:
        rlen = np.asarray(rlen, dtype=np.int64)
        n = rlen.size
        if n == 0:
            return
        # 2-bit pack all reads of the chunk at once: every read padded to a multiple of 4 bases
        plen = (rlen + 3) & ~3
        poff = np.concatenate([[0], np.cumsum(plen)])
        soff = np.concatenate([[0], np.cumsum(rlen)])
        padded = np.zeros(int(poff[-1]), dtype=np.uint8)
        rid = np.repeat(np.arange(n, dtype=np.int64), rlen)
        padded[np.arange(int(soff[-1]), dtype=np.int64) + (poff[:-1] - soff[:-1])[rid]] = bases
        del rid
        q = padded.reshape(-1, 4)
        packed = ((q[:, 0] << 6) | (q[:, 1] << 4) | (q[:, 2] << 2) | q[:, 3]).astype(np.uint8)
        self.bps.write(packed.tobytes())
        rec = np.zeros(n, dtype=np.dtype([("origin", "<i4"), ("rlen", "<i4"), ("fpulse", "<i4"), ("p0", "<i4"),
                                          ("boff", "<i8"), ("coff", "<i8"), ("flags", "<i4"), ("p1", "<i4")]))
        assert rec.dtype.itemsize == DB_READ.size
        rec["origin"] = np.arange(self.n, self.n + n)
        rec["rlen"] = rlen
        rec["boff"] = self.boff + poff[:-1] // 4
        rec["flags"] = DB_BEST
        self.idx.write(rec.tobytes())
        self.n += n
        self.boff += int(poff[-1]) // 4
        self.totlen += int(soff[-1])
        self.maxlen = max(self.maxlen, int(rlen.max()))
        self.cnt += np.bincount(bases, minlength=4)[:4]

    def close(self, nblocks: int = 1, block_bounds=None) -> str:
        n = self.n
        self.bps.close()
        freq = self.cnt / max(1.0, self.cnt.sum())
        self.idx.seek(0)
        self.idx.write(DB_HDR.pack(n, n, 0, DB_ALL, float(freq[0]), float(freq[1]), float(freq[2]),
                                   float(freq[3]), self.maxlen, self.totlen, n, 0, 0, 0, 0, 0, 0, 0, 0, 0))
        self.idx.close()
        if block_bounds is None:
            block_bounds = [int(n * i / nblocks) for i in range(nblocks + 1)]
        with open(self.stub, "w") as f:
            f.write("files = %9d\n" % 1)
            f.write("  %9d %s %s\n" % (n, self.root, self.root))
            f.write("blocks = %9d\n" % (len(block_bounds) - 1))
            f.write("size = %11d cutoff = %9d all = %1d\n" % (max(1, self.totlen // 1000000 + 1), 0, 1))
            for b in block_bounds:
                f.write(" %9d %9d\n" % (b, b))
        return self.stub
