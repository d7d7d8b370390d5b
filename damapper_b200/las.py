""".las reader/writer helpers and the canonical record stream used for parity.

Record layout (reference align.c:3098-3122, align.h:89-95,336-341):
  file header  int64 novl, int32 tspace
  record       int32 tlen, diffs, abpos, bbpos, aepos, bepos; uint32 flags; int32 aread, bread;
               4 padding bytes (uninitialised in the reference, hazard H1) ; then tlen trace
               values, uint8 when tspace <= 125 (TRACE_XOVR, align.h:21) else uint16.
"""
from __future__ import annotations

import struct

import numpy as np

REC = struct.Struct("<iiiiiiIii4x")
assert REC.size == 40
TRACE_XOVR = 125


def read_las(path: str):
    """Return (tspace, [record dict]) with the trace as a numpy array."""
    with open(path, "rb") as f:
        data = f.read()
    return parse_las(data)


def parse_las(data: bytes):
    novl, tspace = struct.unpack_from("<qi", data, 0)
    off = 12
    tb = 1 if tspace <= TRACE_XOVR else 2
    recs = []
    for _ in range(novl):
        tlen, diffs, abpos, bbpos, aepos, bepos, flags, aread, bread = REC.unpack_from(data, off)
        off += 40
        tr = np.frombuffer(data, dtype=np.uint8 if tb == 1 else np.uint16, count=tlen, offset=off)
        off += tlen * tb
        recs.append(dict(tlen=tlen, diffs=diffs, abpos=abpos, bbpos=bbpos, aepos=aepos,
                         bepos=bepos, flags=flags, aread=aread, bread=bread, trace=tr))
    assert off == len(data), "trailing bytes in .las"
    return tspace, recs


def canonical_stream(paths) -> bytes:
    """Concatenate the record streams of per-thread files (in the given order) into one
    canonical byte string: padding bytes 36-39 zeroed (H1), file headers dropped.
    The stream is invariant to the -T thread count."""
    out = bytearray()
    tspace_seen = None
    for p in paths:
        with open(p, "rb") as f:
            data = f.read()
        novl, tspace = struct.unpack_from("<qi", data, 0)
        tspace_seen = tspace if tspace_seen is None else tspace_seen
        assert tspace == tspace_seen
        tb = 1 if tspace <= TRACE_XOVR else 2
        off = 12
        for _ in range(novl):
            tlen = struct.unpack_from("<i", data, off)[0]
            out += data[off:off + 36] + b"\0\0\0\0"
            off += 40
            out += data[off:off + tlen * tb]
            off += tlen * tb
        assert off == len(data)
    return bytes(out)


def stream_records(stream: bytes, tspace: int):
    """Parse a canonical stream (no file header) back into record dicts."""
    tb = 1 if tspace <= TRACE_XOVR else 2
    off = 0
    recs = []
    while off < len(stream):
        tlen, diffs, abpos, bbpos, aepos, bepos, flags, aread, bread = REC.unpack_from(stream, off)
        off += 40
        tr = np.frombuffer(stream, dtype=np.uint8 if tb == 1 else np.uint16, count=tlen, offset=off)
        off += tlen * tb
        recs.append(dict(tlen=tlen, diffs=diffs, abpos=abpos, bbpos=bbpos, aepos=aepos,
                         bepos=bepos, flags=flags, aread=aread, bread=bread, trace=tr))
    return recs


def check_trace_points(recs, tspace: int) -> int:
    """The reference's only invariant checker, Check_Trace_Points (align.c:3194-3236):
    ((aepos-1)/ts - abpos/ts)*2 == tlen-2 and bbpos + sum(b-advances) == bepos.
    Returns the number of records violating it."""
    bad = 0
    for r in recs:
        if r["tlen"] == 0:
            continue
        if ((r["aepos"] - 1) // tspace - r["abpos"] // tspace) * 2 != r["tlen"] - 2:
            bad += 1
            continue
        if r["bbpos"] + int(np.asarray(r["trace"][1::2], dtype=np.int64).sum()) != r["bepos"]:
            bad += 1
    return bad
