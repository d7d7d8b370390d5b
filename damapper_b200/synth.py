"""Deterministic synthetic workloads for the DAMAPPER mapping core.

Generates reference genomes and PacBio-like reads of the shapes named in
BASELINE.json: i.i.d. uniform bases, reads sampled
uniformly over genome and strand, 15 % total error split ins:del:sub = 40:30:30.

Everything is numpy, vectorised over all reads at once, and seeded, so the same
call produces the same bytes here and on the GPU box.  Bases are numeric 0..3
(A,C,G,T) exactly as the DAZZ_DB in-memory form (reference DB.c:1389-1441).
"""
from __future__ import annotations

import numpy as np

__all__ = ["make_genome", "make_repeat_genome", "make_reads", "make_chimeric_reads",
           "CONFIGS", "make_config"]


def make_genome(length: int, seed: int = 1) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return rng.integers(0, 4, size=length, dtype=np.uint8)


def _mutate(seq: np.ndarray, rate: float, rng) -> np.ndarray:
    """Point-substitute a fraction `rate` of positions (used for diverged repeat copies)."""
    out = seq.copy()
    hit = rng.random(seq.size) < rate
    out[hit] = (out[hit] + rng.integers(1, 4, size=int(hit.sum()), dtype=np.uint8)) & 3
    return out


def make_repeat_genome(length: int, seed: int = 1, repeat_frac: float = 0.20,
                       nfamilies: int = 8, fam_len=(300, 3000), divergence: float = 0.03,
                       long_dups: int = 2, long_dup_len: int = 30000,
                       long_dup_div: float = 0.002) -> np.ndarray:
    """Repeat-rich genome (config 3): `repeat_frac` of the bases are copies of a small
    library of repeat families (copies diverged by `divergence`), plus `long_dups`
    near-identical read-length-or-longer duplications that exercise the -n filter."""
    rng = np.random.default_rng(seed)
    g = rng.integers(0, 4, size=length, dtype=np.uint8)
    fams = [rng.integers(0, 4, size=int(rng.integers(fam_len[0], fam_len[1] + 1)), dtype=np.uint8)
            for _ in range(nfamilies)]
    target = int(length * repeat_frac)
    placed = 0
    while placed < target:
        f = fams[int(rng.integers(0, nfamilies))]
        if f.size >= length:
            break
        p = int(rng.integers(0, length - f.size))
        c = _mutate(f, divergence, rng)
        if rng.random() < 0.5:
            c = (3 - c[::-1]).astype(np.uint8)
        g[p:p + f.size] = c
        placed += f.size
    for _ in range(long_dups):
        if long_dup_len * 3 >= length:
            break
        s = int(rng.integers(0, length - long_dup_len))
        d = int(rng.integers(0, length - long_dup_len))
        g[d:d + long_dup_len] = _mutate(g[s:s + long_dup_len].copy(), long_dup_div, rng)
    return g


def _apply_errors(src: np.ndarray, seg_of: np.ndarray, nseg: int, err, rng):
    """Apply indel/substitution noise to the concatenated source bases `src`.

    `seg_of[i]` is the read index of source base i; `err` is a per-base total error
    rate array (or scalar).  Returns (bases, read_lengths)."""
    n = src.size
    err = np.broadcast_to(np.asarray(err, dtype=np.float64), (n,))
    u = rng.random(n)
    p_del = 0.30 * err
    p_sub = 0.30 * err
    p_ins = 0.40 * err
    dele = u < p_del
    sub = (u >= p_del) & (u < p_del + p_sub)
    base = src.copy()
    base[sub] = (base[sub] + rng.integers(1, 4, size=int(sub.sum()), dtype=np.uint8)) & 3
    ins = rng.random(n) < p_ins
    keep = ~dele
    out_cnt = keep.astype(np.int64) + ins.astype(np.int64)
    off = np.cumsum(out_cnt) - out_cnt
    total = int(out_cnt.sum())
    out = np.empty(total, dtype=np.uint8)
    out[off[keep]] = base[keep]
    ins_pos = off[ins] + keep[ins].astype(np.int64)
    out[ins_pos] = rng.integers(0, 4, size=int(ins.sum()), dtype=np.uint8)
    rlen = np.bincount(seg_of, weights=out_cnt, minlength=nseg).astype(np.int64)
    return out, rlen


def make_reads(genome: np.ndarray, nreads: int, src_len: int = 10000, error: float = 0.15,
               seed: int = 2, lognormal_sigma: float = 0.0, contig_bounds=None):
    """Sample `nreads` reads.  Returns (bases, rlen, truth) where bases is the
    concatenation of the reads (0..3), rlen[i] their lengths and truth an (n,3)
    array (start, src_len, strand) for diagnostics.

    If `contig_bounds` (sorted contig start offsets + total) is given, reads never
    straddle a contig boundary."""
    rng = np.random.default_rng(seed)
    G = genome.size
    if lognormal_sigma > 0:
        L = np.exp(rng.normal(np.log(src_len), lognormal_sigma, size=nreads)).astype(np.int64)
        L = np.clip(L, 500, None)
    else:
        L = np.full(nreads, src_len, dtype=np.int64)
    if contig_bounds is None:
        contig_bounds = np.array([0, G], dtype=np.int64)
    cb = np.asarray(contig_bounds, dtype=np.int64)
    clen = np.diff(cb)
    # pick a contig proportional to its length, then a start inside it
    c = np.searchsorted(np.cumsum(clen) / clen.sum(), rng.random(nreads), side="right")
    c = np.clip(c, 0, clen.size - 1)
    L = np.minimum(L, clen[c])
    start = cb[c] + (rng.random(nreads) * (clen[c] - L + 1)).astype(np.int64)
    strand = rng.integers(0, 2, size=nreads, dtype=np.int64)
    total = int(L.sum())
    seg_of = np.repeat(np.arange(nreads, dtype=np.int64), L)
    first = np.cumsum(L) - L
    within = np.arange(total, dtype=np.int64) - first[seg_of]
    fwd = strand[seg_of] == 0
    pos = np.where(fwd, start[seg_of] + within, start[seg_of] + L[seg_of] - 1 - within)
    src = genome[pos]
    src = np.where(fwd, src, 3 - src).astype(np.uint8)
    bases, rlen = _apply_errors(src, seg_of, nreads, error, rng)
    truth = np.stack([start, L, strand], axis=1)
    return bases, rlen, truth


def make_chimeric_reads(genome: np.ndarray, nreads: int, seed: int = 3, piece_len=(2500, 6000),
                        error: float = 0.15, bad_error: float = 0.30, bad_len=(1000, 2000)):
    """Config-5 stress reads: 2-3 genome segments joined (chimeras), each read also carries
    1-2 kbp stretches at 30 % error (low-quality drop-outs)."""
    rng = np.random.default_rng(seed)
    G = genome.size
    npieces = rng.integers(2, 4, size=nreads)
    P = int(npieces.sum())
    plen = rng.integers(piece_len[0], piece_len[1] + 1, size=P).astype(np.int64)
    pstart = (rng.random(P) * (G - plen)).astype(np.int64)
    pstrand = rng.integers(0, 2, size=P, dtype=np.int64)
    read_of_piece = np.repeat(np.arange(nreads, dtype=np.int64), npieces)
    total = int(plen.sum())
    piece_of = np.repeat(np.arange(P, dtype=np.int64), plen)
    first = np.cumsum(plen) - plen
    within = np.arange(total, dtype=np.int64) - first[piece_of]
    fwd = pstrand[piece_of] == 0
    pos = np.where(fwd, pstart[piece_of] + within, pstart[piece_of] + plen[piece_of] - 1 - within)
    src = genome[pos]
    src = np.where(fwd, src, 3 - src).astype(np.uint8)
    seg_of = read_of_piece[piece_of]
    err = np.full(total, error, dtype=np.float64)
    # one bad stretch inside every piece longer than 3 kbp
    blen = rng.integers(bad_len[0], bad_len[1] + 1, size=P).astype(np.int64)
    has_bad = plen > 3000
    boff = (rng.random(P) * np.maximum(plen - blen, 1)).astype(np.int64)
    bad = has_bad[piece_of] & (within >= boff[piece_of]) & (within < (boff + blen)[piece_of])
    err[bad] = bad_error
    bases, rlen = _apply_errors(src, seg_of, nreads, err, rng)
    return bases, rlen, None


# BASELINE.json configs -> generator parameters.  `scale` shrinks a config for tests.
CONFIGS = {
    "C1": dict(genome=1_000_000, contigs=2, nreads=2000, kind="plain"),
    "C2": dict(genome=4_600_000, contigs=2, nreads=13800, kind="plain"),
    "C3": dict(genome=100_000_000, contigs=4, nreads=50000, kind="repeat"),
    "C4": dict(genome=250_000_000, contigs=8, nreads=200000, kind="plain"),
    "C5": dict(genome=2_000_000, contigs=2, nreads=3000, kind="chimeric"),
}


def make_config(name: str, scale: float = 1.0, seed: int = 7):
    """Return (contigs: list[np.ndarray], reads_bases, reads_rlen) for a named config."""
    cfg = CONFIGS[name]
    G = max(20000, int(cfg["genome"] * scale))
    R = max(8, int(cfg["nreads"] * scale))
    if cfg["kind"] == "repeat":
        genome = make_repeat_genome(G, seed=seed)
    else:
        genome = make_genome(G, seed=seed)
    nc = cfg["contigs"]
    cuts = [0] + [int(G * (i + 1) / nc) for i in range(nc)]
    # unequal contigs make contig-index bugs visible
    if nc == 2:
        cuts = [0, int(G * 0.6), G]
    contigs = [genome[cuts[i]:cuts[i + 1]] for i in range(nc)]
    if cfg["kind"] == "chimeric":
        bases, rlen, _ = make_chimeric_reads(genome, R, seed=seed + 1)
    else:
        bases, rlen, _ = make_reads(genome, R, seed=seed + 1, contig_bounds=np.array(cuts))
    return contigs, bases, rlen
