"""CPU tests of the oracle: against the committed golden vectors (outputs of the unmodified
reference, tests/golden/, made by oracle/make_golden.py) and, where oracle/_ref is present,
against the reference run live."""
import os
import shutil
import tempfile

import numpy as np
import pytest

from conftest import ROOT, base_freq, make_case

GOLDEN = os.path.join(ROOT, "tests", "golden")


def _oracle_run(orc, case):
    from oracle.make_golden import CASES, input_digest
    cfg, scale, seed, flags, kw = CASES[case]
    contigs, rb, rl, rd, rf, rc = make_case(cfg, scale, seed)
    out = orc.map_block(orc.HostBlock(*rd), [(orc.HostBlock(*rf), orc.HostBlock(*rc))],
                        orc.HostBlock(*rf), freq=base_freq(contigs), **kw)
    return out, input_digest(contigs, rb, rl)


@pytest.mark.parametrize("case", ["c1_default", "c5_cover_profile", "c3_repeat_n95",
                                  "c1_k16_s50_t20_e80", "c1_k24_s200"])
def test_oracle_matches_golden(oracle_mod, case):
    g = np.load(os.path.join(GOLDEN, case + ".npz"))
    out, digest = _oracle_run(oracle_mod, case)
    assert digest == str(g["digest"]), "synthetic generator drifted: regenerate the fixtures"
    assert out["a"] == g["a"].tobytes()
    assert out["b"] == g["b"].tobytes()
    assert out["prof"] == g["prof"].tobytes()
    assert out["stats"]["h2"] == 0        # the reference's out-of-bounds read (H2) never fires


def test_golden_records_pass_check_trace_points():
    """The reference's only invariant checker (Check_Trace_Points, align.c:3194-3236)."""
    from damapper_b200 import las
    for case, ts in (("c1_default", 100), ("c5_cover_profile", 100), ("c1_k16_s50_t20_e80", 50),
                     ("c1_k24_s200", 200)):
        g = np.load(os.path.join(GOLDEN, case + ".npz"))
        for fam in ("a", "b"):
            recs = las.stream_records(g[fam].tobytes(), ts)
            assert las.check_trace_points(recs, ts) == 0


def test_sort_kmers_is_composite_key_order(oracle_mod):
    """Sort_Kmers output == sort by (code, read, rpos)."""
    orc = oracle_mod
    _, _, _, rd, _, _ = make_case("C1", 0.01, 5)
    idx = orc.sort_kmers(orc.HostBlock(*rd), 20)[:-2]
    order = np.lexsort((idx["rpos"], idx["read"], idx["code"]))
    assert (order == np.arange(len(idx))).all()
    assert idx["rpos"].min() >= 19


def test_seeds_are_fully_ordered(oracle_mod):
    """Sorted seeds == sort by (aread, bread, apos, bpos), no duplicates (item 2)."""
    orc = oracle_mod
    _, _, _, rd, rf, _ = make_case("C1", 0.02, 6)
    hr, hg = orc.HostBlock(*rd), orc.HostBlock(*rf)
    ir, ig = orc.sort_kmers(hr, 20), orc.sort_kmers(hg, 20)
    seeds, nh, lim, histo = orc.merge_join(ir, ig, 64 << 30, hr.sizeof_db, hg.sizeof_db,
                                           hr.maxlen, hr.nreads, hg.nreads)
    s = seeds[:nh]
    bpos = s["apos"] - s["diag"]
    order = np.lexsort((bpos, s["apos"], s["bread"], s["aread"]))
    assert (order == np.arange(nh)).all()
    assert lim == 10000 and nh == int((np.arange(10000) * histo).sum())
    assert seeds[nh]["aread"] == 0x7fffffff


def test_limit_tracks_memory_budget(oracle_mod):
    """map.c:2992-3015: a small -M lowers the run-product cap; -M0 removes it."""
    orc = oracle_mod
    _, _, _, rd, rf, _ = make_case("C3", 0.002, 9)
    hr, hg = orc.HostBlock(*rd), orc.HostBlock(*rf)
    ir, ig = orc.sort_kmers(hr, 12), orc.sort_kmers(hg, 12)
    args = (hr.sizeof_db, hg.sizeof_db, hr.maxlen, hr.nreads, hg.nreads)
    _, n_big, lim_big, _ = orc.merge_join(ir, ig, 64 << 30, *args)
    tight = hr.sizeof_db + hg.sizeof_db + 16 * (len(ir) + len(ig)) + 16 * (n_big // 2)
    _, n_small, lim_small, _ = orc.merge_join(ir, ig, tight, *args)
    _, n_all, lim_all, _ = orc.merge_join(ir, ig, 0, *args)
    assert 1 < lim_small < lim_big == 10000 and n_small < n_big <= n_all
    assert lim_all == 0x7fffffff


def test_align_spec_constants(oracle_mod):
    """-e.85 -> ave_path 51; the C double expressions give 199 / 99 for
    -e.8 / -e.9 (mscore = -table[...]-style check through score[1] = +mscore)."""
    orc = oracle_mod
    ap, score, table = orc.align_spec(0.85, (.25, .25, .25, .25))
    assert ap == 51 and score[1] - score[0] == 1000 and score[0x7fff] == 15 * 150
    _, s8, _ = orc.align_spec(0.8, (.25, .25, .25, .25))
    _, s9, _ = orc.align_spec(0.9, (.25, .25, .25, .25))
    assert s8[0x7fff] == 15 * 199 and s9[0x7fff] == 15 * 99


def test_empty_and_degenerate_inputs(oracle_mod):
    orc = oracle_mod
    from damapper_b200 import dazzdb
    rng = np.random.default_rng(1)
    # reads that share nothing with the reference: no seeds, no records
    ref = [rng.integers(0, 4, 5000, dtype=np.uint8)]
    reads = [rng.integers(0, 4, 700, dtype=np.uint8) for _ in range(3)]
    rd, rf = dazzdb.load_block(reads), dazzdb.load_block(ref)
    rc = dazzdb.load_block(dazzdb.revcomp_contigs(ref))
    out = orc.map_block(orc.HostBlock(*rd), [(orc.HostBlock(*rf), orc.HostBlock(*rc))],
                        orc.HostBlock(*rf))
    assert out["a"] == b"" and out["anrec"] == 0
    # a read identical to a stretch of the reference maps with 0 differences end to end
    reads = [ref[0][1000:3000].copy()]
    rd = dazzdb.load_block(reads)
    out = orc.map_block(orc.HostBlock(*rd), [(orc.HostBlock(*rf), orc.HostBlock(*rc))],
                        orc.HostBlock(*rf))
    from damapper_b200 import las
    recs = las.stream_records(out["a"], 100)
    assert len(recs) == 1 and recs[0]["diffs"] == 0
    assert (recs[0]["abpos"], recs[0]["aepos"], recs[0]["bbpos"], recs[0]["bepos"]) == (0, 2000, 1000, 3000)
    assert recs[0]["flags"] == 0x14        # START | BEST


def _mask_case(orc_or_api, name):
    """Blocks of a -m golden case: (reads, [ref fwd (, ref rc)], whole ref, freq, kwargs)."""
    from damapper_b200 import dazzdb
    from oracle import make_golden as mg
    cfg, scale, seed, flags, kw = mg.MASK_CASES[name]
    contigs, rb, rl, rd, rf, rc = make_case(cfg, scale, seed)
    gd, rdust, rtan = mg.mask_tracks(contigs, rl, seed)
    rm = dazzdb.union_masks(rdust, rtan)
    gc = dazzdb.mirror_masks(gd[0], gd[1], rf[2])
    return contigs, rb, rl, (rd, rm), (rf, gd), (rc, gc), kw


def test_oracle_masks_match_golden(oracle_mod):
    """-m: the oracle's masked extraction + mirrored reference mask against the unmodified
    reference run with -mdust -mtan (two tracks on the reads: merged)."""
    from oracle.make_golden import input_digest
    orc = oracle_mod
    contigs, rb, rl, (rd, rm), (rf, gd), (rc, gc), kw = _mask_case(orc, "c1_masks")
    g = np.load(os.path.join(GOLDEN, "c1_masks.npz"))
    assert input_digest(contigs, rb, rl) == str(g["digest"])
    out = orc.map_block(orc.HostBlock(*rd, mask=rm), [(orc.HostBlock(*rf, mask=gd), orc.HostBlock(*rc, mask=gc))],
                        orc.HostBlock(*rf), freq=base_freq(contigs), **kw)
    assert out["a"] == g["a"].tobytes() and out["b"] == g["b"].tobytes()


def _have_ref():
    from oracle import run_ref
    return run_ref.have_ref()


@pytest.mark.skipif(not _have_ref(), reason="oracle/_ref (compiled reference) not present")
@pytest.mark.parametrize("cfg,scale,seed,flags,kw", [
    ("C1", 0.03, 21, ("-C", "-p"), dict(do_b=1, profile=1)),
    ("C5", 0.05, 22, ("-C", "-n.8"), dict(do_b=1, best_tie=0.8)),
    ("C3", 0.003, 23, ("-p", "-n.7", "-k18"), dict(profile=1, best_tie=0.7, kmer=18)),
    # flag sets drawn by tools/oracle_fuzz.py (92 random cases against the reference, no mismatch)
    ("C1", 0.02, 24, ("-k14", "-n.7", "-M4", "-s126"), dict(kmer=14, best_tie=0.7, mem_limit=4 << 30, spacing=126)),
    ("C3", 0.002, 25, ("-k28", "-C", "-n.95", "-e.9", "-t20"), dict(kmer=28, do_b=1, best_tie=0.95, ave_corr=0.9, suppress=20)),
    ("C1", 0.02, 26, ("-e.75", "-M0", "-C", "-s200"), dict(ave_corr=0.75, mem_limit=0, do_b=1, spacing=200)),
    ("C5", 0.04, 27, ("-k32", "-t3", "-e.7", "-s50"), dict(kmer=32, suppress=3, ave_corr=0.7, spacing=50)),
])
def test_oracle_matches_reference_live(oracle_mod, cfg, scale, seed, flags, kw):
    from damapper_b200 import dazzdb, las
    from oracle import run_ref
    orc = oracle_mod
    contigs, rb, rl, rd, rf, rc = make_case(cfg, scale, seed)
    wd = tempfile.mkdtemp(prefix="orc_ref_")
    try:
        dazzdb.write_db(os.path.join(wd, "ref.dam"), contigs, is_dam=True)
        dazzdb.write_db(os.path.join(wd, "reads.db"), (rb, rl))
        r = run_ref.run_damapper(wd, "ref.dam", "reads.db", flags=flags, threads=2)
        ref_a = las.canonical_stream(r["m_files"])
        ref_b = las.canonical_stream(r["r_files"]) if r["r_files"] else b""
        ref_p = open(r["prof_data"], "rb").read() if r["prof_data"] else b""
    finally:
        shutil.rmtree(wd, ignore_errors=True)
    out = orc.map_block(orc.HostBlock(*rd), [(orc.HostBlock(*rf), orc.HostBlock(*rc))],
                        orc.HostBlock(*rf), freq=base_freq(contigs), **kw)
    assert out["a"] == ref_a and out["b"] == ref_b and out["prof"] == ref_p


def test_stream_db_writer_equals_write_db(tmp_path):
    """dazzdb.StreamDBWriter (chunked appends, used for the full-size runs) writes the same stub,
    .idx and .bps as write_db."""
    import filecmp
    from damapper_b200 import dazzdb, synth
    contigs, rb, rl = synth.make_config("C1", scale=0.02, seed=3)
    d1, d2 = str(tmp_path / "a"), str(tmp_path / "b")
    os.makedirs(d1); os.makedirs(d2)
    dazzdb.write_db(os.path.join(d1, "reads.db"), (rb, rl), nblocks=3)
    w = dazzdb.StreamDBWriter(os.path.join(d2, "reads.db"))
    off = np.concatenate([[0], np.cumsum(rl)])
    for a, b in ((0, 7), (7, 8), (8, len(rl))):
        w.append(rb[off[a]:off[b]], rl[a:b])
    w.close(nblocks=3)
    dazzdb.write_db(os.path.join(d1, "ref.dam"), contigs, is_dam=True)
    w = dazzdb.StreamDBWriter(os.path.join(d2, "ref.dam"), is_dam=True)
    w.append(np.concatenate(contigs), [c.size for c in contigs]); w.close()
    for f in ("reads.db", ".reads.idx", ".reads.bps", "ref.dam", ".ref.idx", ".ref.bps"):
        assert filecmp.cmp(os.path.join(d1, f), os.path.join(d2, f), shallow=False), f


@pytest.mark.skipif(not _have_ref(), reason="oracle/_ref (compiled reference) not present")
@pytest.mark.parametrize("cfg,scale,seed,flags,kw", [
    ("C5", 0.04, 81, ("-k28", "-C", "-p", "-n.9", "-e.8", "-s80", "-t3"),
     dict(kmer=28, do_b=1, profile=1, best_tie=0.9, ave_corr=0.8, spacing=80, suppress=3)),
    ("C3", 0.002, 82, ("-C", "-p", "-n.7", "-M0"), dict(do_b=1, profile=1, best_tie=0.7, mem_limit=0)),
])
def test_oracle_masks_match_reference_live(oracle_mod, cfg, scale, seed, flags, kw):
    """-mdust -mtan with other flags on top (-t run suppression after masking, -k, -s, -M0), against a live
    run of the reference; drawn by `tools/oracle_fuzz.py <n> <seed> masks` (30 cases, no mismatch)."""
    from damapper_b200 import dazzdb, las
    from oracle import make_golden as mg, run_ref
    orc = oracle_mod
    contigs, rb, rl, rd, rf, rc = make_case(cfg, scale, seed)
    wd = tempfile.mkdtemp(prefix="orc_ref_")
    try:
        gd, rm = mg.write_mask_case(wd, contigs, rb, rl, seed)
        r = run_ref.run_damapper(wd, "ref.dam", "reads.db", flags=tuple(flags) + ("-mdust", "-mtan"), threads=1)
        ref_a = las.canonical_stream(r["m_files"])
        ref_b = las.canonical_stream(r["r_files"]) if r["r_files"] else b""
        ref_p = open(r["prof_data"], "rb").read() if r["prof_data"] else b""
    finally:
        shutil.rmtree(wd, ignore_errors=True)
    gc = dazzdb.mirror_masks(gd[0], gd[1], rf[2])
    out = orc.map_block(orc.HostBlock(*rd, mask=rm), [(orc.HostBlock(*rf, mask=gd), orc.HostBlock(*rc, mask=gc))],
                        orc.HostBlock(*rf), freq=base_freq(contigs), **kw)
    assert len(ref_a) > 1000
    assert out["a"] == ref_a and out["b"] == ref_b and out["prof"] == ref_p
