"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI
(damapper_b200.api -> libdamgpu.so), against the oracle on the same seeded inputs, against the
committed golden vectors of the reference, and through size-independent properties at the
benchmark's full size.  All work is integer: every comparison is bit-exact."""
import os

import numpy as np
import pytest

from conftest import ROOT, base_freq, make_case

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def api():
    from damapper_b200 import api as a
    from conftest import install_fatal_hook
    a.init()          # raises without a B200: there is no fallback
    install_fatal_hook(a)
    return a


def _gpu_vs_oracle(api, orc, cfg, scale, seed, reads_filter="auto", **kw):
    contigs, rb, rl, rd, rf, rc = make_case(cfg, scale, seed)
    freq = base_freq(contigs)
    o = orc.map_block(orc.HostBlock(*rd), [(orc.HostBlock(*rf), orc.HostBlock(*rc))],
                      orc.HostBlock(*rf), freq=freq, **kw)
    g = api.map_block(api.HostBlock(*rd), [api.HostBlock(*rf)], api.HostBlock(*rf), freq=freq,
                      want_candidates=True, reads_filter=reads_filter, **kw)
    oc, ojc, oj = o["candidates"]
    gc, gjc, gj = g["candidates"]
    assert oc.tobytes() == gc.tobytes(), "candidate chains differ"
    assert (ojc == gjc).all() and oj.tobytes() == gj.tobytes(), "Jump lists differ"
    assert g["a"] == o["a"], "M records differ"
    assert g["b"] == o["b"], "R records differ"
    assert g["prof"] == o["prof"], "-p track differs"
    for key in ("nalign", "nwaves", "ncells"):
        assert g["stats"][key] == o["stats"][key], key
    assert g["stats"]["trace_fails"] == 0             # Check_Trace_Points of every record, on the device
    return g


@pytest.mark.parametrize("case", ["c1_default", "c5_cover_profile", "c3_repeat_n95",
                                  "c1_k16_s50_t20_e80", "c1_k24_s200"])
def test_gpu_matches_reference_golden(api, case):
    """CUDA output == the unmodified reference's output (committed fixtures)."""
    from oracle.make_golden import CASES, input_digest
    cfg, scale, seed, flags, kw = CASES[case]
    contigs, rb, rl, rd, rf, rc = make_case(cfg, scale, seed)
    g = np.load(os.path.join(GOLDEN, case + ".npz"))
    assert input_digest(contigs, rb, rl) == str(g["digest"])
    out = api.map_block(api.HostBlock(*rd), [api.HostBlock(*rf)], api.HostBlock(*rf),
                        freq=base_freq(contigs), **kw)
    assert out["a"] == g["a"].tobytes()
    assert out["b"] == g["b"].tobytes()
    assert out["prof"] == g["prof"].tobytes()


@pytest.mark.parametrize("cfg,scale,seed,kw", [
    ("C1", 0.1, 4, dict(do_b=1)),
    ("C1", 0.25, 31, dict()),
    ("C5", 0.2, 14, dict(do_b=1, best_tie=0.8)),
    ("C5", 0.3, 32, dict(do_b=1, profile=1)),
    ("C3", 0.004, 16, dict(do_b=1, profile=1, best_tie=0.7)),
    ("C3", 0.01, 33, dict(profile=1, best_tie=0.95)),
    ("C2", 0.02, 34, dict()),
    ("C1", 0.05, 18, dict(kmer=14, ave_corr=0.8, suppress=20, do_b=1)),
    ("C1", 0.05, 35, dict(kmer=32, spacing=75)),
    ("C1", 0.05, 36, dict(kmer=12, ave_corr=0.7, spacing=126, do_b=1)),
])
def test_pipeline_matches_oracle(api, oracle_mod, cfg, scale, seed, kw):
    _gpu_vs_oracle(api, oracle_mod, cfg, scale, seed, **kw)


@pytest.mark.parametrize("tier,slots", [(0, 0)])
@pytest.mark.parametrize("cfg,scale,seed,kw", [
    ("C1", 0.1, 41, dict(do_b=1, profile=1)),
    ("C5", 0.1, 42, dict(do_b=1)),            # wide bands: many jobs fall through to the warp kernel
    ("C3", 0.004, 43, dict(best_tie=0.9)),
])
def test_alignment_tiers_give_the_same_records(api, oracle_mod, tier, slots, cfg, scale, seed, kw):
    """The warp-per-job kernel (the tier that re-runs what outgrows the default duo kernel) as the FIRST
    tier vs the oracle; every other GPU test runs the duo kernel + k_unwind."""
    contigs, rb, rl, rd, rf, rc = make_case(cfg, scale, seed)
    freq = base_freq(contigs)
    o = oracle_mod.map_block(oracle_mod.HostBlock(*rd), [(oracle_mod.HostBlock(*rf), oracle_mod.HostBlock(*rc))],
                             oracle_mod.HostBlock(*rf), freq=freq, **kw)
    L = api.load()
    L.damgpu_set_align_tier(tier, slots)
    try:
        g = api.map_block(api.HostBlock(*rd), [api.HostBlock(*rf)], api.HostBlock(*rf), freq=freq, **kw)
    finally:
        L.damgpu_set_align_tier(1, 0)
    assert g["a"] == o["a"], "M records differ"
    assert g["b"] == o["b"], "R records differ"
    assert g["prof"] == o["prof"], "-p track differs"


@pytest.mark.parametrize("cfg,scale,seed,kmer,suppress", [
    ("C1", 0.1, 4, 20, 0), ("C3", 0.002, 5, 14, 0), ("C1", 0.05, 6, 16, 10),
    ("C5", 0.1, 7, 32, 0), ("C1", 0.05, 8, 12, 0), ("C1", 0.05, 9, 9, 3),
])
def test_index_and_seeds_match_oracle(api, oracle_mod, cfg, scale, seed, kmer, suppress):
    """Sort_Kmers (a2-a5) and merge-join + seed sort (a6-a9) arrays, byte for byte."""
    orc = oracle_mod
    contigs, rb, rl, rd, rf, rc = make_case(cfg, scale, seed)
    api.set_filter_params(kmer, suppress, 4)
    api.set_options()
    hr, hg = api.HostBlock(*rd), api.HostBlock(*rf)
    dr, dg = api.DeviceBlock(hr), api.DeviceBlock(hg)
    ir, ig = api.Index(dr), api.Index(dg)
    o_r = orc.sort_kmers(orc.HostBlock(*rd), kmer, suppress)
    o_g = orc.sort_kmers(orc.HostBlock(*rf), kmer, suppress)
    assert ir.download().tobytes() == o_r.tobytes()
    assert ig.download().tobytes() == o_g.tobytes()
    s = api.Seeds(ir, dr, ig, dg)
    os_, nh, lim, histo = orc.merge_join(o_r, o_g, 64 << 30, hr.sizeof_db, hg.sizeof_db,
                                         hr.maxlen, hr.nreads, hg.nreads)
    assert s.count == nh and s.limit == lim
    assert (s.histogram() == histo).all()
    assert s.download().tobytes() == os_.tobytes()
    # complement_DB on the device (damapper.c:433-469)
    dg.complement()
    n = int(hg.boff[-1]) + 1
    assert (dg.download_bases()[:n] == rc[0][:n]).all()
    # complemented index too
    igc = api.Index(dg)
    assert igc.download().tobytes() == orc.sort_kmers(orc.HostBlock(*rc), kmer, suppress).tobytes()


@pytest.mark.parametrize("cfg,scale,seed,kmer,bits", [
    ("C1", 0.1, 51, 20, 0), ("C3", 0.004, 52, 14, 0), ("C5", 0.1, 53, 32, 0), ("C1", 0.05, 54, 12, 0),
    ("C1", 0.1, 55, 20, 12),      # a 4096-bit bitmap: nearly every k-mer is a false positive
])
def test_deferred_reads_index(api, oracle_mod, cfg, scale, seed, kmer, bits):
    """The reads-side Sort_Kmers left unbuilt: Match_Filter's merge-join runs on the sub-list of the
    records whose code occurs in the reference block (either orientation) and yields the same
    histogram, cap and seed array as the full list; asking for the list builds all of it."""
    orc = oracle_mod
    contigs, rb, rl, rd, rf, rc = make_case(cfg, scale, seed)
    api.set_filter_params(kmer, 0, 4)
    api.set_options()
    api.set_reads_filter("always", bits)
    try:
        hr, hg = api.HostBlock(*rd), api.HostBlock(*rf)
        dr, dg = api.DeviceBlock(hr), api.DeviceBlock(hg)
        ir = api.Index(dr, deferred=True)
        o_r = orc.sort_kmers(orc.HostBlock(*rd), kmer, 0)
        assert ir.is_deferred and len(ir) == o_r.size - 2
        for comp, ohb in ((0, rf), (1, rc)):
            ig = api.Index(dg)
            s = api.Seeds(ir, dr, ig, dg)
            o_g = orc.sort_kmers(orc.HostBlock(*ohb), kmer, 0)
            os_, nh, lim, histo = orc.merge_join(o_r, o_g, 64 << 30, hr.sizeof_db, hg.sizeof_db,
                                                 hr.maxlen, hr.nreads, hg.nreads)
            assert s.count == nh and s.limit == lim
            assert (s.histogram() == histo).all()
            assert s.download().tobytes() == os_.tobytes()
            if comp == 0:
                survivors = api.last_filter_times()["survivors"]
                assert 0 < survivors <= o_r.size - 2
                if bits == 0:
                    assert survivors < (o_r.size - 2) // 2
            s.free(); ig.free()
            dg.complement()
        assert ir.is_deferred                              # both orientations ran on the filtered list
        assert ir.download().tobytes() == o_r.tobytes()    # the whole list on demand
        assert not ir.is_deferred
    finally:
        api.set_reads_filter("auto")


@pytest.mark.parametrize("cfg,scale,seed,kw", [
    ("C1", 0.1, 56, dict(do_b=1, profile=1)),
    ("C3", 0.01, 57, dict(profile=1, best_tie=0.95)),
    ("C5", 0.2, 58, dict(do_b=1, best_tie=0.8, kmer=16)),
    ("C1", 0.05, 59, dict(mem_limit=0)),
])
def test_pipeline_with_filtered_reads_index(api, oracle_mod, cfg, scale, seed, kw):
    """Whole path with the filtered reads list forced on small inputs: candidates, records, -p."""
    _gpu_vs_oracle(api, oracle_mod, cfg, scale, seed, reads_filter="always", **kw)


def test_filtered_reads_index_over_reference_blocks(api, oracle_mod):
    """Three reference blocks: a filtered list per block (signature changes), results as the oracle's;
    a reads block without any k-mer of the reference gives no seeds."""
    orc = oracle_mod
    contigs, rb, rl, rd, rf, rc = make_case("C1", 0.12, 60)
    from damapper_b200 import dazzdb
    freq = base_freq(contigs)
    g = np.concatenate(contigs)
    cut = [0, g.size // 3, 2 * g.size // 3, g.size]
    parts = [[g[cut[i]:cut[i + 1]]] for i in range(3)]
    fwd, pairs, first = [], [], 0
    for p in parts:
        f, c = dazzdb.load_block(p), dazzdb.load_block(dazzdb.revcomp_contigs(p))
        fwd.append(api.HostBlock(*f, tfirst=first))
        pairs.append((orc.HostBlock(*f, tfirst=first), orc.HostBlock(*c, tfirst=first)))
        first += 1
    whole = dazzdb.load_block([p[0] for p in parts])
    o = orc.map_block(orc.HostBlock(*rd), pairs, orc.HostBlock(*whole), freq=freq, do_b=1)
    for mode in ("always", "auto", "off"):
        out = api.map_block(api.HostBlock(*rd), fwd, api.HostBlock(*whole), freq=freq, do_b=1,
                            reads_filter=mode)
        assert out["a"] == o["a"] and out["b"] == o["b"], mode
    # reads that share nothing with the reference (poly-A reads against a poly-C contig)
    api.set_filter_params(20, 0, 4)
    api.set_options()
    api.set_reads_filter("always")
    try:
        ra = dazzdb.load_block((np.zeros(4000, dtype=np.uint8), np.array([2000, 2000], dtype=np.int32)))
        gc_ = dazzdb.load_block([np.ones(3000, dtype=np.uint8)])
        dr, dg = api.DeviceBlock(api.HostBlock(*ra)), api.DeviceBlock(api.HostBlock(*gc_))
        ir, ig = api.Index(dr, deferred=True), api.Index(dg)
        s = api.Seeds(ir, dr, ig, dg)
        assert s.count == 0 and s.limit == 10000
    finally:
        api.set_reads_filter("auto")


def test_memory_cap_limit(api, oracle_mod):
    """`limit` from -M (map.c:2992-3015), incl. -M0 = no cap."""
    orc = oracle_mod
    contigs, rb, rl, rd, rf, rc = make_case("C3", 0.002, 9)
    hr, hg = api.HostBlock(*rd), api.HostBlock(*rf)
    api.set_filter_params(12, 0, 4)
    o_r = orc.sort_kmers(orc.HostBlock(*rd), 12)
    o_g = orc.sort_kmers(orc.HostBlock(*rf), 12)
    args = (hr.sizeof_db, hg.sizeof_db, hr.maxlen, hr.nreads, hg.nreads)
    _, n_big, _, _ = orc.merge_join(o_r, o_g, 64 << 30, *args)
    tight = hr.sizeof_db + hg.sizeof_db + 16 * (len(o_r) + len(o_g)) + 16 * (n_big // 2)
    dr, dg = api.DeviceBlock(hr), api.DeviceBlock(hg)
    ir, ig = api.Index(dr), api.Index(dg)
    for mem in (tight, 0):
        api.set_options(mem_limit=mem)
        s = api.Seeds(ir, dr, ig, dg)
        os_, nh, lim, _ = orc.merge_join(o_r, o_g, mem, *args)
        assert (s.count, s.limit) == (nh, lim)
        assert s.download().tobytes() == os_.tobytes()
    api.set_options()


def test_edge_cases(api, oracle_mod):
    from damapper_b200 import dazzdb, las
    orc = oracle_mod
    rng = np.random.default_rng(1)
    ref = [rng.integers(0, 4, 5000, dtype=np.uint8), rng.integers(0, 4, 333, dtype=np.uint8)]
    rf = dazzdb.load_block(ref)
    rc = dazzdb.load_block(dazzdb.revcomp_contigs(ref))

    def both(reads, **kw):
        rd = dazzdb.load_block(reads)
        o = orc.map_block(orc.HostBlock(*rd), [(orc.HostBlock(*rf), orc.HostBlock(*rc))],
                          orc.HostBlock(*rf), **kw)
        g = api.map_block(api.HostBlock(*rd), [api.HostBlock(*rf)], api.HostBlock(*rf), **kw)
        assert g["a"] == o["a"] and g["b"] == o["b"] and g["prof"] == o["prof"]
        return g

    # no shared k-mers at all: empty output
    g = both([rng.integers(0, 4, 700, dtype=np.uint8) for _ in range(3)], do_b=1, profile=1)
    assert g["a"] == b"" and g["b"] == b""
    # exact copies, forward and reverse complement, ragged lengths down to exactly k
    reads = [ref[0][1000:3000].copy(), (3 - ref[0][200:1500][::-1]).astype(np.uint8),
             ref[0][4000:4020].copy(), ref[1].copy(), ref[0][0:61].copy()]
    g = both(reads, do_b=1, profile=1)
    recs = las.stream_records(g["a"], 100)
    assert [r["diffs"] for r in recs] == [0] * len(recs) and len(recs) >= 3
    assert recs[0]["flags"] == 0x14 and recs[1]["flags"] == 0x15
    # a single read, a single contig, k-mers of length 32
    both([ref[0][500:2500].copy()], kmer=32)


def test_layer1_map_h_calls_write_las_files(api, oracle_mod, tmp_path):
    """The four map.h entry points, called as damapper.c calls them, write per-thread .las files
    whose concatenation is the oracle's record stream (T-invariant)."""
    import ctypes as C
    from damapper_b200 import dazzdb, las
    orc = oracle_mod
    contigs, rb, rl, rd, rf, rc = make_case("C5", 0.05, 41)
    freq = base_freq(contigs)
    o = orc.map_block(orc.HostBlock(*rd), [(orc.HostBlock(*rf), orc.HostBlock(*rc))],
                      orc.HostBlock(*rf), freq=freq, do_b=1, profile=1)
    L = api.load()
    api.set_options(profile=1, sort_path=str(tmp_path))
    assert L.damgpu_Set_Filter_Params(20, 0, 4) == 0
    hr, hg, hc = api.HostBlock(*rd), api.HostBlock(*rf), api.HostBlock(*rc)
    blen, alen = C.c_int(0), C.c_int(0)
    bindex = L.damgpu_Sort_Kmers(C.byref(hr.c), C.byref(blen))
    aindex = L.damgpu_Sort_Kmers(C.byref(hg.c), C.byref(alen))
    L.damgpu_Match_Filter(C.byref(hr.c), C.byref(hg.c), bindex, blen, aindex, alen, 0, 1)
    aindex = L.damgpu_Sort_Kmers(C.byref(hc.c), C.byref(alen))
    L.damgpu_Match_Filter(C.byref(hr.c), C.byref(hc.c), bindex, blen, aindex, alen, 1, 0)
    spec = api.CAlignSpec(0.85, 100, (C.c_float * 4)(*freq))
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        L.damgpu_Reporter(b"reads", C.byref(hr.c), b"ref", C.byref(hg.c), C.byref(spec), 3)
    finally:
        os.chdir(cwd)
    L.damgpu_index_free(bindex)
    m = [str(tmp_path / ("reads.ref.M%d.las" % i)) for i in range(1, 5)]
    r = [str(tmp_path / ("ref.reads.R%d.las" % i)) for i in range(1, 5)]
    assert las.canonical_stream(m) == o["a"]
    assert las.canonical_stream(r) == o["b"]
    assert open(tmp_path / ".reads.prof.data", "rb").read() == o["prof"]
    anno = np.fromfile(tmp_path / ".reads.prof.anno", dtype=np.uint8)
    assert len(anno) == 8 + 8 * (hr.nreads + 1)


def test_full_size_properties(api):
    """BASELINE config 2 at full size (4.6 Mbp + 13.8 k reads): properties that do not need the
    oracle -- sortedness of the index and of the seeds, k-mer count, histogram/limit consistency,
    Check_Trace_Points on every record, every read mapped."""
    from damapper_b200 import las
    contigs, rb, rl, rd, rf, rc = make_case("C2", 1.0, 7)
    api.set_filter_params(20, 0, 4)
    api.set_options()
    hr, hg = api.HostBlock(*rd), api.HostBlock(*rf)
    dr, dg = api.DeviceBlock(hr), api.DeviceBlock(hg)
    ir, ig = api.Index(dr), api.Index(dg)
    assert len(ir) == int(hr.boff[-1]) - 20 * hr.nreads
    idx = ir.download()
    code = idx["code"][:-2]
    assert (np.diff(code.astype(np.int64)) >= 0).all()
    tie = np.diff(code.astype(np.int64)) == 0
    key = idx["read"][:-2].astype(np.int64) * (1 << 32) + idx["rpos"][:-2]
    assert (np.diff(key)[tie] > 0).all()                       # (code, read, rpos) order
    assert idx["code"][-2] == 0xffffffffffffffff and idx["code"][-1] == 0
    # the multiset of codes is that of a direct recount on a sample of reads
    s = api.Seeds(ir, dr, ig, dg)
    seeds = s.download()[:s.count]
    skey = [seeds["apos"] - seeds["diag"], seeds["apos"], seeds["bread"], seeds["aread"]]
    assert (np.lexsort(skey) == np.arange(s.count)).all()
    assert s.limit == 10000 and s.count == int((np.arange(10000) * s.histogram()).sum())
    del idx, seeds
    s.free(); ig.free(); ir.free(); dg.free(); dr.free()
    out = api.map_block(hr, [hg], hg, freq=base_freq(contigs), do_b=1)
    ra = las.stream_records(out["a"], 100)
    rb_ = las.stream_records(out["b"], 100)
    assert las.check_trace_points(ra, 100) == 0 and las.check_trace_points(rb_, 100) == 0
    mapped = {r["aread"] for r in ra}
    assert len(mapped) >= 0.995 * hr.nreads
    assert out["anrec"] == out["bnrec"] == len(ra)
    assert out["stats"]["overflow_jobs"] == 0 or out["stats"]["overflow_jobs"] < 50


@pytest.mark.parametrize("cfg,scale,seed,kw", [
    ("C1", 0.1, 61, dict(do_b=1)),
    ("C3", 0.004, 62, dict(profile=1, best_tie=0.9)),
])
def test_mask_tracks(api, oracle_mod, cfg, scale, seed, kw):
    """-m: k-mers touching a masked base are not indexed (tuple_thread,
    map.c:481-543); the mask of the complemented reference block is mirrored (damapper.c:471-522).
    Index, complemented index and the whole pipeline against the oracle."""
    from damapper_b200 import dazzdb
    contigs, rb, rl, rd, rf, rc = make_case(cfg, scale, seed)
    freq = base_freq(contigs)
    mr = dazzdb.random_masks(rd[2], seed=seed, max_intervals=3, max_len=600)
    mg = dazzdb.random_masks(rf[2], seed=seed + 1, max_intervals=40, max_len=2000)
    mc = dazzdb.mirror_masks(mg[0], mg[1], rf[2])
    O, A = oracle_mod.HostBlock, api.HostBlock
    api.set_filter_params(20, 0, 4)
    api.set_options()
    # index of the masked reads block, and of the masked reference in both orientations
    dr = api.DeviceBlock(A(*rd, mask=mr))
    ir = api.Index(dr)
    oi = oracle_mod.sort_kmers(O(*rd, mask=mr), 20, 0)
    assert ir.download().tobytes() == oi.tobytes()
    assert len(ir) < int(rd[1][-1]) - 20 * len(rd[2])          # something was masked
    dg = api.DeviceBlock(A(*rf, mask=mg))
    ig = api.Index(dg)
    assert ig.download().tobytes() == oracle_mod.sort_kmers(O(*rf, mask=mg), 20, 0).tobytes()
    ig.free(); dg.complement(); ig = api.Index(dg)
    assert ig.download().tobytes() == oracle_mod.sort_kmers(O(*rc, mask=mc), 20, 0).tobytes()
    ig.free(); dg.free(); ir.free(); dr.free()
    # whole pipeline (the Reporter sees the unmasked whole reference, damapper.c:870)
    o = oracle_mod.map_block(O(*rd, mask=mr), [(O(*rf, mask=mg), O(*rc, mask=mc))], O(*rf), freq=freq, **kw)
    g = api.map_block(A(*rd, mask=mr), [A(*rf, mask=mg)], A(*rf), freq=freq, **kw)
    assert g["anrec"] == o["anrec"] > 0
    assert g["a"] == o["a"] and g["b"] == o["b"] and g["prof"] == o["prof"]


def test_mask_tracks_match_reference_golden(api, tmp_path):
    """-m end to end against the unmodified reference (fixture c1_masks: damapper -C -mdust -mtan):
    through the API with the merged masks, and through the host driver, which reads the .anno/.data
    track files itself and merges the two tracks of the reads DB."""
    import subprocess
    from damapper_b200 import dazzdb, las
    from oracle import make_golden as mg
    cfg, scale, seed, flags, kw = mg.MASK_CASES["c1_masks"]
    contigs, rb, rl, rd, rf, rc = make_case(cfg, scale, seed)
    g = np.load(os.path.join(GOLDEN, "c1_masks.npz"))
    assert mg.input_digest(contigs, rb, rl) == str(g["digest"])
    gd, rdust, rtan = mg.mask_tracks(contigs, rl, seed)
    rm = dazzdb.union_masks(rdust, rtan)
    out = api.map_block(api.HostBlock(*rd, mask=rm), [api.HostBlock(*rf, mask=gd)], api.HostBlock(*rf),
                        freq=base_freq(contigs), **kw)
    assert out["a"] == g["a"].tobytes() and out["b"] == g["b"].tobytes()
    # the driver
    exe = os.path.join(ROOT, "damapper_b200", "damapper")
    wd = str(tmp_path)
    mg.write_mask_case(wd, contigs, rb, rl, seed)
    bindir, keep = os.path.join(wd, "bin"), os.path.join(wd, "keep")
    os.makedirs(bindir); os.makedirs(keep); os.makedirs(os.path.join(wd, "tmp"))
    with open(os.path.join(bindir, "LAsort"), "w") as f:
        f.write('#!/bin/bash\nfor a in "$@"; do case "$a" in -*) ;; *) pat="$a";; esac; done\n'
                'for f in ${pat/@/[0-9]*}; do [ -e "$f" ] && cp "$f" "%s"/; done\nexit 0\n' % keep)
    for n in ("LAcat", "LAmerge"):
        with open(os.path.join(bindir, n), "w") as f:
            f.write("#!/bin/bash\nexit 0\n")
    for n in ("LAsort", "LAcat", "LAmerge"):
        os.chmod(os.path.join(bindir, n), 0o755)
    env = dict(os.environ)
    env["PATH"] = bindir + os.pathsep + env["PATH"]
    p = subprocess.run([exe, "-T4", "-P" + os.path.join(wd, "tmp")] + list(flags) + ["-mnone", "ref.dam", "reads.db"],
                       cwd=wd, env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr
    assert "Track none given but never used" in p.stdout
    m = [os.path.join(keep, "reads.ref.M%d.las" % i) for i in range(1, 5)]
    r = [os.path.join(keep, "ref.reads.R%d.las" % i) for i in range(1, 5)]
    assert las.canonical_stream(m) == g["a"].tobytes()
    assert las.canonical_stream(r) == g["b"].tobytes()


def test_packed_bps_upload_decodes_on_the_device(api, oracle_mod):
    """the .bps image (2 bits per base) expanded on the device gives the
    Load_All_Reads image byte for byte, for ragged read lengths (all residues mod 4), and the
    index built from it is the index of the plain upload."""
    rng = np.random.default_rng(77)
    lens = np.array([20, 21, 22, 23, 24, 57, 1001, 4099, 10000, 33], dtype=np.int32)
    bases = rng.integers(0, 4, size=int(lens.sum()), dtype=np.uint8)
    from damapper_b200 import dazzdb
    rd = dazzdb.load_block((bases, lens))
    hb = api.HostBlock(*rd)
    api.set_filter_params(20, 0, 4)
    api.set_options()
    plain, packed = api.DeviceBlock(hb), api.DeviceBlock(hb, packed=True)
    a, b = plain.download_bases(), packed.download_bases()
    n1 = int(hb.boff[-1]) + 1                            # leading 4 + every read and its terminator
    bad = np.nonzero(a[:n1] != b[:n1])[0]
    assert bad.size == 0, ("packed != plain", bad[:8], a[bad[:8]], b[bad[:8]])
    assert b[:n1].tobytes() == hb.bases[:n1].tobytes()
    ia, ib = api.Index(plain), api.Index(packed)
    assert ia.download().tobytes() == ib.download().tobytes()
    ia.free(); ib.free(); plain.free(); packed.free()
    empty = api.HostBlock(np.array([4], dtype=np.uint8), np.zeros(1, dtype=np.int64), np.zeros(0, dtype=np.int32))
    e = api.DeviceBlock(empty, packed=True)
    assert e.download_bases().tobytes() == empty.bases.tobytes()
    e.free()


def test_full_size_c2_matches_oracle(api, oracle_mod):
    """BASELINE config 2 at FULL size (4.6 Mbp + 13.8 k reads, the bench workload): every M and R
    record and the -p track bit for bit against the oracle (about half a minute of one host core)."""
    contigs, rb, rl, rd, rf, rc = make_case("C2", 1.0, 7)
    freq = base_freq(contigs)
    kw = dict(do_b=1, profile=1)
    o = oracle_mod.map_block(oracle_mod.HostBlock(*rd), [(oracle_mod.HostBlock(*rf), oracle_mod.HostBlock(*rc))],
                             oracle_mod.HostBlock(*rf), freq=freq, **kw)
    g = api.map_block(api.HostBlock(*rd), [api.HostBlock(*rf)], api.HostBlock(*rf), freq=freq, **kw)
    assert g["anrec"] == o["anrec"] > 13000
    assert g["a"] == o["a"], "M records differ"
    assert g["b"] == o["b"], "R records differ"
    assert g["prof"] == o["prof"], "-p track differs"
    for key in ("nalign", "nwaves", "ncells"):
        assert g["stats"][key] == o["stats"][key], key


def test_host_driver_cli(api, tmp_path):
    """The C host driver (damapper_b200/damapper): reference command line, DAZZ_DB input read
    by its own loader, per-thread .las + .prof output equal to the reference's golden stream."""
    import shutil
    import subprocess
    from damapper_b200 import dazzdb, las
    from oracle.make_golden import CASES
    exe = os.path.join(ROOT, "damapper_b200", "damapper")
    assert os.path.exists(exe), "host driver not built"
    cfg, scale, seed, flags, kw = CASES["c5_cover_profile"]
    contigs, rb, rl, rd, rf, rc = make_case(cfg, scale, seed)
    g = np.load(os.path.join(GOLDEN, "c5_cover_profile.npz"))
    wd = str(tmp_path)
    dazzdb.write_db(os.path.join(wd, "ref.dam"), contigs, is_dam=True)
    dazzdb.write_db(os.path.join(wd, "reads.db"), (rb, rl))
    # LAsort/LAcat/LAmerge are DALIGNER programs the driver shells out to (as the reference
    # does, damapper.c:893-911); stubs that keep the per-thread files stand in for them
    bindir = os.path.join(wd, "bin")
    keep = os.path.join(wd, "keep")
    os.makedirs(bindir); os.makedirs(keep); os.makedirs(os.path.join(wd, "tmp"))
    with open(os.path.join(bindir, "LAsort"), "w") as f:
        f.write('#!/bin/bash\nfor a in "$@"; do case "$a" in -*) ;; *) pat="$a";; esac; done\n'
                'for f in ${pat/@/[0-9]*}; do [ -e "$f" ] && cp "$f" "%s"/; done\nexit 0\n' % keep)
    for n in ("LAcat", "LAmerge"):
        with open(os.path.join(bindir, n), "w") as f:
            f.write("#!/bin/bash\nexit 0\n")
    for n in ("LAsort", "LAcat", "LAmerge"):
        os.chmod(os.path.join(bindir, n), 0o755)
    env = dict(os.environ)
    env["PATH"] = bindir + os.pathsep + env["PATH"]
    p = subprocess.run([exe, "-T4", "-P" + os.path.join(wd, "tmp")] + list(flags) + ["ref.dam", "reads.db"],
                       cwd=wd, env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr
    m = [os.path.join(keep, "reads.ref.M%d.las" % i) for i in range(1, 5)]
    r = [os.path.join(keep, "ref.reads.R%d.las" % i) for i in range(1, 5)]
    assert las.canonical_stream(m) == g["a"].tobytes()
    assert las.canonical_stream(r) == g["b"].tobytes()
    assert open(os.path.join(wd, ".reads.prof.data"), "rb").read() == g["prof"].tobytes()
    # the same run without the DALIGNER programs: the driver sorts and merges the per-thread files
    # itself (las_post.c) and leaves the final reads.ref.las / ref.reads.las
    env2 = dict(env)
    env2["DAMGPU_BUILTIN_SORT"] = "1"
    os.remove(os.path.join(wd, ".reads.prof.data"))
    p = subprocess.run([exe, "-v", "-T4", "-P" + os.path.join(wd, "tmp")] + list(flags) + ["ref.dam", "reads.db"],
                       cwd=wd, env=env2, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr
    assert "built-in sort" in p.stdout
    for fname, stream in (("reads.ref.las", g["a"].tobytes()), ("ref.reads.las", g["b"].tobytes())):
        ts, recs = las.read_las(os.path.join(wd, fname))
        want = las.stream_records(stream, ts)
        assert len(recs) == len(want) > 0

        def chains(rs):
            out = []
            for r in rs:
                key = (r["tlen"], r["diffs"], r["abpos"], r["bbpos"], r["aepos"], r["bepos"], r["flags"],
                       r["aread"], r["bread"], r["trace"].tobytes())
                if r["flags"] & 0x8 and out:          # NEXT: continues the chain
                    out[-1].append(key)
                else:
                    out.append([key])
            return out
        got_c, want_c = chains(recs), chains(want)
        assert sorted(map(tuple, got_c)) == sorted(map(tuple, want_c))     # same chains, whole
        heads = [(c[0][7], c[0][2]) for c in got_c]                         # (aread, abpos), -a order
        assert heads == sorted(heads)
        assert las.check_trace_points(recs, ts) == 0
    # error behaviour of the reference CLI is kept
    p = subprocess.run([exe, "-N", "ref.dam", "reads.db"], cwd=wd, env=env, capture_output=True, text=True)
    assert p.returncode == 1 and "Cannot specify N flag without C also" in p.stderr
    p = subprocess.run([exe, "-k33", "ref.dam", "reads.db"], cwd=wd, env=env, capture_output=True, text=True)
    assert p.returncode == 1 and "K-mer length must be 32 or less" in p.stderr


@pytest.mark.gpu
def test_host_driver_multiblock_reference_cache(api, tmp_path):
    """Several reads blocks against a reference split into two blocks: the driver keeps the reference
    indices resident between reads blocks (DAMGPU_REF_CACHE) and prefetches the next reads block; the
    per-thread .las streams and the -p tracks must equal those of the uncached run (the reference's
    own schedule, damapper.c:839-863) and, when the compiled reference is present, the reference's."""
    import glob
    import shutil
    import subprocess
    from damapper_b200 import dazzdb, las, synth
    from oracle import run_ref
    exe = os.path.join(ROOT, "damapper_b200", "damapper")
    assert os.path.exists(exe), "host driver not built"
    contigs, rb, rl = synth.make_config("C1", scale=0.06, seed=31)
    wd = str(tmp_path)
    dazzdb.write_db(os.path.join(wd, "ref.dam"), contigs, is_dam=True, nblocks=2)
    w = dazzdb.StreamDBWriter(os.path.join(wd, "reads.db"))
    off = np.concatenate([[0], np.cumsum(rl)])
    half = len(rl) // 2
    w.append(rb[:off[half]], rl[:half]); w.append(rb[off[half]:], rl[half:])
    w.close(nblocks=3)
    bindir = os.path.join(wd, "bin"); os.makedirs(bindir)
    with open(os.path.join(bindir, "LAsort"), "w") as f:
        f.write('#!/bin/bash\nfor a in "$@"; do case "$a" in -*) ;; *) pat="$a";; esac; done\n'
                'for f in ${pat/@/[0-9]*}; do [ -e "$f" ] && cp "$f" "$DAMAPPER_KEEP_DIR"/; done\nexit 0\n')
    for n in ("LAcat", "LAmerge"):
        with open(os.path.join(bindir, n), "w") as f:
            f.write("#!/bin/bash\nexit 0\n")
    for n in ("LAsort", "LAcat", "LAmerge"):
        os.chmod(os.path.join(bindir, n), 0o755)

    def run(binary, tag, extra_env, flags):
        keep = os.path.join(wd, "keep_" + tag); os.makedirs(keep)
        tmp = os.path.join(wd, "tmp_" + tag); os.makedirs(tmp)
        env = dict(os.environ, DAMAPPER_KEEP_DIR=keep, **extra_env)
        env["PATH"] = bindir + os.pathsep + env["PATH"]
        p = subprocess.run([binary, "-T4", "-P" + tmp, "-M16"] + flags + ["ref.dam", "reads.1", "reads.2", "reads.3"],
                           cwd=wd, env=env, capture_output=True, text=True, timeout=900)
        assert p.returncode == 0, p.stderr
        out = {}
        for b in (1, 2, 3):
            m = run_ref._thread_sorted(glob.glob(os.path.join(keep, "reads.%d.ref.M[0-9]*.las" % b)))
            r = run_ref._thread_sorted(glob.glob(os.path.join(keep, "ref.reads.%d.R[0-9]*.las" % b)))
            assert len(m) == 4 and len(r) == 4
            prof = os.path.join(wd, ".reads.%d.prof.data" % b)
            pdata = b""
            if "-p" in flags:
                pdata = open(prof, "rb").read()
                os.remove(prof)
            out[b] = (las.canonical_stream(m), las.canonical_stream(r), pdata)
        return out, p.stderr

    cached, err = run(exe, "cached", {"DAMGPU_TIMING": "1"}, ["-C"])
    assert "match, both strands (cached)" in err, "the reference cache was not used"
    plain, err = run(exe, "plain", {"DAMGPU_REF_CACHE": "0", "DAMGPU_TIMING": "1"}, ["-C"])
    assert "(cached)" not in err
    assert sum(len(v[0]) for v in cached.values()) > 10000
    assert cached == plain
    if run_ref.have_ref():
        ref, _ = run(run_ref.REF_BIN, "ref", {}, ["-C"])
        assert cached == ref
    # with -p as well (the unmodified reference segfaults on -p with a reference of two blocks and
    # several reads blocks, so this half has no third arm)
    cached, _ = run(exe, "cached_p", {}, ["-C", "-p"])
    plain, _ = run(exe, "plain_p", {"DAMGPU_REF_CACHE": "0"}, ["-C", "-p"])
    assert all(len(v[2]) > 0 for v in cached.values())
    assert cached == plain



def test_output_pools_grow_over_several_overflow_rounds(api, oracle_mod):
    """-e.9 on 15 % reads ends alignments early: a candidate chain yields many short alignments, far more than the
    two records per job the output pools are first sized for.  The jobs that find the pools full are re-run by the
    warp kernel with doubled pools, and again, until they fit (report.cu overflow rounds).  A case drawn by
    tools/gpu_fuzz.py (k=14, two reference blocks, -C -p): candidates, Jump lists, both record families and the
    -p track as the oracle's."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gpu_fuzz
    res, info = gpu_fuzz.run_case(api, oracle_mod, "C2", 0.03769, 723234,
                                  dict(kmer=14, do_b=1, profile=1, ave_corr=0.9), 2, "always")
    assert info["overflow_jobs"] > 100 and info["records"] > 1000
    for key in ("candidates", "jumps", "M", "R", "prof", "trace_check"):
        assert res[key], key
