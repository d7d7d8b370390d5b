"""GPU tests of the command lines: the compiled drop-in (the reference's own damapper.c re-linked
against libdamgpu), the -G multi-GPU mode of the shipped driver, block-level mask tracks, and the
on-device Check_Trace_Points verifier rejecting a corrupted record."""
import glob
import os

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

EXE = os.path.join(ROOT, "damapper_b200", "damapper")


def _streams(r, with_prof=True):
    from damapper_b200 import las
    prof = b""
    if with_prof and r["prof_data"]:
        prof = open(r["prof_data"], "rb").read()
        os.remove(r["prof_data"])
    return (las.canonical_stream(r["m_files"]) if r["m_files"] else b"",
            las.canonical_stream(r["r_files"]) if r["r_files"] else b"", prof)


@pytest.mark.parametrize("cfg,scale,seed,flags", [
    ("C1", 0.08, 51, ()),
    ("C5", 0.15, 52, ("-C", "-p")),
    # no -p here: on this repeat-rich input the UNMODIFIED reference itself is not deterministic in its
    # -p track under -T4 (one run in six returned a track of all 40s, same .las records; SURVEY hazard list)
    ("C3", 0.004, 53, ("-n.9", "-C", "-k16", "-s80")),
    ("C1", 0.05, 54, ("-t20", "-e.8", "-C")),
])
def test_reference_driver_relinked_against_libdamgpu(tmp_path, cfg, scale, seed, flags):
    """map.h:25-39 as a drop-in: /root/reference/damapper.c + DB.c + QV.c + align.c, compiled unmodified,
    linked with oracle/map_gpu.c (the INTEGRATION.md shim) and libdamgpu.so instead of map.c, against the
    unmodified reference on the same databases: same .las record streams, same -p track."""
    from damapper_b200 import dazzdb, synth
    from oracle import run_ref
    gpu = os.path.join(run_ref.REF_DIR, "damapper_gpu")
    if not (run_ref.have_ref() and os.access(gpu, os.X_OK)):
        pytest.skip("oracle/_ref/damapper_gpu is not built (needs /root/reference at build time)")
    contigs, rb, rl = synth.make_config(cfg, scale=scale, seed=seed)
    wd = str(tmp_path)
    dazzdb.write_db(os.path.join(wd, "ref.dam"), contigs, is_dam=True)
    dazzdb.write_db(os.path.join(wd, "reads.db"), (rb, rl))
    want = _streams(run_ref.run_damapper(wd, "ref.dam", "reads.db", flags=("-M16",) + tuple(flags), threads=4))
    got = _streams(run_ref.run_damapper(wd, "ref.dam", "reads.db", flags=("-M16",) + tuple(flags), threads=4, exe=gpu))
    assert len(want[0]) + len(want[1]) > 1000
    assert got[0] == want[0], "M records differ"
    assert got[1] == want[1], "R records differ"
    assert got[2] == want[2], "-p track differs"


def _three_block_case(wd, seed=61, nblocks_ref=1):
    from damapper_b200 import dazzdb, synth
    contigs, rb, rl = synth.make_config("C1", scale=0.06, seed=seed)
    dazzdb.write_db(os.path.join(wd, "ref.dam"), contigs, is_dam=True, nblocks=nblocks_ref)
    w = dazzdb.StreamDBWriter(os.path.join(wd, "reads.db"))
    off = np.concatenate([[0], np.cumsum(rl)])
    cuts = [0, len(rl) // 3, 2 * len(rl) // 3, len(rl)]
    for a, b in zip(cuts[:-1], cuts[1:]):
        w.append(rb[off[a]:off[b]], rl[a:b])
    w.close(nblocks=3)
    return contigs, rb, rl, cuts


def _run_blocks(wd, tag, exe, flags, env_extra=None, blocks=(1, 2, 3)):
    import subprocess
    from damapper_b200 import las
    from oracle import run_ref
    keep = os.path.join(wd, "keep_" + tag); os.makedirs(keep)
    tmp = os.path.join(wd, "tmp_" + tag); os.makedirs(tmp)
    env = dict(os.environ, DAMAPPER_KEEP_DIR=keep, **(env_extra or {}))
    env["PATH"] = os.path.join(run_ref.REF_DIR, "bin") + os.pathsep + env["PATH"]
    p = subprocess.run([exe, "-T4", "-P" + tmp, "-M16"] + list(flags) + ["ref.dam"] + ["reads.%d" % b for b in blocks],
                       cwd=wd, env=env, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stderr
    out = {}
    for b in blocks:
        m = run_ref._thread_sorted(glob.glob(os.path.join(keep, "reads.%d.ref.M[0-9]*.las" % b)))
        r = run_ref._thread_sorted(glob.glob(os.path.join(keep, "ref.reads.%d.R[0-9]*.las" % b)))
        prof = os.path.join(wd, ".reads.%d.prof.data" % b)
        pdata = b""
        if os.path.exists(prof):
            pdata = open(prof, "rb").read()
            os.remove(prof)
        out[b] = (las.canonical_stream(m) if m else b"", las.canonical_stream(r) if r else b"", pdata)
    return out, p


def test_multi_gpu_workers_give_the_single_device_stream(tmp_path):
    """damapper -G2: two worker processes, reads blocks round robin (damapper.c:825-914 is a loop of
    independent blocks).  Both workers are put on device 0 here (DAMGPU_DEVICES=0,0) so that the test runs on
    a one-GPU box; with two devices visible the same command line uses both (tools/gpu_multi.py)."""
    from oracle import run_ref
    if not run_ref.have_ref():
        pytest.skip("stubs of oracle/_ref/bin are not built")
    wd = str(tmp_path)
    _three_block_case(wd)
    one, _ = _run_blocks(wd, "g1", EXE, ["-C", "-p"])
    two, p = _run_blocks(wd, "g2", EXE, ["-C", "-p", "-G2"], {"DAMGPU_DEVICES": "0,0"})
    assert sum(len(v[0]) for v in one.values()) > 10000
    assert all(len(v[2]) > 0 for v in one.values())
    assert one == two
    # more workers than blocks, and a device list that is too short
    four, _ = _run_blocks(wd, "g4", EXE, ["-C", "-G8"], {"DAMGPU_DEVICES": "0,0,0"})
    assert {b: v[:2] for b, v in four.items()} == {b: v[:2] for b, v in one.items()}
    import subprocess
    q = subprocess.run([EXE, "-G3", "ref.dam", "reads.1", "reads.2", "reads.3"], cwd=wd,
                       env=dict(os.environ, DAMGPU_DEVICES="0"), capture_output=True, text=True)
    assert q.returncode == 1 and "only 1 devices" in q.stderr


def test_block_level_mask_tracks(tmp_path):
    """-m with tracks stored per block (.<root>.<block>.<track>.anno/.data, what DBdust etc. leave when run
    on blocks) against the same intervals stored as one whole-DB track, and against the reference."""
    from damapper_b200 import dazzdb
    from oracle import run_ref
    if not run_ref.have_ref():
        pytest.skip("oracle/_ref is not built")
    wd = str(tmp_path)
    contigs, rb, rl, cuts = _three_block_case(wd, seed=62)
    off, pts = dazzdb.random_masks(rl, seed=5, max_intervals=4, max_len=600)
    dazzdb.write_mask_track(os.path.join(wd, "reads.db"), "dust", off, pts)
    whole, _ = _run_blocks(wd, "whole", EXE, ["-C", "-mdust"])
    ref, _ = _run_blocks(wd, "ref", run_ref.REF_BIN, ["-C", "-mdust"])
    assert whole == ref
    nomask, _ = _run_blocks(wd, "nomask", EXE, ["-C"])
    assert nomask != whole, "the mask changed nothing: the case does not test it"
    for f in glob.glob(os.path.join(wd, ".reads.dust.*")):
        os.remove(f)
    for b, (a, e) in enumerate(zip(cuts[:-1], cuts[1:]), start=1):
        o = off[a:e + 1] - off[a]
        dazzdb.write_mask_track(os.path.join(wd, "reads.%d" % b), "dust", o, pts[off[a]:off[e]])
    per_block, p = _run_blocks(wd, "blocks", EXE, ["-C", "-mdust"])
    assert "never used" not in p.stdout
    assert per_block == whole


def test_trace_point_verifier_rejects_a_corrupted_record(tmp_path):
    """k_check_trace (Check_Trace_Points, align.c:3194-3236, on the device): with one trace byte of the
    first record changed before the check (DAMGPU_TEST_CORRUPT_TRACE, a test hook in report.cu) the
    Reporter must stop through Clean_Exit; without it the same run succeeds."""
    import subprocess
    from damapper_b200 import dazzdb, synth
    contigs, rb, rl = synth.make_config("C1", scale=0.03, seed=63)
    wd = str(tmp_path)
    dazzdb.write_db(os.path.join(wd, "ref.dam"), contigs, is_dam=True)
    dazzdb.write_db(os.path.join(wd, "reads.db"), (rb, rl))
    os.makedirs(os.path.join(wd, "tmp"))
    cmd = [EXE, "-T2", "-P" + os.path.join(wd, "tmp"), "ref.dam", "reads.db"]
    ok = subprocess.run(cmd, cwd=wd, env=dict(os.environ, DAMGPU_BUILTIN_SORT="1"), capture_output=True, text=True, timeout=600)
    assert ok.returncode == 0, ok.stderr
    assert os.path.getsize(os.path.join(wd, "reads.ref.las")) > 1000
    os.remove(os.path.join(wd, "reads.ref.las"))
    bad = subprocess.run(cmd, cwd=wd, env=dict(os.environ, DAMGPU_BUILTIN_SORT="1", DAMGPU_TEST_CORRUPT_TRACE="1"),
                         capture_output=True, text=True, timeout=600)
    assert bad.returncode == 1
    assert "fail Check_Trace_Points" in bad.stderr
    assert not os.path.exists(os.path.join(wd, "reads.ref.las"))
    assert glob.glob(os.path.join(wd, "tmp", "damapper.*")) == [], "Clean_Exit must remove the sort directory"


def test_layer1_match_filter_with_an_empty_side(tmp_path):
    """map.c:2955-2956: Match_Filter returns at once when either list is empty, and the run goes on to
    write (empty) .las files.  Layer 1 must do the same: a reference block without k-mers, then Reporter."""
    import ctypes as C
    import struct
    from damapper_b200 import api, dazzdb, synth
    from conftest import install_fatal_hook
    L = api.init(0)
    install_fatal_hook(api)
    contigs, rb, rl = synth.make_config("C1", scale=0.01, seed=64)
    hr = api.HostBlock(*dazzdb.load_block((rb, rl)))
    empty = api.HostBlock(np.array([4], dtype=np.uint8), np.zeros(1, dtype=np.int64), np.zeros(0, dtype=np.int32))
    hw = api.HostBlock(*dazzdb.load_block(contigs))
    api.set_filter_params(20, 0, 4)
    api.set_options(sort_path=str(tmp_path), verbose=1)   # -v: the closing lines are printed on this path too
    spec = api.CAlignSpec(0.85, 100, (C.c_float * 4)(.25, .25, .25, .25))
    blen, alen = C.c_int(0), C.c_int(0)
    bindex = L.damgpu_Sort_Kmers(C.byref(hr.c), C.byref(blen))
    aindex = L.damgpu_Sort_Kmers(C.byref(empty.c), C.byref(alen))
    assert blen.value > 0 and alen.value == 0 and not aindex
    L.damgpu_Match_Filter(C.byref(hr.c), C.byref(empty.c), bindex, blen, aindex, alen, 0, 1)
    L.damgpu_Match_Filter(C.byref(hr.c), C.byref(empty.c), bindex, blen, aindex, alen, 1, 0)
    L.damgpu_Reporter(b"reads", C.byref(hr.c), b"ref", C.byref(hw.c), C.byref(spec), 3)
    L.damgpu_index_free(bindex)
    for i in range(1, 5):
        for name in ("reads.ref.M%d.las" % i, "ref.reads.R%d.las" % i):
            data = open(os.path.join(str(tmp_path), name), "rb").read()
            assert len(data) == 12 and struct.unpack("<qi", data) == (0, 100)
    api.set_options()


@pytest.mark.parametrize("flags", [("-M16",), ("-M0", "-C", "-t20"), ("-M1", "-C", "-p", "-k16")])
def test_verbose_statistics_match_the_reference(tmp_path, flags):
    """-v: what Sort_Kmers, Match_Filter and Reporter tell the user (map.c:692-697,792-814,2990-3071,
    3185-3208,3289-3293) -- k-mer counts, index sizes, the hit cap, hit counts, candidates added / removed
    per call and in total, mapped segments, and the block-size warning on stderr -- must read exactly as the
    reference's, from the shipped driver and from the reference's own driver linked against libdamgpu."""
    import re
    from damapper_b200 import dazzdb, synth
    from oracle import run_ref
    if not run_ref.have_ref():
        pytest.skip("oracle/_ref/damapper is not built")
    contigs, rb, rl = synth.make_config("C1", scale=0.05, seed=66)
    wd = str(tmp_path)
    dazzdb.write_db(os.path.join(wd, "ref.dam"), contigs, is_dam=True)
    dazzdb.write_db(os.path.join(wd, "reads.db"), (rb, rl))

    def said(exe):
        r = run_ref.run_damapper(wd, "ref.dam", "reads.db", flags=("-v",) + tuple(flags), threads=4, exe=exe)
        for f in ("prof_anno", "prof_data"):
            if r[f]:
                os.remove(r[f])
        return re.sub(r"damapper\.\d+", "damapper.PID", r["stdout"]), r["stderr"]

    want = said(None)
    assert "Kmer count = " in want[0] and "candidates added" in want[0] and "mapped segments" in want[0]
    assert said(EXE) == want
    gpu = os.path.join(run_ref.REF_DIR, "damapper_gpu")
    if os.access(gpu, os.X_OK):
        assert said(gpu) == want
