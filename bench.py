#!/usr/bin/env python
"""bench.py -- mapped read bases/sec of the B200 mapping core on BASELINE.json's metric config.

Workload (config.workload): BASELINE config 2 -- synthetic 4.6 Mbp reference (2 contigs) and
30x coverage of simulated 10 kbp reads at 15 % error (13 800 reads, ~138 Mbp), k=20, defaults.
A step = the whole hot path for one reads block: Sort_Kmers(reads), then per reference block
Sort_Kmers + Match_Filter for both orientations, then Reporter (chain extension, selection,
records copied back).  N > 1 (torchrun, one rank per GPU): every rank maps its own reads
block (weak scaling); the forward and complement reference indices are built on rank 0 and
broadcast over NCCL, reads need no collective.

  value  inputs resident in HBM when the timed region starts (blocks already uploaded)
  e2e    through the reference-facing C ABI (the four map.h calls) with host buffers in pinned
         memory: H2D of the blocks and D2H of the records are inside the timed region

--impl reference times the unmodified reference damapper (oracle/_ref, built from
/root/reference by oracle/Makefile) on the host cores, on a bounded sample of the same
workload; the oracle directory is touched only there and in the cpu_baseline leg.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "C2: synthetic 4.6 Mbp reference (2 contigs) + 13,800 simulated 10 kbp reads at 15% error (30x), k=20"
METRIC = "mapped read bases/sec"
UNIT = "bases/s"


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = threading.Event()
        self.samples = []

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True,
                                     text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.05)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples for i in range(4)
                          if len(s) > 2 + i and s[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.samples)}


def make_workload(seed: int, reads_scale: float = 1.0):
    from damapper_b200 import synth, dazzdb
    contigs, rb, rl = synth.make_config("C2", scale=1.0, seed=7)     # the reference is shared
    if seed != 7 or reads_scale != 1.0:
        genome = np.concatenate(contigs)
        cuts = np.array([0, contigs[0].size, genome.size])
        n = max(8, int(13800 * reads_scale))
        rb, rl, _ = synth.make_reads(genome, n, seed=seed + 1, contig_bounds=cuts)
    cnt = np.bincount(np.concatenate(contigs), minlength=4).astype(np.float64)
    freq = tuple(float(x) for x in (cnt / cnt.sum()).astype(np.float32))
    return contigs, rb, rl, freq


def pinned_block(api, loaded, torch):
    """HostBlock whose base array lives in pinned host memory."""
    bases, boff, rlen = loaded
    t = torch.empty(bases.size, dtype=torch.uint8, pin_memory=True)
    t.numpy()[:] = bases
    hb = api.HostBlock(t.numpy(), boff, rlen)
    hb._pin = t
    return hb


# ---------------------------------------------------------------------------- reference arm

def run_reference(args):
    """Unmodified reference damapper -T<host cores> on a bounded sample of the workload."""
    from damapper_b200 import dazzdb
    from oracle import run_ref                       # test infrastructure, CPU arm only
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if not run_ref.have_ref():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/damapper is not built"}))
        return
    cores = os.cpu_count() or 1
    threads = 1
    while 2 * threads <= cores:
        threads *= 2                                 # the reference rounds -T down to 2^n (map.c:142-147)
    contigs, rb, rl, freq = make_workload(seed=7)
    sample_reads = len(rl)                           # the whole workload: ~1.2 s per run on 16 threads (~19 CPU-seconds)
    off = np.concatenate([[0], np.cumsum(rl)])
    rb_s, rl_s = rb[:off[sample_reads]], rl[:sample_reads]
    bases = int(rl_s.sum())
    wd = tempfile.mkdtemp(prefix="bench_ref_")
    times = []
    try:
        dazzdb.write_db(os.path.join(wd, "ref.dam"), contigs, is_dam=True)
        dazzdb.write_db(os.path.join(wd, "reads.db"), (rb_s, rl_s))
        for it in range(args.warmup + args.steps):
            r = run_ref.run_damapper(wd, "ref.dam", "reads.db", flags=("-M64",), threads=threads)
            if it >= args.warmup:
                times.append(r["wall_s"])
    finally:
        shutil.rmtree(wd, ignore_errors=True)
    t = float(np.mean(times))
    val = bases / t
    sample = ("the whole workload (%d reads, %d bases) against the full 4.6 Mbp reference, damapper -T%d -M64, "
              "whole process wall clock incl. DB load and .las writes, LAsort/LAcat stubbed" % (sample_reads, bases, threads))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def cpu_baseline():
    """Reference damapper on the host cores, bounded sample (rank 0, N=1 only)."""
    from damapper_b200 import dazzdb
    from oracle import run_ref
    if not run_ref.have_ref():
        return None
    cores = os.cpu_count() or 1
    threads = 1
    while 2 * threads <= cores:
        threads *= 2
    contigs, rb, rl, freq = make_workload(seed=7)
    sample_reads = len(rl)                           # the whole workload (~19 CPU-seconds per run)
    off = np.concatenate([[0], np.cumsum(rl)])
    rb_s, rl_s = rb[:off[sample_reads]], rl[:sample_reads]
    wd = tempfile.mkdtemp(prefix="bench_cpu_")
    try:
        dazzdb.write_db(os.path.join(wd, "ref.dam"), contigs, is_dam=True)
        dazzdb.write_db(os.path.join(wd, "reads.db"), (rb_s, rl_s))
        best = None
        for _ in range(2):
            r = run_ref.run_damapper(wd, "ref.dam", "reads.db", flags=("-M64",), threads=threads)
            best = r["wall_s"] if best is None else min(best, r["wall_s"])
    finally:
        shutil.rmtree(wd, ignore_errors=True)
    bases = int(rl_s.sum())
    return {"value": bases / best, "unit": UNIT, "cores": threads, "kind": "reference",
            "sample": "reference damapper -T%d -M64 on the whole workload (%d reads, %d bases) vs the full 4.6 Mbp "
                      "reference, best of 2, whole process" % (threads, sample_reads, bases)}


# ---------------------------------------------------------------------------------- GPU arm

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="damgpu")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    if args.impl == "reference":
        run_reference(args)
        return

    # only the JSON line goes to the real stdout (NCCL and torchrun banners go to stderr)
    real_out = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    from damapper_b200 import api, dazzdb, shard

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the mapping core has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = api.init(local)
    args.warmup = max(args.warmup, 3)

    # every rank maps its own reads block (different seed) against the same reference
    contigs, rb, rl, freq = make_workload(seed=7 + 100 * rank)
    rd = dazzdb.load_block((rb, rl))
    rf = dazzdb.load_block(contigs)
    rc = dazzdb.load_block(dazzdb.revcomp_contigs(contigs))
    hr, hg, hc = pinned_block(api, rd, torch), pinned_block(api, rf, torch), pinned_block(api, rc, torch)
    bases = int(rl.sum())
    api.set_filter_params(20, 0, 4)
    api.set_options(mem_limit=64 << 30)
    spec = api.CAlignSpec(0.85, 100, (C.c_float * 4)(*freq))

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def ref_index(dref):
        """Reference index of the block's current orientation: built on rank 0, broadcast."""
        if world == 1:
            return api.Index(dref)
        idx, payload = None, None
        if rank == 0:
            idx = api.Index(dref)
            payload = torch.empty((len(idx) + 2) * 16, dtype=torch.uint8, device="cuda")
            L.damgpu_index_export(idx.h, payload.data_ptr())
        buf, ln = shard.broadcast_index(dist, torch, payload, "cuda", src=0)
        if rank == 0:
            return idx
        torch.cuda.synchronize()
        return api.Index(handle=L.damgpu_index_import(buf.data_ptr(), ln))

    stats = {}

    def step_resident(dr, dref_f):
        """One step with the blocks already in HBM.  dref_f is left in forward orientation."""
        ir = api.Index(dr, deferred=True)         # Sort_Kmers(reads): built per reference block, filtered
        stats["kmers"] = len(ir)
        m = api.Mapper(dr, ir)
        ig = ref_index(dref_f)
        stats["ref_sort"] = api.last_sort_times(); stats["ref_kmers"] = len(ig)
        m.match(dref_f, ig, 0, 1)
        stats["filter"] = api.last_filter_times()
        stats["hits_fwd"] = m.last_hits
        stats["join_fwd"] = api.last_join_times()
        ig.free()
        dref_f.complement()
        ig = ref_index(dref_f)
        m.match(dref_f, ig, 1, 0)
        stats["join_rc"] = api.last_join_times()
        ig.free()
        dref_f.complement()                       # back to forward: Reporter wants the plain reference
        rep = m.report(dref_f, 0.85, 100, freq, 1)
        nrec = rep.records(0)
        nbytes = L.damgpu_report_bytes(rep.h, 0)
        stats["report"] = rep.stats()
        rep.free(); m.free(); ir.free()
        return nrec, nbytes

    def step_e2e(tmpdir):
        """The four map.h calls with host buffers (H2D and D2H inside)."""
        blen, alen = C.c_int(0), C.c_int(0)
        bindex = L.damgpu_Sort_Kmers(C.byref(hr.c), C.byref(blen))
        aindex = L.damgpu_Sort_Kmers(C.byref(hg.c), C.byref(alen))
        L.damgpu_Match_Filter(C.byref(hr.c), C.byref(hg.c), bindex, blen, aindex, alen, 0, 1)
        aindex = L.damgpu_Sort_Kmers(C.byref(hc.c), C.byref(alen))
        L.damgpu_Match_Filter(C.byref(hr.c), C.byref(hc.c), bindex, blen, aindex, alen, 1, 0)
        L.damgpu_Reporter(b"reads", C.byref(hr.c), b"ref", C.byref(hg.c), C.byref(spec), 1)
        L.damgpu_index_free(bindex)

    # ---- value: resident inputs
    dr, dg = api.DeviceBlock(hr), api.DeviceBlock(hg)
    L.damgpu_time_kernels(1)
    for _ in range(args.warmup):
        nrec, nbytes = step_resident(dr, dg)
    sampler = ClockSampler(local)
    sync_all()
    sampler.start()
    l0 = L.damgpu_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    t0 = time.perf_counter()
    sort_ms, sort_n, ext_ms, aln_ms, join_ms, flt, rsort = [], 0, [], [], [], [], []
    for _ in range(args.steps):
        nrec, nbytes = step_resident(dr, dg)
        flt.append(stats["filter"]); rsort.append(stats["ref_sort"])
        aln_ms.append(stats["report"]["align_ms"])
        join_ms.append(stats["join_fwd"]["lut_ms"] + stats["join_fwd"]["match_ms"] +
                       stats["join_rc"]["lut_ms"] + stats["join_rc"]["match_ms"])
    ev1.record()
    sync_all()
    t1 = time.perf_counter()
    launches = L.damgpu_launch_count() - l0
    dev_ms = ev0.elapsed_time(ev1)
    wall_ms = (t1 - t0) * 1e3
    step_ms = max(dev_ms, wall_ms) / args.steps          # host-side syncs are part of the step
    if world > 1:
        t = torch.tensor([step_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        step_ms = float(t.item())
        b = torch.tensor([bases], dtype=torch.float64, device="cuda")
        dist.all_reduce(b, op=dist.ReduceOp.SUM)
        total_bases = float(b.item())
    else:
        total_bases = float(bases)
    value = total_bases / (step_ms / 1e3)

    # ---- roofline kernel: k_radix_pass on the WHOLE reads list (Sort_Kmers as the reference does it:
    # the form taken with -t, -m, more than two reference blocks, or when the list is asked for), timed
    # live with CUDA events on the library's stream, the list (2.2 GB) far larger than L2
    if rank == 0:
        for it in range(2 + args.steps):
            ir = api.Index(dr)
            st = api.last_sort_times()
            if it >= 2:
                sort_ms.append(st["sort_ms"]); ext_ms.append(st["extract_ms"]); sort_n = st["npass"]
            ir.free()
    dg.free(); dr.free()
    L.damgpu_time_kernels(0)

    # ---- e2e: host buffers through the map.h-shaped C ABI
    tmpdir = tempfile.mkdtemp(prefix="bench_e2e_")
    api.set_options(mem_limit=64 << 30, sort_path=tmpdir)
    try:
        for _ in range(2):
            step_e2e(tmpdir)
        sync_all()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_e2e(tmpdir)
        sync_all()
        e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
        sampler.stop_flag.set()                          # clocks were sampled over both timed regions
        sampler.join()
    finally:
        sampler.stop_flag.set()
        shutil.rmtree(tmpdir, ignore_errors=True)
    if world > 1:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    h2d = int(hr.bases.size + 2 * hg.bases.size + hg.bases.size +
              8 * (hr.boff.size + 3 * hg.boff.size) + 4 * (hr.rlen.size + 3 * hg.rlen.size))
    d2h = int(nbytes + 16 * hr.nreads)

    if rank == 0:
        peak, which = peaks()
        n = stats["kmers"]
        pass_ms = float(np.mean(sort_ms)) / max(sort_n, 1)
        achieved = 32.0 * n / (pass_ms / 1e3) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "r01_radix_pass_traffic.json")) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        except Exception:
            pass
        rs = stats["report"]
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "reads_per_gpu": int(hr.nreads), "read_bases_per_gpu": bases,
                       "kmer": 20, "l2": "inputs larger than L2 (126 MB): the reads block is %.0f MB (1 byte per base), the whole k-mer "
                             "list sorted in the roofline region %.1f GB; no flush between steps" % (bases / 1e6, 16.0 * n / 1e9),
                       "index": "built on rank 0 and broadcast with NCCL" if world > 1 else "built locally"},
            "e2e": {"value": total_bases / (e2e_ms / 1e3), "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "roofline": {"kernel": "k_radix_pass (one LSD pass over the whole reads k-mer list)", "bound": "hbm",
                         "note": "timed in its own region after the steps: a step builds the reads list only "
                                 "from the k-mers that occur in the reference block (roofline_filter), the "
                                 "same pass kernel then runs on %d instead of %d records"
                                 % (int(np.mean([f["survivors"] for f in flt])), n),
                         "achieved": achieved, "peak": peak, "peak_source": which + " copy bandwidth, burst",
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "algorithmic_bytes_per_launch": 32 * n, "avg_launch_ms": pass_ms,
                         "launches_per_step": sort_n},
            # merge-join of both orientations of a step: SURVEY 8(d) counts 16 B per input record per
            # scan; here the shorter list drives and finds its codes in the longer one through a prefix
            # table (built once per reads block), so the second call never scans the reads list
            "roofline_merge": {"kernel": "k_build_lut + k_join_match, both orientations of a step", "bound": "hbm",
                               "algorithmic_bytes": 2 * 16 * (stats["join_fwd"]["alen"] + stats["join_fwd"]["blen"]),
                               "ms": float(np.mean(join_ms)),
                               "achieved": 2 * 16.0 * (stats["join_fwd"]["alen"] + stats["join_fwd"]["blen"])
                                           / (max(float(np.mean(join_ms)), 1e-9) / 1e3) / 1e9,
                               "peak": peak, "unit": "GB/s",
                               "frac": 2 * 16.0 * (stats["join_fwd"]["alen"] + stats["join_fwd"]["blen"])
                                       / (max(float(np.mean(join_ms)), 1e-9) / 1e3) / 1e9 / peak},
            # the reads-side index of a step: hash bitmap of the reference codes (both orientations),
            # extraction with a membership test + ordered compaction, radix passes over the survivors.
            # Bound by bitmap lookups in L2 (one 32-byte sector per k-mer), not by HBM: algorithmic bytes
            # = 1 B per base + 16 B per survivor.
            "roofline_filter": {"kernel": "k_extract_filtered", "bound": "l2 lookups",
                                "ms": float(np.mean([f["extract_ms"] for f in flt])),
                                "lookups_per_s": n / (max(float(np.mean([f["extract_ms"] for f in flt])), 1e-9) / 1e3),
                                "survivors": int(np.mean([f["survivors"] for f in flt])), "kmers": n,
                                "algorithmic_bytes": int(bases + hr.nreads + 16 * np.mean([f["survivors"] for f in flt]))},
            # the same pass kernel where a step runs it on rank 0: the forward reference list of the step
            # (launch-bound at this size: a pass is ~25 us), timed live inside the timed region
            "roofline_in_step": {"kernel": "k_radix_pass on the forward reference list inside the timed steps",
                                 "bound": "hbm", "records": int(stats["ref_kmers"]),
                                 "avg_launch_ms": float(np.mean([r["sort_ms"] / max(r["npass"], 1) for r in rsort])),
                                 "achieved": 32.0 * stats["ref_kmers"] / (max(float(np.mean([r["sort_ms"] / max(r["npass"], 1) for r in rsort])), 1e-9) / 1e3) / 1e9,
                                 "peak": peak, "unit": "GB/s",
                                 "frac": 32.0 * stats["ref_kmers"] / (max(float(np.mean([r["sort_ms"] / max(r["npass"], 1) for r in rsort])), 1e-9) / 1e3) / 1e9 / peak},
            "phases_ms": {"ref_bitmap": float(np.mean([f["bitmap_ms"] for f in flt])),
                          "extract_filtered": float(np.mean([f["extract_ms"] for f in flt])),
                          "radix_sort_survivors": float(np.mean([f["sort_ms"] for f in flt])),
                          "align_kernel": float(np.mean(aln_ms)),
                          "full_sort_extract": float(np.mean(ext_ms)), "full_sort_radix": float(np.mean(sort_ms))},
            "extension": {"cells_per_s": rs["ncells"] / (max(float(np.mean(aln_ms)), 1e-9) / 1e3),
                          "waves": rs["nwaves"], "alignments": rs["nalign"], "records": int(nrec)},
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                out["cpu_baseline"] = cpu_baseline()
            except Exception as e:                       # the baseline must not take the line down
                out["cpu_baseline"] = {"error": str(e)[:200]}
        os.write(real_out, (json.dumps(out) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
