#!/usr/bin/env python
"""bench.py -- mapped read bases/sec of the B200 mapping core on BASELINE.json's metric config.

Workload (config.workload): BASELINE config 2 -- synthetic 4.6 Mbp reference (2 contigs) and
30x coverage of simulated 10 kbp reads at 15 % error (13 800 reads, ~140 Mbp), k=20, defaults.
A step = the whole hot path for one reads block: Sort_Kmers(reads), then per reference block
Sort_Kmers + Match_Filter for both orientations, then Reporter (chain extension, selection,
records copied back).  N > 1 (torchrun, one rank per GPU): every rank maps its own reads
block (weak scaling); reads need no collective.  The reference index (2 x 74 MB here) is built by
every rank for itself: measured on 2 x B200, the local build is never slower than an NCCL broadcast
or a sharded build + all-gather, at 4.6 Mbp and at 250 Mbp; --index broadcast
keeps the NCCL path measurable.  Both arms keep DB files and -P sort directories in one scratch
directory (config.scratch: /dev/shm when it is a tmpfs with room).

  value  inputs resident in HBM when the timed region starts (blocks already uploaded)
  e2e    through the reference-facing C ABI (the four map.h calls) with host buffers in pinned
         memory: H2D of the blocks and D2H of the records are inside the timed region.  The reads
         block is handed over as the .bps image (2 bits per base, damgpu_block.packed) -- the form
         the DB holds on disk -- the reference blocks as Load_All_Reads leaves them.
  whole_process  exec -> exit of the two command lines on the same DB files:
         damapper_b200/damapper and oracle/_ref/damapper, LAsort/LAcat stubbed for both.

--impl reference times the unmodified reference damapper (oracle/_ref, built from /root/reference
by oracle/Makefile) on the host cores on the whole workload, whole process; the oracle directory
is touched only there and in the cpu_baseline / whole_process legs.
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


def profile_json(name):
    try:
        with open(os.path.join(ROOT, "profiles", name)) as f:
            return json.load(f)
    except Exception:
        return {}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = threading.Event()
        self.samples = []

    def run(self):
        if os.environ.get("BENCH_NO_SMI"):
            return
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
            self.stop_flag.wait(0.2)                      # the recipe's -lms 200

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
    from damapper_b200 import synth
    contigs, rb, rl = synth.make_config("C2", scale=1.0, seed=7)     # the reference is shared
    if seed != 7 or reads_scale != 1.0:
        genome = np.concatenate(contigs)
        cuts = np.array([0, contigs[0].size, genome.size])
        n = max(8, int(13800 * reads_scale))
        rb, rl, _ = synth.make_reads(genome, n, seed=seed + 1, contig_bounds=cuts)
    cnt = np.bincount(np.concatenate(contigs), minlength=4).astype(np.float64)
    freq = tuple(float(x) for x in (cnt / cnt.sum()).astype(np.float32))
    return contigs, rb, rl, freq


def scratch_base():
    """Where both arms keep their DB files and -P sort directories: BENCH_SCRATCH, else /dev/shm when it is a
    writable tmpfs with room (what -P is for: the root volume of these boxes is ext4 mounted with `discard`,
    on which unlinking / truncating a few MB costs milliseconds), else the default temporary directory."""
    e = os.environ.get("BENCH_SCRATCH")
    if e:
        return e if os.path.isdir(e) and os.access(e, os.W_OK) else None
    d = "/dev/shm"
    try:
        if os.path.isdir(d) and os.access(d, os.W_OK):
            st = os.statvfs(d)
            if st.f_bavail * st.f_frsize >= (4 << 30):
                return d
    except OSError:
        pass
    return None


def host_threads():
    cores = os.cpu_count() or 1
    threads = 1
    while 2 * threads <= cores:
        threads *= 2                                 # the reference rounds -T down to 2^n (map.c:142-147)
    return threads


def config_of(nreads, bases, n_kmers, index):
    """The same keys from both arms (the driver compares them)."""
    return {"workload": WORKLOAD, "reads_per_gpu": int(nreads), "read_bases_per_gpu": int(bases), "kmer": 20,
            "l2": "inputs larger than L2 (126 MB): the reads block is %.0f MB at one byte per base (35 MB packed), its "
                  "k-mer list %.1f GB; no flush between steps" % (bases / 1e6, 16.0 * n_kmers / 1e9),
            "index": index,
            "scratch": "%s for both arms (DB files, -P sort directories)" % (scratch_base() or tempfile.gettempdir())}


# ------------------------------------------------------------------ the two command lines

def run_cli(wd, exe, threads, env_extra=None):
    """exec -> exit of a damapper command line in wd (stubs for LAsort/LAcat/LAmerge on PATH)."""
    from oracle import run_ref                       # test infrastructure: CPU legs only
    r = run_ref.run_damapper(wd, "ref.dam", "reads.db", flags=("-M64",), threads=threads, exe=exe,
                             env_extra=env_extra)
    return r


def whole_process(contigs, rb, rl, device, runs=3):
    """whole process, both command lines, same DB files, same stubs.  Returns
    (dict for the JSON line, cpu_baseline dict)."""
    from damapper_b200 import dazzdb
    from oracle import run_ref
    if not run_ref.have_ref():
        return None, None
    threads = host_threads()
    bases = int(rl.sum())
    wd = tempfile.mkdtemp(prefix="bench_wp_", dir=scratch_base())
    gpu_exe = os.path.join(ROOT, "damapper_b200", "damapper")
    timed = os.path.join(run_ref.REF_DIR, "damapper_timed")
    ref_w, gpu_w, core = [], [], None
    try:
        dazzdb.write_db(os.path.join(wd, "ref.dam"), contigs, is_dam=True)
        dazzdb.write_db(os.path.join(wd, "reads.db"), (rb, rl))
        for it in range(runs):
            ref_w.append(run_cli(wd, None, threads)["wall_s"])
            if os.access(gpu_exe, os.X_OK):
                gpu_w.append(run_cli(wd, gpu_exe, threads, {"DAMGPU_DEVICE": str(device)})["wall_s"])
        if os.access(timed, os.X_OK):
            core = run_cli(wd, timed, threads)["core_s"]
    finally:
        shutil.rmtree(wd, ignore_errors=True)
    sample = ("reference damapper -T%d -M64 on the whole workload (%d reads, %d bases) vs the full 4.6 Mbp reference, "
              "whole process (exec -> exit, DB load and .las writes, LAsort/LAcat stubbed), median of %d runs"
              % (threads, len(rl), bases, runs))
    wp = {"boundary": "exec -> exit of the command line on the same DB files, LAsort/LAcat/LAmerge stubbed for both",
          "reference": {"wall_s_median": float(np.median(ref_w)), "wall_s_min": float(min(ref_w)), "runs_s": ref_w,
                        "threads": threads, "value": bases / float(np.median(ref_w)), "unit": UNIT,
                        "core_only_s": core,
                        "core_only_note": "sum of the Sort_Kmers/Match_Filter/Reporter calls (damapper.c:833-875), "
                                          "clocked by ld --wrap around the unmodified sources (oracle/ref_timing.c)"}}
    if gpu_w:
        wp["gpu"] = {"wall_s_median": float(np.median(gpu_w)), "wall_s_min": float(min(gpu_w)), "runs_s": gpu_w,
                     "value": bases / float(np.median(gpu_w)), "unit": UNIT,
                     "note": "a fresh process pays cuInit + context creation (0.6-3 s on these boxes, DESIGN section 7) "
                             "before ~60 ms of work; e2e is the same path in a warm process"}
    cpu = {"value": bases / float(np.median(ref_w)), "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample}
    return wp, cpu


# ---------------------------------------------------------------------------- reference arm

def run_reference(args):
    """Unmodified reference damapper -T<host cores> on the whole workload, whole process."""
    from damapper_b200 import dazzdb
    from oracle import run_ref                       # test infrastructure, CPU arm only
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if not run_ref.have_ref():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/damapper is not built"}))
        return
    threads = host_threads()
    contigs, rb, rl, freq = make_workload(seed=7)
    bases = int(rl.sum())
    n_kmers = int((rl - 19).sum())
    wd = tempfile.mkdtemp(prefix="bench_ref_", dir=scratch_base())
    times, core = [], None
    try:
        dazzdb.write_db(os.path.join(wd, "ref.dam"), contigs, is_dam=True)
        dazzdb.write_db(os.path.join(wd, "reads.db"), (rb, rl))
        for it in range(args.warmup + args.steps):
            r = run_cli(wd, None, threads)
            if it >= args.warmup:
                times.append(r["wall_s"])
        timed = os.path.join(run_ref.REF_DIR, "damapper_timed")
        if os.access(timed, os.X_OK):
            core = run_cli(wd, timed, threads)["core_s"]
    finally:
        shutil.rmtree(wd, ignore_errors=True)
    t = float(np.mean(times))
    val = bases / t
    sample = ("the whole workload (%d reads, %d bases) against the full 4.6 Mbp reference, damapper -T%d -M64, "
              "whole process wall clock incl. DB load and .las writes, LAsort/LAcat stubbed" % (len(rl), bases, threads))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": config_of(len(rl), bases, n_kmers, "built locally"),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "whole_process": {"reference": {"wall_s_mean": t, "threads": threads, "core_only_s": core}},
    }))


# ---------------------------------------------------------------------------------- GPU arm

def pinned_array(torch, a):
    t = torch.empty(a.size, dtype=torch.uint8, pin_memory=True)
    t.numpy()[:] = a.reshape(-1).view(np.uint8)
    return t


def pinned_block(api, loaded, torch, packed=False):
    """HostBlock whose base array (and .bps image) live in pinned host memory."""
    bases, boff, rlen = loaded
    t = pinned_array(torch, bases)
    hb = api.HostBlock(t.numpy(), boff, rlen)
    hb._pin = [t]
    if packed:
        pk, poff = api.pack_bps(hb)
        tp = pinned_array(torch, pk)
        hb.attach_packed(tp.numpy(), poff)
        hb._pin.append(tp)
    return hb


def run_c4(args):
    """BASELINE config 4, the one north_star names for sharding: 250 Mbp reference (8 contigs, one block of
    250 M k-mers per strand) + 200 000 simulated 10 kbp reads in 8 reads blocks of 25 000; rank r maps blocks
    r, r+N, ... (damapper.c:825-914: blocks are independent), STRONG scaling.  A step = the whole job of a rank:
    both reference indices built locally (at 4 GB per strand the local build costs what the NCCL
    transfer costs), then every block of the rank: packed H2D, reads index, Match_Filter on both strands against the
    resident indices, Reporter, records D2H -- i.e. the e2e boundary; `value` and `e2e` are the same measurement."""
    real_out = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    from damapper_b200 import api, dazzdb, synth
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = api.init(local)
    G, NB = 250_000_000, int(os.environ.get("C4_NB", "8"))
    per_block = max(8, int(25000 * args.c4_scale))
    genome = synth.make_genome(G, seed=11)
    cuts = np.array([int(G * i / 8) for i in range(9)])
    contigs = [genome[cuts[i]:cuts[i + 1]] for i in range(8)]
    cnt = np.bincount(genome[:10_000_000], minlength=4).astype(np.float64)
    freq = tuple(float(x) for x in (cnt / cnt.sum()).astype(np.float32))
    hg = pinned_block(api, dazzdb.load_block(contigs), torch, packed=True)
    mine = list(range(rank, NB, world))
    blocks, bases = [], 0
    for b in mine:
        rb, rl, _ = synth.make_reads(genome, per_block, seed=1000 + b, contig_bounds=cuts)
        blocks.append(pinned_block(api, dazzdb.load_block((rb, rl)), torch, packed=True))
        bases += int(rl.sum())
    del genome
    api.set_filter_params(20, 0, 4)
    api.set_options(mem_limit=64 << 30)
    nrec = [0]

    def job():
        tj = time.perf_counter()
        dg = api.DeviceBlock.from_host(hg)        # the reference too crosses PCIe at 2 bits per base
        igf = api.Index(dg)
        dg.complement()
        igr = api.Index(dg)                       # dg stays complemented: only its sizes are used by the matches
        dw = api.DeviceBlock.from_host(hg)        # the whole reference the Reporter aligns against
        nrec[0] = 0
        torch.cuda.synchronize(); tb = [time.perf_counter()]
        detail = os.environ.get("C4_TIMES") == "2"
        for hb in blocks:
            tc = [time.perf_counter()]
            dr = api.DeviceBlock.from_host(hb)             # 2 bits per base cross PCIe
            tc.append(time.perf_counter())
            ir = api.Index(dr, deferred=True)
            m = api.Mapper(dr, ir)
            tc.append(time.perf_counter())
            m.match(dg, igf, 0, 1)
            tc.append(time.perf_counter())
            m.match(dg, igr, 1, 0)
            tc.append(time.perf_counter())
            rep = m.report(dw, 0.85, 100, freq, 1)
            tc.append(time.perf_counter())
            nrec[0] += rep.records(0)
            _ = rep.a                             # the record stream comes back to the host
            tc.append(time.perf_counter())
            rep.free(); m.free(); ir.free(); dr.free()
            tc.append(time.perf_counter())
            if detail:
                print("[c4]   upload %.1f | index+mapper %.1f | match f %.1f | match r %.1f | report %.1f | bytes %.1f | free %.1f"
                      % tuple((b - a) * 1e3 for a, b in zip(tc[:-1], tc[1:])), file=sys.stderr, flush=True)
            tb.append(time.perf_counter())
        igf.free(); igr.free(); dg.free(); dw.free()
        if os.environ.get("C4_TIMES"):
            fr, to = C.c_uint64(0), C.c_uint64(0)
            L.damgpu_device_memory(C.byref(fr), C.byref(to))
            print("[c4] rank %d: setup %.1f ms, blocks %s ms; device memory not in use (free + cached) %.2f GB" % (
                  rank, (tb[0] - tj) * 1e3, " ".join("%.1f" % ((b - a) * 1e3) for a, b in zip(tb[:-1], tb[1:])),
                  fr.value / 1e9), file=sys.stderr, flush=True)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(); torch.cuda.synchronize()

    for _ in range(max(1, min(args.warmup, 2))):
        job()
    sampler = ClockSampler(local)
    sync_all()
    if rank == 0:
        sampler.start()
    l0 = L.damgpu_launch_count()
    per_step = []
    for _ in range(args.steps):                          # every job bracketed by a barrier + synchronize
        t0 = time.perf_counter()
        job()
        sync_all()
        per_step.append((time.perf_counter() - t0) * 1e3)
    launches = L.damgpu_launch_count() - l0
    sampler.stop_flag.set()
    if rank == 0:
        sampler.join()
    ps = torch.tensor(per_step, dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(bases), float(nrec[0])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ps, op=dist.ReduceOp.MAX)        # a job ends when its slowest rank ends
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    per_step = [float(x) for x in ps.tolist()]
    step_ms = float(np.mean(per_step))
    total_bases, total_rec = float(tot[0].item()), int(tot[1].item())
    if rank == 0:
        val = total_bases / (step_ms / 1e3)
        out = {"metric": METRIC, "value": val, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": step_ms, "ms_per_step_median": float(np.median(per_step)), "ms_per_step_min": float(min(per_step)),
               "ms_steps": per_step, "value_median": total_bases / (float(np.median(per_step)) / 1e3),
               "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int64",
               "data": "synthetic",
               "config": {"workload": "C4: synthetic 250 Mbp reference (8 contigs, one block) + %d simulated 10 kbp reads at 15%% error in "
                                      "8 reads blocks, blocks round robin over the GPUs, k=20" % (NB * per_block),
                          "reads_blocks": NB, "reads_per_block": per_block, "read_bases": int(total_bases), "kmer": 20,
                          "l2": "inputs larger than L2: a reads block is 250 MB, the reference lists 4 GB per strand",
                          "index": "built locally by every rank, once per job"},
               "e2e": {"value": val, "unit": UNIT, "ms_per_step": step_ms,
                       "h2d_bytes_per_step": int(sum(b.packed.size for b in blocks) + 2 * hg.packed.size),
                       "d2h_bytes_per_step": None},
               "records": total_rec, "gpu_launches": int(launches), "clocks": sampler.summary()}
        os.write(real_out, (json.dumps(out) + "\n").encode())
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="damgpu")
    ap.add_argument("--index", default="local", choices=["local", "broadcast"])
    ap.add_argument("--workload", default="C2", choices=["C2", "C4"])
    ap.add_argument("--c4-scale", type=float, default=1.0, help="C4 only: fraction of the 200k reads (dev)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    if args.impl == "reference":
        run_reference(args)
        return
    if args.workload == "C4":
        run_c4(args)
        return

    # only the JSON line goes to the real stdout (NCCL and torchrun banners go to stderr)
    real_out = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    from damapper_b200 import api, dazzdb

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
    bcast = (world > 1 and args.index == "broadcast")

    # every rank maps its own reads block (different seed) against the same reference
    contigs, rb, rl, freq = make_workload(seed=7 + 100 * rank)
    rd = dazzdb.load_block((rb, rl))
    rf = dazzdb.load_block(contigs)
    rc = dazzdb.load_block(dazzdb.revcomp_contigs(contigs))
    hr = pinned_block(api, rd, torch, packed=True)
    hg, hc = pinned_block(api, rf, torch), pinned_block(api, rc, torch)
    bases = int(rl.sum())
    api.set_filter_params(20, 0, 4)
    api.set_options(mem_limit=64 << 30)
    spec = api.CAlignSpec(0.85, 100, (C.c_float * 4)(*freq))

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    nref = [0]

    def ref_index(dref):
        """Reference index of the block's current orientation."""
        if not bcast:
            return api.Index(dref)
        idx, payload = None, None
        if rank == 0:
            idx = api.Index(dref)
            nref[0] = len(idx)
            payload = torch.empty((len(idx) + 2) * 16, dtype=torch.uint8, device="cuda")
            L.damgpu_index_export(idx.h, payload.data_ptr())
        # every rank knows the length from the block itself (k-mers = sum(rlen - k + 1), map.c:676): no
        # length exchange, no host synchronisation before the payload
        ln = int((rf[2].astype(np.int64) - 19).clip(min=0).sum())
        buf = payload if rank == 0 else torch.empty((ln + 2) * 16, dtype=torch.uint8, device="cuda")
        dist.broadcast(buf, 0)
        if rank == 0:
            return idx
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

    def step_e2e(sortdir):
        """The four map.h calls with host buffers (H2D and D2H inside).  Every step writes its .las files into a
        directory of its own, as a run does: re-writing the previous step's files would add their truncation
        (5 ms per step on an ext4 volume mounted with discard) to a step that has nothing to truncate."""
        api.set_options(mem_limit=64 << 30, sort_path=sortdir)
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
    rtot = (C.c_double * 4)()
    L.damgpu_radix_totals(rtot, 1)                       # reset: the timed steps only
    sampler = ClockSampler(local)
    sync_all()
    if rank == 0:                                        # one poller per node: nvidia-smi takes driver locks
        sampler.start()
    l0 = L.damgpu_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    t0 = time.perf_counter()
    sort_ms, sort_n, ext_ms, aln_ms, join_ms, flt = [], 0, [], [], [], []
    for _ in range(args.steps):
        nrec, nbytes = step_resident(dr, dg)
        flt.append(stats["filter"])
        aln_ms.append(stats["report"]["align_ms"])
        join_ms.append(stats["join_fwd"]["lut_ms"] + stats["join_fwd"]["match_ms"] +
                       stats["join_rc"]["lut_ms"] + stats["join_rc"]["match_ms"])
    ev1.record()
    sync_all()
    t1 = time.perf_counter()
    launches = L.damgpu_launch_count() - l0
    L.damgpu_radix_totals(rtot, 1)
    radix_bytes, radix_ms, radix_launches, radix_sorts = [float(x) for x in rtot]
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

    # ---- side region (NOT part of a step): k_radix_pass on the WHOLE reads list, Sort_Kmers as the
    # reference does it -- the form taken with -t, -m, more than two reference blocks, or when the list
    # is asked for -- timed live with CUDA events on the library's stream, the list (2.2 GB) far larger than L2
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
    tmpdir = tempfile.mkdtemp(prefix="bench_e2e_", dir=scratch_base())
    sortdirs = [os.path.join(tmpdir, "s%d" % i) for i in range(2 + args.steps)]
    for d in sortdirs:
        os.makedirs(d)
    try:
        for i in range(2):
            step_e2e(sortdirs[i])
        sync_all()
        t0 = time.perf_counter()
        for i in range(args.steps):
            step_e2e(sortdirs[2 + i])
        sync_all()
        e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
        sampler.stop_flag.set()                          # clocks were sampled over both timed regions
        if rank == 0:
            sampler.join()
    finally:
        sampler.stop_flag.set()
        shutil.rmtree(tmpdir, ignore_errors=True)
    if world > 1:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    # what one e2e step copies: the packed reads + their offsets, the reference block in both orientations
    # and once more for the Reporter
    h2d = int(hr.packed.size + 8 * hr.poff.size + 8 * hr.boff.size + 4 * hr.rlen.size
              + 3 * hg.bases.size + 3 * (8 * hg.boff.size + 4 * hg.rlen.size))
    d2h = int(nbytes + 16 * hr.nreads)

    if rank == 0:
        peak, which = peaks()
        n = stats["kmers"]
        pass_ms = float(np.mean(sort_ms)) / max(sort_n, 1)
        achieved_full = 32.0 * n / (pass_ms / 1e3) / 1e9
        traffic_full = profile_json("r01_radix_pass_traffic.json").get("dram_bytes_per_launch")
        instep = profile_json("r02_radix_pass_in_step.json")
        alnp = profile_json("r02_k_align_duo.json")
        rs = stats["report"]
        in_ms = radix_ms / max(radix_launches, 1.0)
        in_ach = radix_bytes / max(radix_ms, 1e-9) / 1e6          # bytes per ms -> GB/s
        fe = float(np.mean([f["extract_ms"] for f in flt]))
        jb = 2 * 16.0 * (stats["join_fwd"]["alen"] + stats["join_fwd"]["blen"])
        jm = max(float(np.mean(join_ms)), 1e-9)
        am = max(float(np.mean(aln_ms)), 1e-9)
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
            "config": config_of(hr.nreads, bases, n, "built on rank 0 and broadcast with NCCL" if bcast else "built locally"),
            "e2e": {"value": total_bases / (e2e_ms / 1e3), "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            # the HBM-bound kernel of the timed path: every k_radix_pass launch of the timed steps (reference
            # lists of both orientations, the surviving reads k-mers, the seed lists), CUDA events around the
            # passes of each sort on the launching stream.  Lists of 4-9 M records: a pass is 25-45 us.
            "roofline": {"kernel": "k_radix_pass, all launches inside the timed steps", "bound": "hbm",
                         "achieved": in_ach, "peak": peak, "peak_source": which + " copy bandwidth, burst",
                         "unit": "GB/s", "frac": in_ach / peak,
                         "traffic": instep.get("dram_bytes_per_launch"),
                         "algorithmic_bytes_per_launch": radix_bytes / max(radix_launches, 1.0),
                         "avg_launch_ms": in_ms,
                         "launches_per_step": radix_launches / args.steps, "sorts_per_step": radix_sorts / args.steps,
                         "share_of_step": radix_ms / args.steps / step_ms},
            # the kernel that dominates the step; no byte roofline applies (integer ALU / dependent latency):
            # furthest-reaching cells per second live, issue-slot / ALU-pipe / lane figures from the committed
            # ncu capture of the same kernel on the same workload
            "roofline_align": {"kernel": "k_align_duo + k_unwind", "bound": "issue slots / dependent latency",
                               "ms": am, "share_of_step": am / step_ms,
                               "cells_per_s": rs["ncells"] / (am / 1e3), "waves_per_s": rs["nwaves"] / (am / 1e3),
                               "cells_per_wave": rs["ncells"] / max(rs["nwaves"], 1),
                               "ncu": alnp},
            # side region, not in a step (launches_per_step 0): the same pass kernel on the whole reads list
            "roofline_sort_full": {"kernel": "k_radix_pass (one LSD pass over the whole reads k-mer list)", "bound": "hbm",
                                   "note": "timed in its own region after the steps: a step builds the reads list only from "
                                           "the k-mers that occur in the reference block (roofline_filter), %d of %d records"
                                           % (int(np.mean([f["survivors"] for f in flt])), n),
                                   "achieved": achieved_full, "peak": peak, "unit": "GB/s", "frac": achieved_full / peak,
                                   "traffic": traffic_full, "algorithmic_bytes_per_launch": 32 * n,
                                   "avg_launch_ms": pass_ms, "launches_per_step": 0},
            # merge-join of both orientations of a step: SURVEY 8(d) counts 16 B per input record per
            # scan; here the shorter list drives and finds its codes in the longer one through a prefix
            # table (built once per reads block), so the second call never scans the reads list
            "roofline_merge": {"kernel": "k_build_lut + k_join_match, both orientations of a step", "bound": "hbm",
                               "algorithmic_bytes": jb, "ms": jm, "achieved": jb / (jm / 1e3) / 1e9,
                               "peak": peak, "unit": "GB/s", "frac": jb / (jm / 1e3) / 1e9 / peak},
            # the reads-side index of a step: hash bitmap of the reference codes (both orientations),
            # extraction with a membership test + ordered compaction, radix passes over the survivors.
            # Bound by bitmap lookups in L2 (one 32-byte sector per k-mer), not by HBM.
            "roofline_filter": {"kernel": "k_extract_filtered", "bound": "l2 lookups", "ms": fe,
                                "lookups_per_s": n / (max(fe, 1e-9) / 1e3),
                                "survivors": int(np.mean([f["survivors"] for f in flt])), "kmers": n,
                                "algorithmic_bytes": int(bases + hr.nreads + 16 * np.mean([f["survivors"] for f in flt]))},
            "phases_ms": {"ref_bitmap": float(np.mean([f["bitmap_ms"] for f in flt])),
                          "extract_filtered": fe,
                          "radix_sort_survivors": float(np.mean([f["sort_ms"] for f in flt])),
                          "radix_passes_in_step": radix_ms / args.steps,
                          "align_kernel": am,
                          "full_sort_extract": float(np.mean(ext_ms)), "full_sort_radix": float(np.mean(sort_ms))},
            "extension": {"cells_per_s": rs["ncells"] / (am / 1e3),
                          "waves": rs["nwaves"], "alignments": rs["nalign"], "records": int(nrec)},
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                wp, cpu = whole_process(contigs, rb, rl, local)
                if wp is not None:
                    out["whole_process"] = wp
                    out["cpu_baseline"] = cpu
            except Exception as e:                       # the baseline must not take the line down
                out["cpu_baseline"] = {"error": str(e)[:300]}
        os.write(real_out, (json.dumps(out) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
