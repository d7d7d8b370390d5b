"""Evidence for SURVEY 8(e) / DESIGN section 5: how should N GPUs come by the sorted reference index?
Run under torchrun (one rank per GPU).  For a reference block of GBP Mbp (default 250), per strand:
  local      every rank extracts and sorts the whole list itself
  broadcast  rank 0 builds it, one NCCL broadcast of the 16 B x N_g records, the others adopt the buffer
  sharded    every rank builds 1/N of the list, one NCCL all-gather of the shards (prefix sharding is
             emulated by a shard of equal size: 1/N of the contigs; the gathered buffer is not adopted)
Times are CUDA-event / wall maxima over the ranks, best of 3."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from damapper_b200 import api, dazzdb, synth
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
L = api.init(local)
G = int(float(os.environ.get("GBP", "250")) * 1e6)
nc = 8 * max(world, 1)
genome = synth.make_genome(G, seed=5)
cuts = [int(G * i / nc) for i in range(nc + 1)]
contigs = [genome[cuts[i]:cuts[i + 1]] for i in range(nc)]
api.set_filter_params(20, 0, 4); api.set_options()
whole = api.DeviceBlock(api.HostBlock(*dazzdb.load_block(contigs)), packed=False)
per = nc // max(world, 1)
shard = api.DeviceBlock(api.HostBlock(*dazzdb.load_block(contigs[rank * per:(rank + 1) * per])), packed=False)
def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier(); torch.cuda.synchronize()
def timed(fn):
    best = None
    for _ in range(3):
        sync(); t0 = time.perf_counter(); fn(); sync(); dt = (time.perf_counter() - t0) * 1e3
        best = dt if best is None else min(best, dt)
    if world > 1:
        t = torch.tensor([best], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); best = float(t.item())
    return best
n = int(sum(c.size - 19 for c in contigs))
def do_local():
    i = api.Index(whole); i.free()
def do_bcast():
    buf = torch.empty((n + 2) * 16, dtype=torch.uint8, device="cuda")
    if rank == 0:
        i = api.Index(whole); L.damgpu_index_export(i.h, buf.data_ptr()); i.free()
    if world > 1:
        dist.broadcast(buf, 0)
    if rank != 0:
        api.Index(handle=L.damgpu_index_import(buf.data_ptr(), n)).free()
ns = int(sum(c.size - 19 for c in contigs[rank * per:(rank + 1) * per]))
def do_sharded():
    i = api.Index(shard)
    mine = torch.empty((ns + 2) * 16, dtype=torch.uint8, device="cuda")
    L.damgpu_index_export(i.h, mine.data_ptr()); i.free()
    if world > 1:
        allb = torch.empty(world * (ns + 2) * 16, dtype=torch.uint8, device="cuda")
        dist.all_gather_into_tensor(allb, mine)
out = {"gpus": world, "reference_mbp": G / 1e6, "records_per_strand": n, "list_gb": 16 * n / 1e9,
       "local_ms": timed(do_local), "broadcast_ms": timed(do_bcast), "sharded_allgather_ms": timed(do_sharded)}
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
