"""Dev: wall-clock per phase of one resident step at C2 size (optionally with torch imported)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if os.environ.get("WITH_TORCH"):
    import torch
    torch.cuda.init(); torch.zeros(1, device="cuda")
from damapper_b200 import synth, dazzdb, api

api.init()
contigs, rb, rl = synth.make_config("C2", scale=float(os.environ.get("SCALE", "1.0")), seed=7)
rd = dazzdb.load_block((rb, rl)); rf = dazzdb.load_block(contigs)
api.set_filter_params(20, 0, 4); api.set_options()
hr, hg = api.HostBlock(*rd), api.HostBlock(*rf)
L = api.load()
if os.environ.get("TIMEK"): L.damgpu_time_kernels(1)
t = time.perf_counter
for it in range(4):
    t0 = t(); dr = api.DeviceBlock(hr); dg = api.DeviceBlock(hg); t1 = t()
    ir = api.Index(dr); t2 = t()
    m = api.Mapper(dr, ir); t3 = t()
    ig = api.Index(dg); t4 = t()
    m.match(dg, ig, 0, 1); t5 = t()
    ig.free(); dg.complement(); ig = api.Index(dg); t6 = t()
    m.match(dg, ig, 1, 0); t7 = t()
    ig.free(); dg.complement(); t8 = t()
    rep = m.report(dg, 0.85, 100, (.25,.25,.25,.25), 1); t9 = t()
    st = rep.stats(); n = rep.records(0)
    rep.free(); m.free(); ir.free(); dg.free(); dr.free(); t10 = t()
    print("upload %.1f | index reads %.1f | mapper_new %.1f | index ref %.1f | match fwd %.1f | comp+index %.1f | match rc %.1f | comp %.1f | report %.1f (align kernel %.1f) | free %.1f | total %.1f ms | recs %d" % tuple(
        [1e3*x for x in (t1-t0, t2-t1, t3-t2, t4-t3, t5-t4, t6-t5, t7-t6, t8-t7, t9-t8)] + [st["align_ms"], 1e3*(t10-t9), 1e3*(t10-t0), n]), flush=True)
