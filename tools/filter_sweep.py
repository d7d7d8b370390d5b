"""Bitmap size sweep of the filtered reads index at C2 size (dev aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from damapper_b200 import synth, dazzdb, api
api.init()
contigs, rb, rl = synth.make_config("C2", scale=1.0, seed=7)
rd = dazzdb.load_block((rb, rl)); rf = dazzdb.load_block(contigs)
api.set_filter_params(20, 0, 4); api.set_options()
hr, hg = api.HostBlock(*rd), api.HostBlock(*rf)
dr = api.DeviceBlock(hr); dg = api.DeviceBlock(hg)
L = api.load()
L.damgpu_time_kernels(1)
ig = api.Index(dg)
for bits in (0, 27, 29):
    api.set_reads_filter("always", bits)
    for it in range(3):
        ir = api.Index(dr, deferred=True)
        t0 = time.perf_counter()
        s = api.Seeds(ir, dr, ig, dg)
        dt = (time.perf_counter() - t0) * 1e3
        f = api.last_filter_times(); j = api.last_join_times()
        cnt = s.count
        s.free(); ir.free()
    print("bits", bits, "hits", cnt, "wall_ms %.2f" % dt, f, j, flush=True)
