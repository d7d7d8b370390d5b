"""Dev: k_extract_filtered time (CUDA events) of resident C2 steps."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from damapper_b200 import synth, dazzdb, api
api.init()
contigs, rb, rl = synth.make_config("C2", scale=1.0, seed=7)
rd = dazzdb.load_block((rb, rl)); rf = dazzdb.load_block(contigs)
api.set_filter_params(20, 0, 4); api.set_options()
dr = api.DeviceBlock(api.HostBlock(*rd)); dg = api.DeviceBlock(api.HostBlock(*rf))
L = api.load(); L.damgpu_time_kernels(1)
out = []
for it in range(5):
    ir = api.Index(dr, deferred=True); m = api.Mapper(dr, ir)
    ig = api.Index(dg); m.match(dg, ig, 0, 1); f = api.last_filter_times(); ig.free()
    out.append("%.3f" % f["extract_ms"]); surv = f["survivors"]
    m.free(); ir.free()
print(os.environ.get("DAMGPU_LIB", "default"), "extract_filtered ms", " ".join(out), "survivors", surv, flush=True)
