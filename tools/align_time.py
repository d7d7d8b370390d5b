"""Dev: alignment-phase time (CUDA events around the first-tier kernel + k_unwind) of resident C2 steps."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from damapper_b200 import synth, dazzdb, api
api.init()
contigs, rb, rl = synth.make_config(os.environ.get("CFG", "C2"), scale=float(os.environ.get("SCALE", "1.0")), seed=7)
rd = dazzdb.load_block((rb, rl)); rf = dazzdb.load_block(contigs)
api.set_filter_params(20, 0, 4); api.set_options()
hr, hg = api.HostBlock(*rd), api.HostBlock(*rf)
dr = api.DeviceBlock(hr); dg = api.DeviceBlock(hg)
L = api.load(); L.damgpu_time_kernels(1)
out = []
for it in range(int(os.environ.get("STEPS", "4"))):
    ir = api.Index(dr, deferred=True); m = api.Mapper(dr, ir)
    ig = api.Index(dg); m.match(dg, ig, 0, 1); ig.free()
    dg.complement(); ig = api.Index(dg); m.match(dg, ig, 1, 0); ig.free(); dg.complement()
    rep = m.report(dg, 0.85, 100, (.25, .25, .25, .25), int(os.environ.get("DOB", "0")))
    st = rep.stats()
    out.append("%.3f" % st["align_ms"])
    last = (rep.records(0), st["nwaves"], st["overflow_jobs"])
    rep.free(); m.free(); ir.free()
print(os.environ.get("DAMGPU_LIB", "default"), "align_ms", " ".join(out), "records/waves/overflow", last, flush=True)
