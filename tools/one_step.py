"""Two resident steps at C2 size (for ncu launch lists / captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from damapper_b200 import synth, dazzdb, api
api.init()
contigs, rb, rl = synth.make_config("C2", scale=1.0, seed=7)
rd = dazzdb.load_block((rb, rl)); rf = dazzdb.load_block(contigs)
api.set_filter_params(20, 0, 4); api.set_options()
hr, hg = api.HostBlock(*rd), api.HostBlock(*rf)
dr = api.DeviceBlock(hr); dg = api.DeviceBlock(hg)
L = api.load()
for it in range(int(os.environ.get("STEPS", "2"))):
    n0 = L.damgpu_launch_count()
    ir = api.Index(dr, deferred=os.environ.get("FULL_SORT") is None); m = api.Mapper(dr, ir)
    ig = api.Index(dg); m.match(dg, ig, 0, 1); ig.free()
    dg.complement(); ig = api.Index(dg); m.match(dg, ig, 1, 0); ig.free(); dg.complement()
    rep = m.report(dg, 0.85, 100, (.25, .25, .25, .25), 1)
    print("step", it, "records", rep.records(0), "launches", L.damgpu_launch_count() - n0, flush=True)
    rep.free(); m.free(); ir.free()
