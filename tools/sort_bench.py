"""Dev: Sort_Kmers of the C2 reads block, kernel times (CUDA events inside the library)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from damapper_b200 import synth, dazzdb, api
api.init()
contigs, rb, rl = synth.make_config("C2", scale=1.0, seed=7)
rd = dazzdb.load_block((rb, rl))
api.set_filter_params(20, 0, 4); api.set_options()
hr = api.HostBlock(*rd); dr = api.DeviceBlock(hr)
L = api.load(); L.damgpu_time_kernels(1)
for it in range(int(os.environ.get("ITERS", "4"))):
    ir = api.Index(dr); tm = api.last_sort_times(); n = len(ir)
    print("kmers %d extract %.3f ms (%.0f GB/s) sort %.3f ms = %.3f ms/pass (%.0f GB/s, %.1f%% of 6544.7)" % (
        n, tm["extract_ms"], 17.0*n/tm["extract_ms"]/1e6, tm["sort_ms"], tm["sort_ms"]/tm["npass"],
        32.0*n*tm["npass"]/tm["sort_ms"]/1e6, 100*32.0*n*tm["npass"]/tm["sort_ms"]/1e6/6544.7), flush=True)
    ir.free()
