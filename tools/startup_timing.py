"""Dev: where a fresh process spends its first seconds (no torch): init, first allocations, first kernels."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
t = time.perf_counter
t0 = t()
from damapper_b200 import synth, dazzdb, api
import numpy as np
contigs, rb, rl = synth.make_config("C2", scale=0.25, seed=7)
rd = dazzdb.load_block((rb, rl))
hr = api.HostBlock(*rd)
t1 = t(); L = api.init(0); t2 = t()
import ctypes as C
f, tot = C.c_uint64(0), C.c_uint64(0)
L.damgpu_device_memory(C.byref(f), C.byref(tot)); t3 = t()
d1 = api.DeviceBlock(hr, packed=True); t4 = t()
d2 = api.DeviceBlock(hr, packed=True); t5 = t()
d3 = api.DeviceBlock(hr); t6 = t()
api.set_filter_params(20, 0, 4); api.set_options()
i1 = api.Index(d1); n = len(i1); x = i1.download()[:1]; t7 = t()
i2 = api.Index(d2); x = i2.download()[:1]; t8 = t()
print("init %.1f | meminfo %.1f | first packed upload %.1f | second %.1f | plain upload %.1f | first index+download %.1f | second %.1f ms" %
      tuple(1e3 * v for v in (t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5, t7 - t6, t8 - t7)))
