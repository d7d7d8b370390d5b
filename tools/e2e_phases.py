"""Dev: wall-clock of the four map.h-shaped calls of one e2e step (host buffers pinned)."""
import sys, os, time, tempfile, shutil, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from damapper_b200 import api, dazzdb
L = api.init(0)
contigs, rb, rl, freq = bench.make_workload(seed=7)
rd = dazzdb.load_block((rb, rl)); rf = dazzdb.load_block(contigs); rc = dazzdb.load_block(dazzdb.revcomp_contigs(contigs))
hr = bench.pinned_block(api, rd, torch, packed=True); hg, hc = bench.pinned_block(api, rf, torch), bench.pinned_block(api, rc, torch)
api.set_filter_params(20, 0, 4)
tmp = tempfile.mkdtemp(prefix="e2e_")
api.set_options(mem_limit=64 << 30, sort_path=tmp)
spec = api.CAlignSpec(0.85, 100, (C.c_float * 4)(*freq))
t = time.perf_counter
for it in range(4):
    blen, alen = C.c_int(0), C.c_int(0)
    t0 = t(); bindex = L.damgpu_Sort_Kmers(C.byref(hr.c), C.byref(blen)); torch.cuda.synchronize(); t1 = t()
    aindex = L.damgpu_Sort_Kmers(C.byref(hg.c), C.byref(alen)); t2 = t()
    L.damgpu_Match_Filter(C.byref(hr.c), C.byref(hg.c), bindex, blen, aindex, alen, 0, 1); t3 = t()
    aindex = L.damgpu_Sort_Kmers(C.byref(hc.c), C.byref(alen)); t4 = t()
    L.damgpu_Match_Filter(C.byref(hr.c), C.byref(hc.c), bindex, blen, aindex, alen, 1, 0); t5 = t()
    L.damgpu_Reporter(b"reads", C.byref(hr.c), b"ref", C.byref(hg.c), C.byref(spec), 1); t6 = t()
    L.damgpu_index_free(bindex); torch.cuda.synchronize(); t7 = t()
    print("Sort_Kmers(reads) %.2f | Sort_Kmers(ref) %.2f | Match fwd %.2f | Sort_Kmers(refc) %.2f | Match rc %.2f | Reporter %.2f | free %.2f | total %.2f ms" %
          tuple(1e3 * x for x in (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5, t7 - t6, t7 - t0)), flush=True)
shutil.rmtree(tmp, ignore_errors=True)
