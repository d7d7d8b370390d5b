"""Dev check: k-mer index + merge-join on the GPU vs the oracle, then timing at C2 size."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from damapper_b200 import synth, dazzdb, api
from oracle import oracle as orc

def check(cfg, scale, seed, kmer=20, suppress=0):
    contigs, rb, rl = synth.make_config(cfg, scale=scale, seed=seed)
    rd = dazzdb.load_block((rb, rl)); rf = dazzdb.load_block(contigs)
    api.set_filter_params(kmer, suppress, 4)
    api.set_options()
    hr, hg = api.HostBlock(*rd), api.HostBlock(*rf)
    dr, dg = api.DeviceBlock(hr), api.DeviceBlock(hg)
    ir, ig = api.Index(dr), api.Index(dg)
    o_r = orc.sort_kmers(orc.HostBlock(*rd), kmer, suppress)
    o_g = orc.sort_kmers(orc.HostBlock(*rf), kmer, suppress)
    g_r, g_g = ir.download(), ig.download()
    ok1 = (len(g_r) == len(o_r)) and g_r.tobytes() == o_r.tobytes()
    ok2 = (len(g_g) == len(o_g)) and g_g.tobytes() == o_g.tobytes()
    s = api.Seeds(ir, dr, ig, dg)
    os_, nh, lim, histo = orc.merge_join(o_r, o_g, 64 << 30, hr.sizeof_db, hg.sizeof_db, hr.maxlen, hr.nreads, hg.nreads)
    gs = s.download()
    ok3 = (s.count == nh) and (s.limit == lim) and gs.tobytes() == os_.tobytes() and (s.histogram() == histo).all()
    # complement on device
    dg.complement()
    rc = dazzdb.load_block(dazzdb.revcomp_contigs(contigs))
    ok4 = (dg.download_bases() == rc[0]).all()
    print(cfg, scale, seed, "k", kmer, "t", suppress, "reads idx", ok1, len(g_r), "ref idx", ok2, len(g_g), "seeds", ok3, s.count, nh, "limit", s.limit, lim, "comp", ok4, flush=True)
    return ok1 and ok2 and ok3 and ok4

if __name__ == "__main__":
    api.init()
    good = True
    good &= check("C1", 0.02, 3)
    good &= check("C1", 0.1, 4)
    good &= check("C3", 0.002, 5, kmer=14)
    good &= check("C1", 0.05, 6, kmer=16, suppress=10)
    good &= check("C5", 0.1, 7, kmer=32)
    good &= check("C1", 0.05, 8, kmer=12)
    print("ALL OK" if good else "MISMATCH", flush=True)
    # timing at C2 size
    contigs, rb, rl = synth.make_config("C2", scale=1.0, seed=7)
    rd = dazzdb.load_block((rb, rl))
    api.set_filter_params(20, 0, 4)
    hr = api.HostBlock(*rd)
    load = api.load(); load.damgpu_time_kernels(1)
    t0 = time.time(); dr = api.DeviceBlock(hr); t1 = time.time()
    for it in range(4):
        t2 = time.time(); ir = api.Index(dr); t3 = time.time()
        n = len(ir); tm = api.last_sort_times()
        print("C2 reads: upload %.1f ms, Sort_Kmers wall %.1f ms, kmers %d, extract %.3f ms (%.0f GB/s), sort %.3f ms %d passes (%.0f GB/s/pass-avg)" % (
            (t1-t0)*1e3, (t3-t2)*1e3, n, tm["extract_ms"], 17.0*n/tm["extract_ms"]/1e6, tm["sort_ms"], tm["npass"], 32.0*n*tm["npass"]/tm["sort_ms"]/1e6), flush=True)
        if it < 3: ir.free()
    hg = api.HostBlock(*dazzdb.load_block(contigs)); dg = api.DeviceBlock(hg); ig = api.Index(dg)
    for it in range(3):
        t4 = time.time(); s = api.Seeds(ir, dr, ig, dg); t5 = time.time()
        print("C2 merge-join+seed sort wall %.1f ms, seeds %d" % ((t5-t4)*1e3, s.count), flush=True)
        s.free()
