"""Dev check: whole pipeline on the GPU vs the oracle (candidates, records, -p track)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from damapper_b200 import synth, dazzdb, api
from oracle import oracle as orc

def check(cfg, scale, seed, **kw):
    contigs, rb, rl = synth.make_config(cfg, scale=scale, seed=seed)
    rd = dazzdb.load_block((rb, rl)); rf = dazzdb.load_block(contigs)
    rc = dazzdb.load_block(dazzdb.revcomp_contigs(contigs))
    allb = np.concatenate(contigs); cnt = np.bincount(allb, minlength=4).astype(np.float64)
    freq = tuple(float(x) for x in (cnt / cnt.sum()).astype(np.float32))
    t0 = time.time()
    o = orc.map_block(orc.HostBlock(*rd), [(orc.HostBlock(*rf), orc.HostBlock(*rc))], orc.HostBlock(*rf), freq=freq, **kw)
    t1 = time.time()
    g = api.map_block(api.HostBlock(*rd), [api.HostBlock(*rf)], api.HostBlock(*rf), freq=freq, want_candidates=True, **kw)
    t2 = time.time()
    oc, ojc, oj = o["candidates"]; gc, gjc, gj = g["candidates"]
    okc = oc.tobytes() == gc.tobytes() and (ojc == gjc).all() and oj.tobytes() == gj.tobytes()
    oka = o["a"] == g["a"]; okb = o["b"] == g["b"]; okp = o["prof"] == g["prof"]
    ost = o["stats"]; gst = g["stats"]
    oks = (ost["nalign"], ost["nwaves"], ost["ncells"]) == (gst["nalign"], gst["nwaves"], gst["ncells"])
    print(cfg, scale, seed, kw, "cand", okc, len(oc), len(gc), "A", oka, len(o["a"]), len(g["a"]), "B", okb, len(o["b"]), len(g["b"]),
          "prof", okp, "stats", oks, gst, "oracle %.2fs gpu %.2fs" % (t1 - t0, t2 - t1), flush=True)
    if not oka and len(o["a"]) == len(g["a"]):
        a = np.frombuffer(o["a"], dtype=np.uint8); b = np.frombuffer(g["a"], dtype=np.uint8)
        d = np.nonzero(a != b)[0]; print("  first diff at byte", d[:5], flush=True)
    return okc and oka and okb and okp and oks

if __name__ == "__main__":
    api.init()
    good = True
    good &= check("C1", 0.02, 3)
    good &= check("C1", 0.1, 4, do_b=1)
    good &= check("C1", 0.1, 12, do_b=1, profile=1)
    good &= check("C5", 0.1, 13, do_b=1, profile=1)
    good &= check("C5", 0.2, 14, do_b=1, best_tie=0.8)
    good &= check("C3", 0.002, 15, profile=1, best_tie=0.95)
    good &= check("C3", 0.004, 16, do_b=1, profile=1, best_tie=0.7)
    good &= check("C1", 0.05, 17, kmer=16, spacing=50)
    good &= check("C1", 0.05, 18, kmer=14, ave_corr=0.8, suppress=20, do_b=1)
    good &= check("C1", 0.05, 19, kmer=24, spacing=200, do_b=1)
    print("ALL OK" if good else "MISMATCH", flush=True)
