"""Dev: larger shapes than the unit tests cover, GPU vs the oracle.
   C3-like: repeat-rich reference, -n.95 -p;   C4-like: chromosome-scale reference (250 Mbp, 3 contigs), few reads."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from damapper_b200 import synth, dazzdb, api
from oracle import oracle as orc

api.init()

def run(name, contigs, rb, rl, **kw):
    rd = dazzdb.load_block((rb, rl)); rf = dazzdb.load_block(contigs); rc = dazzdb.load_block(dazzdb.revcomp_contigs(contigs))
    cnt = np.bincount(np.concatenate([c[:1000000] for c in contigs]), minlength=4).astype(np.float64)
    freq = tuple(float(x) for x in (cnt / cnt.sum()).astype(np.float32))
    t1 = time.time()
    o = orc.map_block(orc.HostBlock(*rd), [(orc.HostBlock(*rf), orc.HostBlock(*rc))], orc.HostBlock(*rf), freq=freq, **kw)
    t2 = time.time()
    ok = True
    for mode in ("auto", "always", "off"):          # the reads list: policy, filtered whatever the size, sorted in full
        t0 = time.time()
        g = api.map_block(api.HostBlock(*rd), [api.HostBlock(*rf)], api.HostBlock(*rf), freq=freq, reads_filter=mode, **kw)
        tg = time.time() - t0
        good = (g["a"] == o["a"]) and (g["b"] == o["b"]) and (g["prof"] == o["prof"])
        ok &= good
        print("%s [reads list %s]: ref %.1f Mbp, %d reads, records %d/%d, parity %s, stats %s, gpu %.2fs oracle %.1fs" % (
            name, mode, sum(c.size for c in contigs) / 1e6, len(rl), g["anrec"], o["anrec"], good, g["stats"], tg, t2 - t1), flush=True)
    return ok

ok = True
contigs, rb, rl = synth.make_config("C3", scale=float(os.environ.get("C3SCALE", "0.1")), seed=51)
ok &= run("C3-like -n.95 -p", contigs, rb, rl, best_tie=0.95, profile=1, do_b=1)
if os.environ.get("BIG", "1") == "1":
    G = int(float(os.environ.get("C4MBP", "250")) * 1e6)
    genome = synth.make_genome(G, seed=52)
    cuts = np.array([0, int(G * 0.5), int(G * 0.8), G])
    contigs = [genome[cuts[i]:cuts[i + 1]] for i in range(3)]
    rb, rl, _ = synth.make_reads(genome, 3000, seed=53, contig_bounds=cuts)
    ok &= run("C4-like 250 Mbp", contigs, rb, rl)
print("ALL OK" if ok else "MISMATCH")
