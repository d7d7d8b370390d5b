"""Dev: one small mapping through the deferred (filtered) reads index, for compute-sanitizer runs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from damapper_b200 import api, synth, dazzdb
api.init(0)
contigs, rb, rl = synth.make_config("C1", scale=float(os.environ.get("SCALE", "0.05")), seed=3)
rd = dazzdb.load_block((rb, rl)); rf = dazzdb.load_block(contigs)
g = api.map_block(api.HostBlock(*rd), [api.HostBlock(*rf)], api.HostBlock(*rf), do_b=1, profile=1, reads_filter="always")
print("records", g["anrec"], "deferred", g["deferred"])
