"""cProfile of repeated make().run() of a small BASELINE config (host overhead)."""
import cProfile
import os
import pstats
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from pmda_b200 import engine, synthetic  # noqa: E402
from pmda_b200.rdf import InterRDF  # noqa: E402

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 1
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
u, kw = synthetic.sharded_config(cfg, {1: 100}[cfg])
make = lambda: InterRDF(kw["g1"], kw["g2"], nbins=kw["nbins"], range=kw["range"])
engine.Config.cache_bytes = 64 << 30
for _ in range(10):
    make().run()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(reps):
    make().run()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(32)
