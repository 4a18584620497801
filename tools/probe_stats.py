"""librdfk's counters (fp64 re-checks, evaluated pairs, ...) of one run of a
BASELINE config shape, for the library RDFK_LIB points at."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from pmda_b200 import _cabi, engine, synthetic  # noqa: E402
from pmda_b200.rdf import InterRDF  # noqa: E402

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
nf = int(sys.argv[2]) if len(sys.argv) > 2 else 16
u, a = synthetic.CONFIGS[cfg](n_frames=nf)
r = InterRDF(a["g1"], a["g2"], nbins=a["nbins"], range=a["range"])
r.run()
st = engine._state(torch.device("cuda", torch.cuda.current_device()))
ws = st.workspaces[0]
ws = ws[(-ws.data_ptr()) % 256:]
stats = _cabi.rdf_stats(ws, st.compute_streams[0].cuda_stream)
print(os.environ.get("RDFK_LIB", "default"), "cfg%d x %d frames:" % (cfg, nf),
      stats, "binned", int(r.count.sum()))
