import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, torch
from pmda_b200 import engine, _cabi
from pmda_b200.rdf import InterRDF
from pmda_b200.universe import Universe
from helpers import oracle_interrdf
rng = np.random.default_rng(102)
L = 55.0
pos = (rng.random((2, 9000, 3)) * L).astype(np.float32)
u = Universe(pos, [L, L, L, 90, 90, 90])
perm = rng.permutation(9000)
small, large = u.atoms[np.sort(perm[:500])], u.atoms[np.sort(perm[500:])]
for name, (g1, g2) in (("small,large", (small, large)), ("large,small", (large, small))):
    want, _, _ = oracle_interrdf(u, g1, g2)
    for dbg in (None, 1, 2, 3):
        os.environ.pop("RDFK_DEBUG", None)
        if dbg is not None: os.environ["RDFK_DEBUG"] = str(dbg)
        r = InterRDF(g1, g2).run()
        st = engine._state(torch.device("cuda", 0))
        stats = _cabi.rdf_stats(st.workspace[(-st.workspace.data_ptr()) % 256:], st.compute_stream.cuda_stream)
        print(name, "dbg", dbg, "sum gpu", r.count.sum(), "want", want.sum(), "equal", np.array_equal(r.count, want), stats)
