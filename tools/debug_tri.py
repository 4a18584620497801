import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from pmda_b200 import synthetic, engine, _cabi
from pmda_b200.rdf import InterRDF
from pmda_b200.universe import Universe
from helpers import oracle_interrdf
dims = [20.0, 22.0, 24.0, 70.0, 80.0, 65.0]
rng = np.random.default_rng(17)
pos = synthetic.uniform_triclinic(1, 400, dims, rng)
u = Universe(pos, dims)
r = InterRDF(u.atoms, u.atoms).run()
want, _, _ = oracle_interrdf(u, u.atoms, u.atoms)
print("gpu ", r.count[:20], r.count.sum())
print("want", want[:20], want.sum())
print("diff", (r.count - want)[:40], (r.count-want).sum())
st = engine._state(torch.device("cuda", 0))
print(_cabi.rdf_stats(st.workspace[(-st.workspace.data_ptr()) % 256:], st.compute_stream.cuda_stream))
