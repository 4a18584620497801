"""File-backed ingest: InterRDF over an NpyTrajectory vs the same frames in
memory (cfg2 shape), frames/s and GB/s of coordinates end to end."""
import sys, os, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pmda_b200 import engine
from pmda_b200.rdf import InterRDF
from pmda_b200.universe import Universe, NpyTrajectory
from pmda_b200.dcd import write_dcd
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
L = 100.0 * (n / 1e5) ** (1 / 3)
rng = np.random.default_rng(3)
pos = np.empty((nf, n, 3), np.float32)
for f in range(nf):
    pos[f] = (rng.random((n, 3)) * L).astype(np.float32)
d = tempfile.mkdtemp(dir=os.environ.get("TMPDIR", "/tmp"))
p = os.path.join(d, "traj.npy")
NpyTrajectory.write(p, pos)
dims = [L, L, L, 90, 90, 90]
pd = os.path.join(d, "traj.dcd")
write_dcd(pd, pos, dims)
engine.Config.cache_bytes = 0          # every run re-reads and re-uploads
for name, u in (("memory (page-locked array)", Universe(pos, dims)),
                ("npy    (mmap -> pinned ring)", Universe.from_npy(p, dims)),
                ("dcd    (raw bytes -> device unpack)", Universe.from_dcd(pd))):
    for rep in range(3):
        r = InterRDF(u.atoms, u.atoms, nbins=200, range=(0.0, 15.0))
        torch.cuda.synchronize(); t0 = time.time(); r.run(); torch.cuda.synchronize()
        dt = time.time() - t0
        print("%-30s rep%d: %.4f s  %.0f frames/s  %.2f GB/s ingest  kernel %.4f s  io %.4f s" % (
            name, rep, dt, nf / dt, pos.nbytes / dt / 1e9, r.timing.compute.sum(), r.timing.io.sum()), flush=True)
os.remove(p); os.remove(pd)
