"""Contacts throughput on one GPU: two 600-atom groups
of a 20,000-atom system, 2,000 frames, hard_cut and soft_cut, device-resident
and end to end, plus the numpy restatement's time per frame.  Documentation
numbers; not the contract bench."""
import sys, os, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from pmda_b200 import engine
from pmda_b200.contacts import Contacts
from pmda_b200.universe import Universe

rng = np.random.default_rng(5)
F, N = 2000, 20000
base = rng.random((N, 3)) * 60.0
pos = (base[None] + rng.normal(0, 0.4, (F, N, 3))).astype(np.float32)
u = Universe(pos, None)
a, b = u.atoms[:600], u.atoms[600:1200]
ref = Universe(pos[:1].copy(), None)
refs = (ref.atoms[:600], ref.atoms[600:1200])
for method in ("hard_cut", "soft_cut"):
    make = lambda: Contacts((a, b), refs, method=method, radius=6.0)
    res = []
    for cache in (64 << 30, 0):
        engine.Config.cache_bytes = cache
        engine.clear_cache(); make().run()
        best = 1e9
        for _ in range(3):
            c = make(); torch.cuda.synchronize(); t0 = time.perf_counter(); c.run(); torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        res.append(F / best)
    npairs = int(c.initial_contacts[0].sum())
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import numpy_oracle
    t0 = time.perf_counter()
    for f in range(20):
        numpy_oracle.contacts_single_frame(f, pos[f][a.indices], pos[f][b.indices],
                                           c.initial_contacts, c.r0, c.fraction_contacts,
                                           c.fraction_kwargs)
    cpu = 20 / (time.perf_counter() - t0)
    print(json.dumps({"method": method, "contact_pairs": npairs, "frames": F,
                      "device_resident_frames_per_s": res[0], "e2e_frames_per_s": res[1],
                      "kernel_frames_per_s": F / float(c.timing.compute.sum()),
                      "cpu_numpy_frames_per_s": cpu}), flush=True)
