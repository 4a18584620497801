"""Golden vectors for the N-rank invariance check (`parity_nranks` of
bench.py, tests/test_gpu_nranks.py): a fixed seeded small trajectory whose
InterRDF / InterRDF_s results must be the same integers for every number of
GPUs, ranks and blocks -- the invariance the reference pins in
pmda/test/test_rdf.py:76-93 (n_blocks in 1..4) and test_rdf_s.py:110-138.

    python tests/golden/make_nranks_golden.py        # ~2 min on 8 cores

Produced by the ORACLE (oracle/rdf_oracle.c); inputs are regenerated from the
seed wherever they are needed and fingerprinted here.  Besides the totals over
all 64 frames the fixture keeps the counts of frame 0 alone, so the CPU suite
can re-check the oracle against it in a second.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

N_FRAMES = 64
SEED = 20261020

CASES = {
    # name: (n_atoms, dims, nbins, range)
    "ortho": (20_000, [58.0, 60.0, 62.0, 90.0, 90.0, 90.0], 100, (0.0, 12.0)),
    "triclinic": (8_000, [46.0, 48.0, 50.0, 70.0, 80.0, 65.0], 75, (0.0, 10.0)),
}
N_SITES, N_NEAR = 16, 64          # InterRDF_s on the ortho trajectory


def make_positions(name, frames=None):
    """float32 [N_FRAMES, n, 3]; every frame has its own seed, `frames`
    limits which ones are generated (the rest stays zero)"""
    from pmda_b200 import mdamath
    n, dims, _, _ = CASES[name]
    m = mdamath.triclinic_vectors(np.asarray(dims, np.float32),
                                  dtype=np.float64)
    pos = np.zeros((N_FRAMES, n, 3), dtype=np.float32)
    tag = sorted(CASES).index(name)
    for f in (range(N_FRAMES) if frames is None else frames):
        gen = np.random.Generator(np.random.PCG64([SEED, tag, f]))
        frac = 1e-4 + (1.0 - 2e-4) * gen.random((n, 3))
        pos[f] = (frac @ m).astype(np.float32)
    return pos, dims


def site_indices():
    """InterRDF_s site pairs on the ortho trajectory: ion i against N_NEAR
    fixed atoms (seeded choice, no geometry involved)"""
    n = CASES["ortho"][0]
    gen = np.random.Generator(np.random.PCG64([SEED, 99]))
    ions = gen.choice(n, N_SITES, replace=False)
    return [(np.array([i]), np.sort(gen.choice(n, N_NEAR, replace=False)))
            for i in ions]


def oracle_rdf(name, frames):
    from oracle import c_oracle
    pos, dims = make_positions(name, frames)
    _, _, nbins, rng = CASES[name]
    count = np.zeros(nbins, dtype=np.int64)
    for f in frames:
        c_oracle.rdf_single_frame(pos[f], pos[f], dims, nbins, rng,
                                  count=count, nthreads=os.cpu_count(),
                                  cells=(name == "ortho"))
    return count


def oracle_rdf_s(frames):
    from oracle import c_oracle
    pos, dims = make_positions("ortho", frames)
    _, _, nbins, rng = CASES["ortho"]
    sites = site_indices()
    counts = [np.zeros((len(a), len(b), nbins)) for a, b in sites]
    for f in frames:
        c_oracle.rdf_s_single_frame([(pos[f][a], pos[f][b]) for a, b in sites],
                                    dims, nbins, rng, counts=counts)
    return np.stack([c[0] for c in counts])          # [N_SITES, N_NEAR, nbins]


def digest(name):
    pos, _ = make_positions(name)
    return hashlib.sha256(pos.tobytes()).hexdigest()


if __name__ == "__main__":
    out = {}
    for name in sorted(CASES):
        out[name + "__sha256"] = np.frombuffer(digest(name).encode(), np.uint8)
        out[name + "__frame0"] = oracle_rdf(name, [0])
        out[name + "__total"] = oracle_rdf(name, list(range(N_FRAMES)))
        print(name, int(out[name + "__total"].sum()), flush=True)
    out["sites__frame0"] = oracle_rdf_s([0]).astype(np.int32)
    out["sites__total"] = oracle_rdf_s(list(range(N_FRAMES))).astype(np.int32)
    print("sites", int(out["sites__total"].sum()))
    np.savez_compressed(os.path.join(HERE, "nranks_counts.npz"), **out)
