"""Committed golden count vectors of the per-frame path
(tests/golden/oracle_counts.npz, made by tests/golden/make_oracle_golden.py).
CPU: the oracle still reproduces them and the seeded inputs are unchanged.
GPU: the CUDA path through the public API gives the same integers."""
import hashlib
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_oracle_golden as mk        # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "oracle_counts.npz"))


@pytest.mark.parametrize("name", sorted(mk.CASES))
def test_oracle_reproduces_golden_counts(name):
    count, digest = mk.oracle_counts(name)
    assert digest == GOLD[name + "__sha256"].tobytes().decode()
    np.testing.assert_array_equal(count, GOLD[name + "__count"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(mk.CASES))
def test_cuda_path_matches_golden_counts(name):
    from pmda_b200.rdf import InterRDF
    from pmda_b200.universe import Universe
    pos, dims, i1, i2, nbins, rng, excl = mk.make_inputs(name)
    assert hashlib.sha256(pos.tobytes()).hexdigest() == \
        GOLD[name + "__sha256"].tobytes().decode()
    u = Universe(pos, dims)
    r = InterRDF(u.atoms[i1], u.atoms[i2], nbins=nbins, range=rng,
                 exclusion_block=excl).run(n_blocks=min(2, len(pos)))
    np.testing.assert_array_equal(r.count, GOLD[name + "__count"])
