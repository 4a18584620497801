"""Contacts: host logic against the numpy restatement of
pmda/contacts.py:270-289 (CPU, kernels replaced by tests/fake_backend.py) and
GPU parity of rdfk_contacts through the public API.  Mirrors the structure of
pmda/test/test_contacts.py (frame slicing, uneven blocks, own methods, q1q2);
its golden numbers need MDAnalysisTests data files that are absent here."""
import numpy as np
import pytest

from pmda_b200 import contacts, engine
from pmda_b200.contacts import Contacts, q1q2
from pmda_b200.universe import Universe

from oracle import numpy_oracle
from fake_backend import OracleBackend


def make_universe(n_frames=12, n=160, seed=31, spread=14.0):
    rng = np.random.default_rng(seed)
    base = rng.random((n, 3)) * spread
    pos = base[None] + rng.normal(0, 1.2, (n_frames, n, 3)).cumsum(axis=0) * 0.3
    return Universe(pos.astype(np.float32), None)


def oracle_timeseries(u, ca, frames):
    grA, grB = ca._atomgroups
    rows = []
    for i in frames:
        ts = u.trajectory[i]
        rows.append(numpy_oracle.contacts_single_frame(
            ts.frame, ts.positions[grA.indices], ts.positions[grB.indices],
            ca.initial_contacts, ca.r0, ca.fraction_contacts,
            ca.fraction_kwargs))
    return np.array(rows)


@pytest.fixture
def backend():
    be = OracleBackend(n_devices=2)
    engine._set_test_backend(be)
    engine.clear_cache()
    yield be
    engine._set_test_backend(None)
    engine.clear_cache()


def _is_any_closer(r, r0, dist=2.5):
    return np.any(r < dist)


def _weird_own_method(r, r0):
    return 'aaa'


def test_restated_fraction_functions():
    r = np.array([1.0, 2.0, 3.0, 4.0])
    r0 = np.array([1.5, 1.5, 3.0, 3.5])
    assert contacts.hard_cut_q(r, r0) == 0.5
    assert contacts.radius_cut_q(r, r0, 2.0) == 0.5
    want = (1 / (1 + np.exp(5.0 * (r - 1.8 * r0)))).sum() / 4
    assert contacts.soft_cut_q(r, r0) == want
    d = np.array([[1.0, 5.0], [4.5, 4.6]])
    np.testing.assert_array_equal(contacts.contact_matrix(d, 4.5),
                                  [[True, False], [True, False]])
    a = np.random.default_rng(0).random((7, 3)).astype(np.float32) * 9
    b = np.random.default_rng(1).random((5, 3)).astype(np.float32) * 9
    np.testing.assert_array_equal(contacts.distance_array(a, b),
                                  numpy_oracle.distance_array_nobox(a, b))


@pytest.mark.parametrize("method,kwargs", [
    ("hard_cut", None), ("soft_cut", None),
    ("soft_cut", {"beta": 4.0, "lambda_constant": 1.5}),
    (contacts.radius_cut_q, {"radius": 5.0}),
    (_is_any_closer, None), (_is_any_closer, {"dist": 3.5}),
])
def test_host_logic_against_restatement(backend, method, kwargs):
    u = make_universe()
    A, B = u.atoms[:70], u.atoms[70:]
    ca = Contacts((A, B), (A, B), method=method, radius=6.0, kwargs=kwargs)
    ca.run(n_blocks=3)
    want = oracle_timeseries(u, ca, range(12))
    assert ca.timeseries.shape == (12, 2)
    np.testing.assert_array_equal(ca.timeseries[:, 0], np.arange(12))
    np.testing.assert_allclose(ca.timeseries[:, 1], want[:, 1], rtol=1e-13,
                               atol=0)
    if method == "hard_cut":
        np.testing.assert_array_equal(ca.timeseries, want)
        assert ca.timeseries[0, 1] == 1.0       # frame 0 is the reference


def test_slicing_blocks_and_empty(backend):
    u = make_universe()
    A, B = u.atoms[:70], u.atoms[70:]

    def run(**kw):
        return Contacts((A, B), (A, B), radius=6.0).run(**kw)

    assert len(run().timeseries) == u.trajectory.n_frames
    assert len(run(n_blocks=5).timeseries) == u.trajectory.n_frames
    with pytest.warns(UserWarning):
        assert len(run(stop=0).timeseries) == 0
    r = run(start=2, stop=11, step=3, n_blocks=2)
    np.testing.assert_array_equal(r.timeseries[:, 0], [2, 5, 8])
    full = run().timeseries
    np.testing.assert_array_equal(r.timeseries[:, 1], full[[2, 5, 8], 1])
    two = run(n_jobs=2, n_blocks=4).timeseries       # two "devices"
    np.testing.assert_array_equal(two, full)


def test_several_references_and_q1q2(backend):
    u = make_universe()
    A = u.atoms[:90]
    q = q1q2(A, radius=7.0).run(n_blocks=2)
    assert q.timeseries.shape == (12, 3)
    assert q.timeseries[0, 1] == 1.0 and q.timeseries[-1, 2] == 1.0
    want = oracle_timeseries(u, q, range(12))
    np.testing.assert_array_equal(q.timeseries, want)


def test_bad_methods(backend):
    u = make_universe()
    A, B = u.atoms[:70], u.atoms[70:]
    with pytest.raises(ValueError):
        Contacts((A, B), (A, B), method=_weird_own_method).run(stop=2)
    with pytest.raises(ValueError):
        Contacts((A, B), (A, B), method=2)


def test_single_frame_hook(backend):
    u = make_universe()
    A, B = u.atoms[:70], u.atoms[70:]
    ca = Contacts((A, B), (A, B), radius=6.0)
    ca._prepare()
    y = ca._single_frame(u.trajectory[4], (A, B))
    want = oracle_timeseries(u, ca, [4])[0]
    np.testing.assert_array_equal(y, want)


# ------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("method,kwargs,exact", [
    ("hard_cut", None, True), ("soft_cut", None, False),
    (contacts.radius_cut_q, {"radius": 5.5}, True),
    (_is_any_closer, {"dist": 3.0}, True),
])
def test_gpu_contacts_match_restatement(method, kwargs, exact):
    u = make_universe(n_frames=40, n=900, seed=5, spread=30.0)
    A, B = u.atoms[:400], u.atoms[400:]
    ca = Contacts((A, B), (A, B), method=method, radius=6.0, kwargs=kwargs)
    ca.run(n_blocks=3)
    want = oracle_timeseries(u, ca, range(40))
    if exact:
        np.testing.assert_array_equal(ca.timeseries, want)
    else:
        np.testing.assert_allclose(ca.timeseries, want, rtol=1e-12, atol=0)


@pytest.mark.gpu
def test_gpu_q1q2_and_distances_bit_exact():
    u = make_universe(n_frames=25, n=600, seed=6, spread=26.0)
    A = u.atoms[:500]
    q = q1q2(A, radius=8.0).run(n_blocks=4)
    want = oracle_timeseries(u, q, range(25))
    np.testing.assert_array_equal(q.timeseries, want)
    # the emitted distances themselves (callable that returns one of them)
    pick = lambda r, r0: r[len(r) // 2]      # noqa: E731
    ca = Contacts((A, A), (A, A), method=pick, radius=5.0).run()
    want = oracle_timeseries(u, ca, range(25))
    np.testing.assert_array_equal(ca.timeseries, want)
