"""tools/pin_against_mdanalysis.py -- the one-command path that pins oracle/
against the real MDAnalysis where it is importable (it is not, in the build
container: DESIGN.md section 6).

CPU: the script's machinery (case generation, the reference's per-frame body
on a provider, the diff, the fixture writer) runs on a stand-in provider, and
its diff path does flag a provider that computes something else.  If a fixture
produced by the real MDAnalysis has been committed
(tests/golden/mdanalysis_pinned.npz), the oracle (CPU) and the CUDA path (GPU)
must reproduce it; until then those tests skip and parity stays "unpinned".
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import pin_against_mdanalysis as pin       # noqa: E402

PINNED = pin.DEFAULT_OUT
QUICK = ["triclinic_out_of_cell_uniform", "triclinic_on_the_faces",
         "lattice_on_bin_edges", "golden_triclinic_small_box"]


def test_dry_run_with_the_stand_in_provider(tmp_path):
    out = str(tmp_path / "pin.npz")
    bad = pin.run(pin.OracleProvider(), out, verbose=False, only=QUICK)
    assert bad == 0
    z = np.load(out)
    assert b"NOT MDAnalysis" in z["provider"].tobytes()
    for name in QUICK:
        assert z["rdf__%s__count" % name].sum() > 0
        assert len(z["rdf__%s__sha256" % name]) == 64
    for box in pin.BOXES:
        assert z["sweep__%s__dist" % box].shape == (1, 81)
        assert z["angles__%s" % box].shape == (400,)
    assert int(z["n_different"]) == 0


def test_out_of_cell_cases_really_are_out_of_cell():
    """the cases that exist to probe the triclinic pre-wrap must exercise it"""
    from oracle import numpy_oracle
    cases = pin.rdf_cases()
    for name in ("triclinic_out_of_cell_lattice_shifts",
                 "triclinic_out_of_cell_uniform"):
        pos, dims = cases[name][0], cases[name][1]
        m = numpy_oracle.triclinic_vectors(np.asarray(dims, np.float32))
        wrapped = numpy_oracle._wrap_triclinic(pos[0].copy(), m)
        moved = (wrapped != pos[0]).any(axis=1).mean()
        assert moved > 0.8, (name, moved)


class _WrongProvider(pin.OracleProvider):
    """computes the distances in float32 and cuts with `<`: what a careless
    restatement would do"""

    def __init__(self, what):
        pin.OracleProvider.__init__(self)
        self.what = what

    def capped_distance(self, ref, conf, max_cutoff, box, method=None):
        pairs, dist = pin.OracleProvider.capped_distance(self, ref, conf,
                                                         max_cutoff, box)
        if self.what == "float32":
            dist = dist.astype(np.float32).astype(np.float64)
        elif self.what == "scaled":
            dist = dist * (1.0 - 1e-7)
        return pairs, dist

    def distance_array(self, ref, conf, box):
        d = pin.OracleProvider.distance_array(self, ref, conf, box)
        return d.astype(np.float32).astype(np.float64) \
            if self.what == "float32" else d


@pytest.mark.parametrize("what", ["float32", "scaled"])
def test_diff_path_flags_a_provider_that_computes_something_else(tmp_path,
                                                                 what, capsys):
    out = str(tmp_path / "pin.npz")
    bad = pin.run(_WrongProvider(what), out, verbose=True,
                  only=["lattice_on_bin_edges"])
    text = capsys.readouterr().out
    # atoms exactly on bin edges: any change of the arithmetic moves counts
    assert bad >= 1 and "DIFFERENT" in text
    assert int(np.load(out)["n_different"]) == bad


def test_without_mdanalysis_the_script_says_so(monkeypatch, capsys):
    try:
        import MDAnalysis      # noqa: F401
        pytest.skip("MDAnalysis is importable here: run the script for real")
    except ImportError:
        pass
    monkeypatch.setattr(sys, "argv", ["pin_against_mdanalysis.py"])
    assert pin.main() == 2
    assert "cannot import" in capsys.readouterr().err
    # and the stand-in may never overwrite a real fixture
    monkeypatch.setattr(sys, "argv", ["pin", "--provider", "oracle"])
    assert pin.main() == 2


needs_fixture = pytest.mark.skipif(
    not os.path.exists(PINNED),
    reason="parity unpinned: no fixture produced by the real MDAnalysis yet "
           "(run tools/pin_against_mdanalysis.py where MDAnalysis imports)")


@needs_fixture
def test_oracle_reproduces_the_mdanalysis_fixture():
    z = np.load(PINNED)
    assert b"NOT MDAnalysis" not in z["provider"].tobytes()
    for name, case in pin.rdf_cases().items():
        np.testing.assert_array_equal(pin.oracle_counts(*case),
                                      z["rdf__%s__count" % name], err_msg=name)


@needs_fixture
@pytest.mark.gpu
def test_cuda_path_reproduces_the_mdanalysis_fixture():
    import warnings
    from pmda_b200.rdf import InterRDF
    from pmda_b200.universe import Universe
    z = np.load(PINNED)
    for name, (pos, dims, i1, i2, nbins, rng, excl) in pin.rdf_cases().items():
        u = Universe(pos, dims)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)
            r = InterRDF(u.atoms[i1], u.atoms[i2], nbins=nbins, range=rng,
                         exclusion_block=excl).run()
        np.testing.assert_array_equal(r.count, z["rdf__%s__count" % name],
                                      err_msg=name)


@pytest.mark.gpu
def test_cuda_path_matches_the_oracle_on_the_pinning_cases_and_flags_wraps():
    """until the fixture exists: GPU == oracle on the very cases the script
    would pin, and the product SAYS when it is on the unpinned branch (atoms
    outside the primary triclinic cell)"""
    import warnings
    from pmda_b200 import engine
    from pmda_b200.rdf import InterRDF
    from pmda_b200.universe import Universe
    engine._set_test_backend(None)
    for name, case in pin.rdf_cases().items():
        pos, dims, i1, i2, nbins, rng, excl = case
        u = Universe(pos, dims)
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            r = InterRDF(u.atoms[i1], u.atoms[i2], nbins=nbins, range=rng,
                         exclusion_block=excl).run()
        np.testing.assert_array_equal(r.count, pin.oracle_counts(*case),
                                      err_msg=name)
        flagged = [x for x in w if "outside the primary triclinic" in
                   str(x.message)]
        if "out_of_cell" in name and name.startswith("triclinic"):
            assert flagged, name
        elif name.startswith("golden"):
            assert not flagged, name
