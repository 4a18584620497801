"""Generate the golden fixtures in this directory from the REFERENCE's own code.

Run in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

What can be pinned against the reference itself:
  * balanced_slices.json -- pmda/util.py:make_balanced_slices, imported straight
    from /root/reference/pmda/util.py (pure numpy, importable on its own).
  * conclude_rdf.npz / conclude_rdf_s.npz -- InterRDF._conclude / .cdf and
    InterRDF_s._conclude / .cdf (pmda/rdf.py:139-178, 311-360) executed from the
    reference's rdf.py.  pmda/rdf.py imports MDAnalysis and dask at module
    level; neither is installed, so inert stub modules are registered first --
    _conclude/cdf touch neither.  The per-frame path (_single_frame) cannot be
    executed: it needs MDAnalysis' compiled distance code, and it builds a
    ragged np.array that numpy >= 1.24 rejects.
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def load_ref_util():
    spec = importlib.util.spec_from_file_location(
        "ref_pmda_util", os.path.join(REF, "pmda", "util.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_ref_rdf():
    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m
    mda = stub("MDAnalysis")
    lib = stub("MDAnalysis.lib")
    dist = stub("MDAnalysis.lib.distances")
    mda.lib = lib
    lib.distances = dist
    dask = stub("dask")
    dd = stub("dask.delayed", delayed=lambda *a, **k: None)
    ddist = stub("dask.distributed")
    dask.delayed = dd
    dask.distributed = ddist
    sys.path.insert(0, REF)
    import pmda.rdf as ref_rdf          # noqa: E402  (the real reference file)
    return ref_rdf


def slices_to_json(sl):
    return [[None if v is None else int(v) for v in (s.start, s.stop, s.step)]
            for s in sl]


def make_balanced_slices_fixture():
    util = load_ref_util()
    cases = []
    for n_frames in (0, 1, 5, 10, 11, 98, 100, 256, 1000):
        for n_blocks in (1, 2, 3, 4, 5, 7, 10, 11, 64):
            for start, stop, step in ((None, None, None), (0, None, 1),
                                      (1, None, None), (10, None, 2),
                                      (0, 100, 3), (3, 256, 7), (5, 50, 5)):
                if stop is not None:
                    nf = len(range(start or 0, stop, step or 1))
                else:
                    nf = n_frames
                try:
                    out = util.make_balanced_slices(nf, n_blocks, start=start,
                                                    stop=stop, step=step)
                    res = {"slices": slices_to_json(out)}
                except ValueError as exc:
                    res = {"error": str(exc)}
                cases.append({"n_frames": nf, "n_blocks": n_blocks,
                              "start": start, "stop": stop, "step": step,
                              **res})
    with open(os.path.join(HERE, "balanced_slices.json"), "w") as fh:
        json.dump(cases, fh)
    print("balanced_slices.json:", len(cases), "cases")


def make_conclude_fixtures():
    ref = load_ref_rdf()
    rng = np.random.default_rng(20261019)
    out = {}
    case = 0
    for nbins, rrange, nA, nB, excl, n_blocks, n_frames in (
            (75, (0.0, 15.0), 3000, 3000, None, 4, 100),
            (200, (0.0, 15.0), 100000, 100000, None, 3, 17),
            (75, (1.0, 13.0), 700, 2300, (3, 4), 2, 9),
            (412, (0.0, 12.0), 1400, 1400, (7, 7), 1, 5)):
        obj = object.__new__(ref.InterRDF)
        obj.nA, obj.nB = nA, nB
        obj._exclusion_block = excl
        obj.rdf_settings = {"bins": nbins, "range": rrange}
        edges = np.histogram([-1], bins=nbins, range=rrange)[1]
        obj.edges = edges
        obj.bins = 0.5 * (edges[:-1] + edges[1:])
        obj.n_frames = n_frames
        res = np.empty((n_blocks, 2), dtype=object)
        counts = rng.integers(0, 10 ** 9, size=(n_blocks, nbins),
                              dtype=np.int64)
        vols = rng.uniform(8e4, 1e6, size=n_blocks)
        for b in range(n_blocks):
            res[b, 0] = counts[b]
            res[b, 1] = np.array(vols[b], dtype=np.float64)
        obj._results = res
        obj._conclude()
        pre = "c%d_" % case
        out[pre + "nbins"] = nbins
        out[pre + "range"] = np.array(rrange)
        out[pre + "nA"], out[pre + "nB"] = nA, nB
        out[pre + "excl"] = np.array(excl if excl else (0, 0))
        out[pre + "n_frames"] = n_frames
        out[pre + "block_counts"] = counts
        out[pre + "block_volumes"] = vols
        out[pre + "count"] = obj.count
        out[pre + "volume"] = obj.volume
        out[pre + "rdf"] = obj.rdf
        out[pre + "cdf"] = obj.cdf
        out[pre + "edges"] = obj.edges
        out[pre + "bins"] = obj.bins
        case += 1
    out["n_cases"] = case
    np.savez(os.path.join(HERE, "conclude_rdf.npz"), **out)
    print("conclude_rdf.npz:", case, "cases")

    # InterRDF_s: two site pairs with different shapes, density on / off
    out = {}
    case = 0
    for density in (True, False):
        nbins, rrange, n_blocks, n_frames = 75, (0.0, 15.0), 3, 11
        shapes = [(1, 2), (2, 3)]
        obj = object.__new__(ref.InterRDF_s)
        obj._density = density
        obj.n = len(shapes)
        obj.rdf_settings = {"bins": nbins, "range": rrange}
        obj.ag_shape = [list(s) for s in shapes]
        edges = np.histogram([-1], bins=nbins, range=rrange)[1]
        obj.edges = edges
        obj.bins = 0.5 * (edges[:-1] + edges[1:])
        obj.n_frames = n_frames
        res = np.empty((n_blocks, 2), dtype=object)
        vols = rng.uniform(8e4, 1e6, size=n_blocks)
        blocks = []
        for b in range(n_blocks):
            per = np.empty(len(shapes), dtype=object)
            for s, (n1, n2) in enumerate(shapes):
                per[s] = rng.integers(0, 5, size=(n1, n2, nbins)).astype(
                    np.float64)
            res[b, 0] = per
            res[b, 1] = np.array(vols[b], dtype=np.float64)
            blocks.append(per)
        obj._results = res
        obj._conclude()
        pre = "c%d_" % case
        out[pre + "density"] = density
        out[pre + "n_frames"] = n_frames
        out[pre + "block_volumes"] = vols
        for b in range(n_blocks):
            for s in range(len(shapes)):
                out[pre + "block%d_site%d" % (b, s)] = blocks[b][s]
        for s in range(len(shapes)):
            out[pre + "count%d" % s] = obj.count[s]
            out[pre + "rdf%d" % s] = obj.rdf[s]
            out[pre + "cdf%d" % s] = obj.cdf[s]
        out[pre + "volume"] = obj.volume
        case += 1
    out["n_cases"] = case
    np.savez(os.path.join(HERE, "conclude_rdf_s.npz"), **out)
    print("conclude_rdf_s.npz:", case, "cases")


if __name__ == "__main__":
    make_balanced_slices_fixture()
    make_conclude_fixtures()
