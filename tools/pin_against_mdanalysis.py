#!/usr/bin/env python
"""Pin the oracle against the REAL MDAnalysis, wherever it can be imported.

    python tools/pin_against_mdanalysis.py             # needs `import MDAnalysis`
    python tools/pin_against_mdanalysis.py --provider oracle --out /tmp/x.npz
                                                       # dry run of the machinery

Why: the arithmetic under pmda's InterRDF / InterRDF_s / HydrogenBondAnalysis
lives in MDAnalysis (`pmda/rdf.py:33`, pinned only as `MDAnalysis>=1.0.0` in
`setup.py:51`).  MDAnalysis is neither under /root/reference nor installable in
the build container, so `oracle/` restates it from the published algorithm and
every parity claim of this repository is "GPU == oracle".  This script closes
the last link on any machine that has MDAnalysis:

  1. it regenerates the seeded inputs of every committed golden case
     (tests/golden/make_oracle_golden.py, make_hbond_golden.py) plus cases
     that only matter for this question -- triclinic coordinates far OUTSIDE
     the primary cell (the `_triclinic_pbc` pre-wrap, the least certain part
     of the restatement), atoms sitting exactly on bin edges, a two-atom
     distance sweep across an edge, `calc_angles` triples;
  2. it computes each case with the reference's own calls, written as the
     reference writes them (`capped_distance(g1, g2, rmax, box=dimensions)`,
     exclusion mask, `np.histogram` -- pmda/rdf.py:120-137; `distance_array`;
     `calc_angles` -- pmda/hbond_analysis.py:415-470);
  3. it computes the same with oracle/ (C and numpy restatements) and DIFFS:
     integer counts must be identical, distances and angles bit-equal;
  4. it writes the vectors with the MDAnalysis version into
     tests/golden/mdanalysis_pinned.npz.  Once that file is committed,
     tests/test_pinning.py makes the oracle (CPU) and the CUDA path (GPU)
     reproduce it on every run, and DESIGN.md section 6 may say "pinned".

Exit status: 0 = every case identical, 1 = differences (printed), 2 = the
provider could not be imported.

`--provider oracle` swaps MDAnalysis for the numpy restatement: the diff is
then trivially empty, but every line of the case generation, the comparison
and the fixture writer runs -- that is what tests/test_pinning.py exercises
on the CPU runner, so the script cannot rot unnoticed.
"""
import argparse
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)

DEFAULT_OUT = os.path.join(ROOT, "tests", "golden", "mdanalysis_pinned.npz")


# --------------------------------------------------------------------------
# providers: who answers "what does the reference compute?"
# --------------------------------------------------------------------------
class MDAnalysisProvider(object):
    """the real thing"""

    def __init__(self):
        import MDAnalysis
        from MDAnalysis.lib import distances
        self.name = "MDAnalysis " + MDAnalysis.__version__
        self._d = distances

    def capped_distance(self, ref, conf, max_cutoff, box, method=None):
        # pmda/rdf.py:123-126 lets capped_distance pick its method; the north
        # star pins the semantics to the brute-force (distance_array) branch,
        # so both are recorded
        kw = {} if method is None else {"method": method}
        pairs, dist = self._d.capped_distance(ref, conf, max_cutoff, box=box,
                                              **kw)
        return np.asarray(pairs).reshape(-1, 2), np.asarray(dist, np.float64)

    def distance_array(self, ref, conf, box):
        return self._d.distance_array(ref, conf, box=box)

    def calc_angles(self, a, b, c, box):
        return self._d.calc_angles(a, b, c, box=box)


class OracleProvider(object):
    """stand-in for dry runs / the CPU test of this script: the numpy
    restatement answers instead of MDAnalysis"""

    def __init__(self):
        from oracle import numpy_oracle
        self.name = "oracle stand-in (NOT MDAnalysis)"
        self._o = numpy_oracle

    def capped_distance(self, ref, conf, max_cutoff, box, method=None):
        return self._o.capped_distance(ref, conf, max_cutoff, box=box)

    def distance_array(self, ref, conf, box):
        return self._o.distance_array(ref, conf, box=box)

    def calc_angles(self, a, b, c, box):
        return self._o.calc_angles(a, b, c, box=box)


def get_provider(name):
    if name == "oracle":
        return OracleProvider()
    return MDAnalysisProvider()


# --------------------------------------------------------------------------
# cases
# --------------------------------------------------------------------------
def _seed(name):
    return int(hashlib.sha256(("pin" + name).encode()).hexdigest()[:8], 16)


def rdf_cases():
    """name -> (pos [F,N,3] f32, dims or None, i1, i2, nbins, range, excl)"""
    import make_oracle_golden as mk
    from pmda_b200 import mdamath, synthetic
    cases = {}
    for name in sorted(mk.CASES):
        cases["golden_" + name] = mk.make_inputs(name)

    # triclinic coordinates OUTSIDE the primary cell: unwrapped molecules that
    # drifted by whole lattice vectors, by fractions, to negative coordinates
    dims = [24.0, 26.0, 28.0, 70.0, 80.0, 65.0]
    gen = np.random.default_rng(_seed("tri_out"))
    pos = synthetic.uniform_triclinic(2, 700, dims, gen)
    m = mdamath.triclinic_vectors(np.asarray(dims, np.float32),
                                  dtype=np.float64)
    shift = gen.integers(-3, 4, (2, 700, 3)).astype(np.float64)
    far = (pos.astype(np.float64) + shift @ m).astype(np.float32)
    ix = np.arange(700)
    cases["triclinic_out_of_cell_lattice_shifts"] = (
        far, dims, ix, ix, 50, (0.0, 10.0), None)
    frac = gen.random((2, 600, 3)) * 5.0 - 2.5       # fractional in [-2.5, 2.5)
    cases["triclinic_out_of_cell_uniform"] = (
        (frac @ m).astype(np.float32), dims, np.arange(600), np.arange(600),
        50, (0.0, 10.0), None)
    # exactly on the cell faces (fractional 0 and 1 after float32 rounding)
    edge = gen.integers(0, 2, (1, 300, 3)).astype(np.float64)
    edge[:, 150:] = gen.random((1, 150, 3))
    cases["triclinic_on_the_faces"] = (
        (edge @ m).astype(np.float32), dims, np.arange(300), np.arange(300),
        40, (0.0, 11.0), None)
    # orthorhombic, coordinates outside [0, L) (distance_array does not wrap
    # for orthorhombic boxes; the minimum image must still be right)
    L = np.array([21.0, 23.0, 25.0])
    out = ((gen.random((2, 600, 3)) * 6.0 - 3.0) * L).astype(np.float32)
    cases["ortho_out_of_cell"] = (
        out, [21.0, 23.0, 25.0, 90, 90, 90], np.arange(600), np.arange(600),
        60, (0.0, 10.0), None)
    # simple-cubic lattice whose neighbour distances sit EXACTLY on bin edges
    # (a = 2 * dr): pins `edges[i] <= d` and the float32 inverse box
    k = 8
    a = 2.0
    g = np.stack(np.meshgrid(*[np.arange(k)] * 3, indexing="ij"),
                 -1).reshape(-1, 3)
    lat = (g * a).astype(np.float32)[None]
    cases["lattice_on_bin_edges"] = (
        lat, [k * a, k * a, k * a, 90, 90, 90], np.arange(k ** 3),
        np.arange(k ** 3), 7, (0.0, 7.0), None)
    return cases


def sweep_case():
    """two atoms, the second one moved through a bin edge in float32 ulps:
    distances (bit patterns) and bins"""
    xs = np.float32(3.2) + np.arange(-40, 41).astype(np.float32) * \
        np.spacing(np.float32(3.2))
    ref = np.array([[1.25, 2.5, 3.75]], dtype=np.float32)
    conf = np.stack([ref[0, 0] + xs,
                     np.full_like(xs, ref[0, 1] + np.float32(1.7)),
                     np.full_like(xs, ref[0, 2] - np.float32(0.9))],
                    axis=1).astype(np.float32)
    return ref, conf


def angle_case():
    gen = np.random.default_rng(_seed("angles"))
    a, b, c = ((gen.random((400, 3)) * 18.0).astype(np.float32)
               for _ in range(3))
    return a, b, c


BOXES = {
    "none": None,
    "ortho": np.array([16.0, 17.0, 18.0, 90, 90, 90], dtype=np.float32),
    "tri": np.array([16.0, 17.0, 18.0, 75.0, 80.0, 70.0], dtype=np.float32),
}


# --------------------------------------------------------------------------
# the reference's per-frame body, on a provider
# --------------------------------------------------------------------------
def interrdf_counts(provider, pos, dims, i1, i2, nbins, rng, excl,
                    method=None):
    """pmda/rdf.py:120-137 accumulated over the frames (:180-190)"""
    count = np.zeros(nbins, dtype=np.int64)
    box = None if dims is None else np.asarray(dims, dtype=np.float32)
    for f in range(len(pos)):
        g1, g2 = pos[f][i1], pos[f][i2]
        pairs, dist = provider.capped_distance(g1, g2, rng[1], box, method)
        if excl is not None:
            idxA, idxB = pairs[:, 0] // excl[0], pairs[:, 1] // excl[1]
            dist = dist[np.where(idxA != idxB)[0]]
        count += np.histogram(dist, bins=nbins, range=rng)[0]
    return count


def oracle_counts(pos, dims, i1, i2, nbins, rng, excl):
    from oracle import c_oracle
    count = np.zeros(nbins, dtype=np.int64)
    for f in range(len(pos)):
        c_oracle.rdf_single_frame(pos[f][i1], pos[f][i2], dims, nbins, rng,
                                  excl, count=count, nthreads=4)
    return count


def run(provider, out_path, verbose=True, only=None):
    """-> number of differing cases; writes the fixture.  `only`: substrings
    selecting RDF cases (the CPU test runs a quick subset)"""
    from oracle import numpy_oracle
    say = print if verbose else (lambda *a, **k: None)
    fixture = {"provider": np.frombuffer(provider.name.encode(), np.uint8)}
    bad = 0

    for name, (pos, dims, i1, i2, nbins, rng, excl) in rdf_cases().items():
        if only is not None and not any(o in name for o in only):
            continue
        digest = hashlib.sha256(np.ascontiguousarray(pos).tobytes()).hexdigest()
        ref = interrdf_counts(provider, pos, dims, i1, i2, nbins, rng, excl)
        # the branch the north star pins the semantics to
        ref_bf = interrdf_counts(provider, pos, dims, i1, i2, nbins, rng, excl,
                                 method="bruteforce")
        mine = oracle_counts(pos, dims, i1, i2, nbins, rng, excl)
        ok = np.array_equal(ref_bf, mine)
        same_dispatch = np.array_equal(ref, ref_bf)
        bad += 0 if ok else 1
        say("%-40s %s  sum=%d%s" % (
            name, "identical" if ok else "DIFFERENT", int(ref_bf.sum()),
            "" if same_dispatch else
            "   (capped_distance's default method gives other counts than "
            "its brute-force branch: %d bins)" % int((ref != ref_bf).sum())))
        if not ok:
            w = np.flatnonzero(ref_bf != mine)
            say("    bins %s: reference %s, oracle %s" %
                (w[:8], ref_bf[w][:8], mine[w][:8]))
        fixture["rdf__%s__count" % name] = ref_bf
        fixture["rdf__%s__count_default_method" % name] = ref
        fixture["rdf__%s__sha256" % name] = np.frombuffer(digest.encode(),
                                                          np.uint8)

    ref_a, conf = sweep_case()
    for bname, box in BOXES.items():
        d_ref = np.asarray(provider.distance_array(ref_a, conf, box),
                           np.float64)
        d_me = numpy_oracle.distance_array(ref_a, conf, box=box)
        ok = np.array_equal(d_ref.view(np.uint64), d_me.view(np.uint64))
        bad += 0 if ok else 1
        say("%-40s %s" % ("distance sweep, box=" + bname,
                          "bit-identical" if ok else "DIFFERENT (max |diff| "
                          "%.3g)" % np.abs(d_ref - d_me).max()))
        fixture["sweep__%s__dist" % bname] = d_ref

    a, b, c = angle_case()
    for bname, box in BOXES.items():
        g_ref = np.asarray(provider.calc_angles(a, b, c, box), np.float64)
        g_me = numpy_oracle.calc_angles(a, b, c, box=box)
        # calc_angles ends in atan2: libm may differ by an ulp between builds
        ok = np.allclose(g_ref, g_me, rtol=0, atol=4e-16 * np.pi)
        bad += 0 if ok else 1
        say("%-40s %s" % ("calc_angles, box=" + bname,
                          "identical (<= 2 ulp)" if ok else
                          "DIFFERENT (max |diff| %.3g rad)"
                          % np.abs(g_ref - g_me).max()))
        fixture["angles__%s" % bname] = g_ref

    fixture["n_different"] = np.array(bad)
    np.savez(out_path, **fixture)
    say("%d case(s) differ; vectors of %s written to %s"
        % (bad, provider.name, out_path))
    return bad


def main():
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--provider", default="mdanalysis",
                    choices=["mdanalysis", "oracle"])
    ap.add_argument("--out", default=None)
    ap.add_argument("--only", nargs="*", default=None,
                    help="substrings selecting RDF cases (default: all)")
    args = ap.parse_args()
    try:
        provider = get_provider(args.provider)
    except ImportError as exc:
        print("cannot import the provider: %s\n(MDAnalysis is not installable "
              "in the build container; run this where it is)" % exc,
              file=sys.stderr)
        return 2
    out = args.out
    if out is None:
        if args.provider == "oracle":
            print("--provider oracle needs --out: the stand-in must never "
                  "overwrite real MDAnalysis vectors", file=sys.stderr)
            return 2
        out = DEFAULT_OUT
    return 1 if run(provider, out, only=args.only) else 0


if __name__ == "__main__":
    sys.exit(main())
