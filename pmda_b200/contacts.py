"""Native-contacts analysis on B200 -- drop-in for :mod:`pmda.contacts`
(the sibling analysis on the same distance arithmetic).

Same constructor, ``run(start, stop, step, n_jobs, n_blocks)`` and
``timeseries`` as pmda/contacts.py:195-292.  The per-frame work of
``Contacts._single_frame`` (pmda/contacts.py:270-289) --
``distance_array(grA.positions, grB.positions)`` restricted to the contacts of
the reference conformation, then ``fraction_contacts(r, r0, **kwargs)`` --
runs in ``rdfk_contacts`` (include/rdfk.h): fused count / sum kernels for
``hard_cut``, ``soft_cut`` and :func:`radius_cut_q`; for any other callable
(the "Contacts API") the kernel emits the distances and the callable is applied
on the host, frame by frame, exactly like the reference does.  The reference
distances ``r0`` are computed once at construction on the host, as in the
reference.  There is no CPU implementation of the per-frame distances here.
"""
import numpy as np

from . import _cabi, engine
from .parallel import ParallelAnalysisBase
from .universe import Universe


# -- restated from MDAnalysis.analysis.contacts (absent from /root/reference;
#    call sites pmda/contacts.py:188-189, 236-247, 285) ----------------------
def hard_cut_q(r, cutoff):
    """fraction of contacts with ``r <= cutoff`` (element-wise)"""
    r = np.asarray(r)
    cutoff = np.asarray(cutoff)
    y = r <= cutoff
    return y.sum() / r.size


def soft_cut_q(r, r0, beta=5.0, lambda_constant=1.8):
    """soft cut-off fraction of native contacts (Best & Hummer 2013)"""
    r = np.asarray(r)
    r0 = np.asarray(r0)
    result = 1 / (1 + np.exp(beta * (r - lambda_constant * r0)))
    return result.sum() / len(r0)


def radius_cut_q(r, r0, radius):
    """fraction of contacts closer than ``radius``"""
    y = r <= radius
    return y.sum() / r.size


def contact_matrix(d, radius, out=None):
    """boolean matrix of the distances ``d <= radius``"""
    if out is not None:
        out[:] = d <= radius
        return out
    return d <= radius


def distance_array(reference, configuration):
    """``MDAnalysis.lib.distances.distance_array`` without a box, restated:
    float32 difference, float64 squares summed x, y, z, float64 sqrt
    (_calc_distance_array).  Host side, used once per reference pair at
    construction (pmda/contacts.py:252-260), never per frame."""
    ref = np.ascontiguousarray(reference, dtype=np.float32)
    conf = np.ascontiguousarray(configuration, dtype=np.float32)
    d = (conf[None, :, :] - ref[:, None, :]).astype(np.float64)
    d *= d
    return np.sqrt((d[..., 0] + d[..., 1]) + d[..., 2])


class Contacts(ParallelAnalysisBase):
    """Fraction of native contacts *Q* (pmda/contacts.py:195-292).

    Parameters
    ----------
    mobiles : tuple(AtomGroup, AtomGroup) -- the two contacting groups
    refs : tuple(AtomGroup, AtomGroup) or list of such tuples -- the groups in
        their reference conformation
    method : ``'hard_cut'``, ``'soft_cut'`` or ``func(r, r0, **kwargs)``
    radius : float, distance below which a reference contact exists [4.5]
    kwargs : dict of extra arguments for `method`

    ``timeseries`` after :meth:`run`: ``[n_frames, 1 + n_refs]`` (frame
    number, then one *Q* per reference pair).
    """

    def __init__(self, mobiles, refs, method="hard_cut", radius=4.5,
                 kwargs=None):
        universe = mobiles[0].universe
        super(Contacts, self).__init__(universe, mobiles)

        if method == 'hard_cut':
            self.fraction_contacts = hard_cut_q
        elif method == 'soft_cut':
            self.fraction_contacts = soft_cut_q
        else:
            if not callable(method):
                raise ValueError("method has to be callable")
            self.fraction_contacts = method

        # contacts formed in reference
        self.r0 = []
        self.initial_contacts = []
        if isinstance(refs[0], type(mobiles[0])) or hasattr(refs[0], "positions"):
            refs = [refs]
        for refA, refB in refs:
            self.r0.append(distance_array(refA.positions, refB.positions))
            self.initial_contacts.append(contact_matrix(self.r0[-1], radius))

        self.fraction_kwargs = kwargs if kwargs is not None else {}
        self._radius = radius

    def _prepare(self):
        self.timeseries = None

    # ------------------------------------------------------------------ GPU
    def _spec(self):
        """the contact pairs of every reference as index lists into a frame"""
        ia_all, ib_all = self._indices
        fk = dict(self.fraction_kwargs)
        fn = self.fraction_contacts
        method = _cabi.CONTACT_DIST
        radius, beta, lam = 0.0, 5.0, 1.8
        if fn is hard_cut_q and not fk:
            method = _cabi.CONTACT_HARD
        elif fn is radius_cut_q and set(fk) == {"radius"}:
            method, radius = _cabi.CONTACT_RADIUS, float(fk["radius"])
        elif fn is soft_cut_q and set(fk) <= {"beta", "lambda_constant"}:
            method = _cabi.CONTACT_SOFT
            beta = float(fk.get("beta", 5.0))
            lam = float(fk.get("lambda_constant", 1.8))
        refs, col = [], 0
        for init, r0 in zip(self.initial_contacts, self.r0):
            if init.shape != (len(ia_all), len(ib_all)):
                raise ValueError("reference groups and mobile groups differ "
                                 "in size")
            a, b = np.nonzero(init)            # row-major, like d[mask]
            refs.append({"ia": ia_all[a], "ib": ib_all[b],
                         "r0": np.ascontiguousarray(r0[init], np.float64),
                         "col": col})
            col += len(a) if method == _cabi.CONTACT_DIST else 1
        spec = {"refs": refs, "method": method, "radius": radius,
                "beta": beta, "lambda": lam, "width": max(col, 1)}
        if method == _cabi.CONTACT_DIST:
            # the user's callable runs on every batch of distances as it
            # comes off the device (pmda/contacts.py:283-286, per frame)
            def reduce(dist):
                y = np.empty((len(dist), len(refs)))
                for i, ref in enumerate(refs):
                    lo, n = ref["col"], len(ref["ia"])
                    for k in range(len(dist)):
                        y[k, i] = self.fraction_contacts(
                            dist[k, lo:lo + n], ref["r0"],
                            **self.fraction_kwargs)
                return y
            spec["reduce"] = reduce
        return spec

    def _device_blocks(self, blocks, n_jobs):
        spec = self._spec()
        raw = engine.run_contacts_blocks(
            self._trajectory, self._universe.atoms.n_atoms, spec, blocks,
            n_jobs)
        out = []
        for blk, el in zip(blocks, raw):
            series = el[0]
            y = np.empty((len(blk), len(self.r0) + 1))
            y[:, 0] = np.fromiter(blk, dtype=np.float64, count=len(blk))
            for i, ref in enumerate(spec["refs"]):
                n = len(ref["ia"])
                if spec["method"] == _cabi.CONTACT_DIST:
                    y[:, i + 1] = series[:, i]
                else:
                    with np.errstate(divide='ignore', invalid='ignore'):
                        y[:, i + 1] = series[:, ref["col"]] / n
            if y.shape[1] == 1:
                y = y[:, 0]
            out.append((y,) + tuple(el[1:]))
        return out

    def _single_frame(self, ts, atomgroups):
        """One frame on the GPU (API parity with pmda/contacts.py:270-289)."""
        return self._device_blocks([range(ts.frame, ts.frame + 1)], 1)[0][0][0]

    def _conclude(self):
        self.timeseries = np.concatenate(self._results)


def q1q2(atomgroup, radius=4.5):
    """q1-q2 analysis (pmda/contacts.py:292-335): native contacts relative to
    the first and to the last frame of the trajectory."""
    u_orig = atomgroup.universe
    indices = atomgroup.indices

    def create_refs(frame):
        """stand-alone AtomGroups of the selection at `frame`"""
        ts = u_orig.trajectory[frame]
        dims = None if ts.dimensions is None else np.array(ts.dimensions)
        u = Universe(np.array(ts.positions, dtype=np.float32)[None], dims)
        return [u.atoms[indices], u.atoms[indices]]

    first_frame_refs = create_refs(0)
    last_frame_refs = create_refs(-1)
    return Contacts(
        (atomgroup, atomgroup), (first_frame_refs, last_frame_refs),
        radius=radius,
        method=radius_cut_q,
        kwargs={'radius': radius})
