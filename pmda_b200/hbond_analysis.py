"""Hydrogen bond analysis -- the operator interface of pmda.hbond_analysis
(pmda/hbond_analysis.py:43-560) over the B200 search kernel.

Per frame the reference runs (pmda/hbond_analysis.py:415-470)

    pairs, dist = capped_distance(donors.positions, acceptors.positions,
                                  max_cutoff=d_a_cutoff, min_cutoff=1.0, box=box)
    angles = rad2deg(calc_angles(D[pairs[:,0]], H[pairs[:,0]], A[pairs[:,1]], box))
    keep = angles > d_h_a_angle_cutoff

and appends ``[frame, donor, hydrogen, acceptor, distance, angle]`` rows.  Here
a whole batch of frames is searched by one ``rdfk_hbonds`` call (include/rdfk.h)
whose compacted output has the same rows in the same (frame, donor-hydrogen
pair, acceptor) order.

Differences from the reference, all on the topology side (out of scope,
SURVEY.md section 8): selections are the small language of
:class:`pmda_b200.universe.Universe` (``name ...`` / ``index ...``), AtomGroups
or index arrays; donor-hydrogen pairs come from ``d_h_cutoff`` (the reference's
PDB route) or from an explicit ``pairs=(donor_indices, hydrogen_indices)``
argument (the reference's bonded-topology route); atom ``ids`` are the 0-based
atom indices; the charge/mass based ``guess_*`` helpers need topology
attributes this Universe does not carry and raise.
"""
import numpy as np

from . import engine
from .parallel import ParallelAnalysisBase
from .universe import AtomGroup


def _as_indices(universe, sel, what):
    if sel is None:
        raise ValueError(
            "%s is required: this Universe carries no masses / charges to "
            "guess it from" % what)
    if isinstance(sel, str):
        return universe.select_atoms(sel).indices
    if isinstance(sel, AtomGroup) or hasattr(sel, "indices"):
        return np.asarray(sel.indices, dtype=np.intp)
    return np.asarray(sel, dtype=np.intp).reshape(-1)


class HydrogenBondAnalysis(ParallelAnalysisBase):
    """Perform an analysis of hydrogen bonds in a Universe
    (pmda/hbond_analysis.py:43-180).

    Parameters
    ----------
    universe : Universe
    donors_sel, hydrogens_sel, acceptors_sel : str, AtomGroup or indices
    d_h_cutoff : float, donor-hydrogen pairing distance [1.2]
    d_a_cutoff : float, donor-acceptor distance cut-off [3.0]
    d_h_a_angle_cutoff : float, D-H..A angle cut-off in degrees [150]
    update_selections : bool, re-pair donors and hydrogens every frame [True]
    pairs : (donor_indices, hydrogen_indices), optional -- explicit D-H pairs
        (what a bonded topology gives the reference); ``donors_sel`` /
        ``hydrogens_sel`` are then ignored.

    After :meth:`run`: ``hbonds`` float64 ``[n, 6]`` = frame, donor id,
    hydrogen id, acceptor id, distance, angle.
    """

    def __init__(self, universe, donors_sel=None, hydrogens_sel=None,
                 acceptors_sel=None, d_h_cutoff=1.2, d_a_cutoff=3.0,
                 d_h_a_angle_cutoff=150, update_selections=True, pairs=None):
        ag = universe.atoms
        super(HydrogenBondAnalysis, self).__init__(universe, (ag, ))
        self.donors_sel = donors_sel
        self.hydrogens_sel = hydrogens_sel
        self.acceptors_sel = acceptors_sel
        self.d_h_cutoff = d_h_cutoff
        self.d_a_cutoff = d_a_cutoff
        self.d_h_a_angle = d_h_a_angle_cutoff
        self.update_selections = update_selections
        self._pairs = None if pairs is None else (
            np.asarray(pairs[0], dtype=np.intp).reshape(-1),
            np.asarray(pairs[1], dtype=np.intp).reshape(-1))
        if self._pairs is not None and \
                len(self._pairs[0]) != len(self._pairs[1]):
            raise ValueError("pairs needs as many donors as hydrogens")

    # the charge / mass heuristics of pmda/hbond_analysis.py:183-347
    def guess_hydrogens(self, selection='all', max_mass=1.1, min_charge=0.3):
        raise NotImplementedError(
            "guess_hydrogens needs per-atom masses and charges; pass "
            "hydrogens_sel")

    def guess_donors(self, selection='all', max_charge=-0.5):
        raise NotImplementedError(
            "guess_donors needs per-atom charges; pass donors_sel")

    def guess_acceptors(self, selection='all', max_charge=-0.5):
        raise NotImplementedError(
            "guess_acceptors needs per-atom charges; pass acceptors_sel")

    # ------------------------------------------------------------------ GPU
    def _capped_pairs(self, donors, hydrogens, blocks, n_jobs):
        """capped_distance(donors, hydrogens, d_h_cutoff, box) for every frame
        of `blocks` (pmda/hbond_analysis.py:372-378) -> one
        (donor_indices, hydrogen_indices) per frame, in run order"""
        spec = {"donors": donors, "hydrogens": None, "acceptors": hydrogens,
                "max_cutoff": float(self.d_h_cutoff), "min_cutoff": None,
                "angle_cutoff": 0.0}
        raw = engine.run_hbonds_blocks(
            self._trajectory, self._universe.atoms.n_atoms, spec, blocks,
            n_jobs)
        per_frame, pos = [], 0
        for blk, el in zip(blocks, raw):
            frame, k, acc = el[0][:3]
            lo = np.searchsorted(frame, pos + np.arange(len(blk) + 1))
            for i in range(len(blk)):
                sl = slice(lo[i], lo[i + 1])
                per_frame.append((donors[k[sl]], hydrogens[acc[sl]]))
            pos += len(blk)
        return per_frame

    def _get_dh_pairs(self, frame=None, n_jobs=1):
        """donor / hydrogen index arrays that zip into D-H pairs
        (pmda/hbond_analysis.py:349-383) at `frame` (default: current)"""
        if self._pairs is not None:
            return self._pairs
        u = self._universe
        hydrogens = _as_indices(u, self.hydrogens_sel, "hydrogens_sel")
        if self.donors_sel is None:
            raise ValueError(
                'Cannot assign donor-hydrogen pairs via topology as no '
                'bonded information is present. Please either pass '
                'pairs=(donors, hydrogens) or set '
                'HydrogenBondAnalysis.donors_sel so that a distance cutoff '
                'can be used.')
        donors = _as_indices(u, self.donors_sel, "donors_sel")
        if frame is None:
            frame = self._trajectory.ts.frame
        return self._capped_pairs(donors, hydrogens,
                                  [range(frame, frame + 1)], n_jobs)[0]

    def _prepare(self):
        traj = self._trajectory
        self.hbonds = []
        self.frames = np.arange(self.start, self.stop, self.step)
        dt = getattr(traj, "dt", 1.0)
        t0 = getattr(traj[0], "time", 0.0) if len(traj) else 0.0
        self.timesteps = (self.frames * dt) + t0
        self._acceptors_ids = _as_indices(self._universe, self.acceptors_sel,
                                          "acceptors_sel")
        # The reference pairs donors with hydrogens on a Universe it has just
        # built and moved to frame 0 (pmda/hbond_analysis.py:398-408): always
        # frame 0, never the frame the caller's reader happens to stand on.
        # With update_selections the pairs are re-derived for every frame of
        # the run (_device_blocks), so this search would be thrown away.
        if self._repairs_every_frame():
            self._donors_ids = self._hydrogens_ids = None
        else:
            self._donors_ids, self._hydrogens_ids = self._get_dh_pairs(frame=0)

    def _spec(self, donors, hydrogens):
        return {"donors": donors, "hydrogens": hydrogens,
                "acceptors": self._acceptors_ids,
                "max_cutoff": float(self.d_a_cutoff),
                # an atom cannot form a hydrogen bond with itself
                # (pmda/hbond_analysis.py:430)
                "min_cutoff": 1.0,
                "angle_cutoff": float(self.d_h_a_angle)}

    def _rows(self, spec, blocks, n_jobs):
        """engine result -> the reference's six columns, per block"""
        raw = engine.run_hbonds_blocks(
            self._trajectory, self._universe.atoms.n_atoms, spec, blocks,
            n_jobs)
        don, hyd, acc = spec["donors"], spec["hydrogens"], spec["acceptors"]
        out, pos = [], 0
        for blk, el in zip(blocks, raw):
            frame, k, j, dist, ang = el[0]
            frames = np.fromiter(blk, dtype=np.float64, count=len(blk))
            tab = np.empty((len(k), 6), dtype=np.float64)
            if len(k):
                tab[:, 0] = frames[frame - pos]
            tab[:, 1] = don[k]
            tab[:, 2] = hyd[k]
            tab[:, 3] = acc[j]
            tab[:, 4] = dist
            tab[:, 5] = ang
            out.append((tab,) + tuple(el[1:]))
            pos += len(blk)
        return out

    def _repairs_every_frame(self):
        return (self.update_selections and self._pairs is None
                and self.donors_sel is not None)

    def _device_blocks(self, blocks, n_jobs):
        repair = self._repairs_every_frame()
        if not repair:
            return self._rows(self._spec(self._donors_ids,
                                         self._hydrogens_ids), blocks, n_jobs)
        # update_selections: the D-H pairs are re-derived in every frame
        # (pmda/hbond_analysis.py:421-423).  One capped search over the run
        # gives all of them; frames that share a pair list run together.
        u = self._universe
        donors = _as_indices(u, self.donors_sel, "donors_sel")
        hydrogens = _as_indices(u, self.hydrogens_sel, "hydrogens_sel")
        per_frame = self._capped_pairs(donors, hydrogens, blocks, n_jobs)
        out, pos = [], 0
        for blk in blocks:
            n = len(blk)
            lists = per_frame[pos:pos + n]
            pieces, extras = [], []
            lo = 0
            while lo < n:
                hi = lo + 1
                while hi < n and np.array_equal(lists[hi][0], lists[lo][0]) \
                        and np.array_equal(lists[hi][1], lists[lo][1]):
                    hi += 1
                el = self._rows(self._spec(lists[lo][0], lists[lo][1]),
                                [blk[lo:hi]], n_jobs)[0]
                pieces.append(el[0])
                extras.append(el[1:])
                lo = hi
            tab = np.concatenate(pieces) if pieces \
                else np.empty((0, 6), dtype=np.float64)
            out.append((tab,
                        np.concatenate([e[0] for e in extras]),
                        np.concatenate([e[1] for e in extras]),
                        0.0, min(e[3] for e in extras),
                        sum(e[4] for e in extras), sum(e[5] for e in extras)))
            pos += n
        return out

    def _single_frame(self, ts, atomgroups):
        """One frame on the GPU (API parity with
        pmda/hbond_analysis.py:415-470)."""
        return self._device_blocks([range(ts.frame, ts.frame + 1)], 1)[0][0]

    def _conclude(self):
        blocks = [np.asarray(r, dtype=np.float64).reshape(-1, 6)
                  for r in self._results if len(r)]
        self.hbonds = np.vstack(blocks) if blocks \
            else np.empty((0, 6), dtype=np.float64)

    # ---------------------------------------------------------- summaries
    def count_by_time(self):
        """Number of hydrogen bonds per analysed frame
        (pmda/hbond_analysis.py:475-494)."""
        indices, tmp_counts = np.unique(self.hbonds[:, 0], axis=0,
                                        return_counts=True)
        indices = (indices - self.start) / self.step
        counts = np.zeros_like(self.frames)
        counts[indices.astype(np.intp)] = tmp_counts
        return counts

    def count_by_type(self):
        """Counts per (donor resname:type, acceptor resname:type)
        (pmda/hbond_analysis.py:496-528); needs ``universe.resnames`` and
        ``universe.types``."""
        u = self._universe
        resnames = getattr(u, "resnames", None)
        types = getattr(u, "types", None)
        if resnames is None or types is None:
            raise ValueError("count_by_type needs universe.resnames and "
                             "universe.types")
        d = self.hbonds[:, 1].astype(np.intp)
        a = self.hbonds[:, 3].astype(np.intp)
        tmp_hbonds = np.array([np.asarray(resnames)[d], np.asarray(types)[d],
                               np.asarray(resnames)[a], np.asarray(types)[a]],
                              dtype=str).T
        hbond_type, type_counts = np.unique(tmp_hbonds, axis=0,
                                            return_counts=True)
        hbond_type_list = []
        for hb_type, hb_count in zip(hbond_type, type_counts):
            hbond_type_list.append(
                [":".join(hb_type[:2]), ":".join(hb_type[2:4]), hb_count])
        return np.array(hbond_type_list)

    def count_by_ids(self):
        """Counts per unique (donor, hydrogen, acceptor) triple, most frequent
        first (pmda/hbond_analysis.py:530-560)."""
        tmp_hbonds = self.hbonds[:, 1:4].astype(np.intp)
        hbond_ids, ids_counts = np.unique(tmp_hbonds, axis=0,
                                          return_counts=True)
        unique_hbonds = np.concatenate((hbond_ids, ids_counts[:, None]),
                                       axis=1)
        return unique_hbonds[unique_hbonds[:, 3].argsort()[::-1]]
