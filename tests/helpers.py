"""Shared helpers for the parity tests: run the ORACLE over a Universe."""
import numpy as np

from oracle import c_oracle


def oracle_interrdf(u, g1, g2, nbins=75, rng=(0.0, 15.0), exclusion_block=None,
                    frames=None, nthreads=8):
    """counts int64[nbins], summed volume, n_frames -- the reference's
    _single_frame + _reduce over `frames` (pmda/rdf.py:120-137, 180-190)."""
    traj = u.trajectory
    frames = range(traj.n_frames) if frames is None else frames
    count = np.zeros(nbins, dtype=np.int64)
    volume = 0.0
    for i in frames:
        ts = traj[i]
        _, v = c_oracle.rdf_single_frame(
            ts.positions[g1.indices], ts.positions[g2.indices], ts.dimensions,
            nbins, rng, exclusion_block, count=count, nthreads=nthreads)
        volume += v
    return count, volume, len(frames)


def oracle_interrdf_s(u, ags, nbins=75, rng=(0.0, 15.0), frames=None):
    traj = u.trajectory
    frames = range(traj.n_frames) if frames is None else frames
    counts = [np.zeros((len(a), len(b), nbins), dtype=np.float64)
              for a, b in ags]
    volume = 0.0
    for i in frames:
        ts = traj[i]
        pairs = [(ts.positions[a.indices], ts.positions[b.indices])
                 for a, b in ags]
        _, v = c_oracle.rdf_s_single_frame(pairs, ts.dimensions, nbins, rng,
                                           counts=counts)
        volume += v
    return counts, volume, len(frames)


def reference_rdf(count, volume, n_frames, nA, nB, edges, exclusion_block=None):
    """InterRDF._conclude arithmetic (pmda/rdf.py:143-160)"""
    N = nA * nB
    if exclusion_block:
        xA, xB = exclusion_block
        N -= xA * xB * (nA / xA)
    vol = np.power(edges[1:], 3) - np.power(edges[:-1], 3)
    vol *= 4 / 3.0 * np.pi
    with np.errstate(divide='ignore', invalid='ignore'):
        density = np.float64(N) / (np.float64(volume) / n_frames)
        return count / (density * vol * n_frames)
