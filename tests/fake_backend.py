"""CPU stand-in for the kernel backend, built on the ORACLE.  Installed through
engine._set_test_backend so that the host logic (block partition, frame
sharding over devices / ranks, result assembly, all-reduce) can be exercised
without a GPU.  Never used by the product path."""
import numpy as np
import torch

from oracle import c_oracle, numpy_oracle


def _dims_from(kind, box9):
    if kind == 0:
        return None
    if kind == 1:
        return ("ortho", np.asarray(box9[:3], dtype=np.float32))
    return ("tri", np.asarray(box9, dtype=np.float32))


class OracleBackend(object):
    name = "oracle-cpu"

    def __init__(self, n_devices=1):
        self.n_devices = n_devices
        self.calls = 0

    def devices(self, n_jobs):
        devs = [torch.device("cpu"), torch.device("cpu", 0)]
        return devs[:max(1, min(int(n_jobs), self.n_devices))]

    @staticmethod
    def rdf_workspace_bytes(F, n1, n2, nbins, same_group):
        return 256

    @staticmethod
    def host_hash_rows(arr, seed=0, nthreads=1):
        # host code of the product library (no CUDA involved)
        from pmda_b200 import _cabi
        return _cabi.host_hash_rows(arr, seed, nthreads)

    def _frame(self, g1, g2, kind, box9, nbins, rng, excl, count):
        lib = c_oracle.lib()
        edges = c_oracle.edges_of(nbins, rng)
        xa, xb = (0, 0) if excl is None else excl
        b = np.zeros(9, dtype=np.float32)
        b[:] = box9
        lib.oracle_rdf_frame(np.ascontiguousarray(g1), len(g1),
                             np.ascontiguousarray(g2), len(g2), int(kind), b,
                             float(rng[1]), float(rng[0]), float(rng[1]),
                             int(nbins), edges, int(xa), int(xb), count, 2)

    def rdf_hist(self, xyz, F, natoms, idx1, n1, idx2, n2, same_group, boxkind,
                 box, r0, r1, nbins, edges, excl, hist, workspace, stream):
        self.calls += 1
        pos = xyz.numpy()
        i1, i2 = idx1.numpy(), idx2.numpy()
        cnt = np.zeros(nbins, dtype=np.int64)
        for f in range(F):
            self._frame(pos[f][i1], pos[f][i2], boxkind, box[f].numpy(), nbins,
                        (r0, r1), excl, cnt)
        hist += torch.from_numpy(cnt)

    def rdf_sites(self, xyz, F, natoms, idx_a, idx_b, off_a, off_b, cell_off,
                  n_sites, total_cells, boxkind, box, r0, r1, nbins, edges,
                  count, stream):
        self.calls += 1
        lib = c_oracle.lib()
        pos = xyz.numpy()
        ia, ib = idx_a.numpy(), idx_b.numpy()
        oa, ob, co = off_a.numpy(), off_b.numpy(), cell_off.numpy()
        e = c_oracle.edges_of(nbins, (r0, r1))
        dense = np.zeros((total_cells, nbins), dtype=np.float64)
        for f in range(F):
            b = np.zeros(9, dtype=np.float32)
            b[:] = box[f].numpy()
            for s in range(n_sites):
                pa = np.ascontiguousarray(pos[f][ia[oa[s]:oa[s + 1]]])
                pb = np.ascontiguousarray(pos[f][ib[ob[s]:ob[s + 1]]])
                view = dense[co[s]:co[s + 1]].reshape(-1)
                lib.oracle_rdf_s_frame(pa, len(pa), pb, len(pb), int(boxkind),
                                       b, float(r1), float(r0), float(r1),
                                       int(nbins), e, view)
        count += torch.from_numpy(dense.reshape(-1).astype(np.int32))

    @staticmethod
    def host_gather_rows(src, idx, dst, nthreads=1):
        from pmda_b200 import _cabi          # host code of the product library
        return _cabi.host_gather_rows(src, idx, dst, nthreads)

    @staticmethod
    def host_sites_finish(count, nbins, den=None, want_total=False,
                          nthreads=1):
        from pmda_b200 import _cabi          # host code of the product library
        return _cabi.host_sites_finish(count, nbins, den, want_total, nthreads)

    def contacts(self, xyz, F, natoms, idx_a, idx_b, r0, n_pairs, method,
                 radius, beta, lam, out, out_stride, stream):
        self.calls += 1
        pos = xyz.numpy()
        a, b = idx_a.numpy(), idx_b.numpy()
        r0 = None if r0 is None else r0.numpy()
        for f in range(F):
            dx = (pos[f][b] - pos[f][a]).astype(np.float64)   # float32 diff
            d = np.sqrt((dx[:, 0] * dx[:, 0] + dx[:, 1] * dx[:, 1]) +
                        dx[:, 2] * dx[:, 2])
            if method == 0:
                out[f, 0] += float((d <= r0).sum())
            elif method == 1:
                out[f, 0] += float((d <= radius).sum())
            elif method == 2:
                out[f, 0] += float((1 / (1 + np.exp(beta * (d - lam * r0)))).sum())
            else:
                out[f, :n_pairs] = torch.from_numpy(d)

    @staticmethod
    def hbonds_workspace_bytes(F, n_dh, n_acc):
        return 256

    def hbonds(self, xyz, F, natoms, donors, hydrogens, n_dh, acceptors, n_acc,
               boxkind, box, max_cutoff, min_cutoff, angle_cutoff, capacity,
               out_row, out_acc, out_dist, out_angle, total, workspace,
               stream):
        self.calls += 1
        pos = xyz.numpy()
        d = donors.numpy()
        h = None if hydrogens is None else hydrogens.numpy()
        a = acceptors.numpy()
        n = 0
        for f in range(F):
            dims = None
            if boxkind == 1:
                dims = ("ortho", box[f].numpy()[:3].astype(np.float32))
            elif boxkind == 2:
                dims = ("tri_vecs",
                        box[f].numpy().astype(np.float32).reshape(3, 3))
            k, j, dist, ang = numpy_oracle.hbond_rows(
                pos[f], d, h, a, dims, max_cutoff, min_cutoff, angle_cutoff)
            for i in range(len(k)):
                if n < capacity:
                    out_row[n] = f * n_dh + int(k[i])
                    out_acc[n] = int(j[i])
                    out_dist[n] = float(dist[i])
                    out_angle[n] = float(ang[i])
                n += 1
        total[0] = n
