"""N-rank / N-block invariance against a committed oracle golden
(tests/golden/nranks_counts.npz, made by make_nranks_golden.py): the
reference pins that the result does not depend on how the frames are cut
(pmda/test/test_rdf.py:76-93, test_rdf_s.py:110-138); here "how the frames are
cut" also means over how many GPUs / ranks.  bench.py prints the same check as
`parity_nranks` at every N.

CPU: the oracle still reproduces frame 0 of the fixture and the seeded inputs
are unchanged.  GPU: one process with n_blocks in 1..4, and TWO ranks (gloo
rendezvous, both driving the one GPU of the test lease with the real kernels)
must give the fixture's integers.
"""
import os
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_nranks_golden as mk        # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "nranks_counts.npz"))


@pytest.mark.parametrize("name", sorted(mk.CASES))
def test_inputs_and_oracle_reproduce_the_fixture(name):
    assert mk.digest(name) == GOLD[name + "__sha256"].tobytes().decode()
    np.testing.assert_array_equal(mk.oracle_rdf(name, [0]),
                                  GOLD[name + "__frame0"])


def test_oracle_reproduces_the_site_fixture():
    np.testing.assert_array_equal(mk.oracle_rdf_s([0]).astype(np.int32),
                                  GOLD["sites__frame0"])
    assert GOLD["sites__total"].sum() > GOLD["sites__frame0"].sum() > 0


def _run_all(n_blocks_list, out):
    """every analysis of the fixture through run(); `out` collects results"""
    from pmda_b200.rdf import InterRDF, InterRDF_s
    from pmda_b200.universe import Universe
    for name in sorted(mk.CASES):
        pos, dims = mk.make_positions(name)
        _, _, nbins, rng = mk.CASES[name]
        u = Universe(pos, dims)
        for nb in n_blocks_list:
            r = InterRDF(u.atoms, u.atoms, nbins=nbins, range=rng).run(
                n_blocks=nb)
            out["%s_nb%d" % (name, nb)] = r.count
        r = InterRDF(u.atoms, u.atoms, nbins=nbins, range=rng).run(
            start=0, stop=1)
        out["%s_frame0" % name] = r.count
        if name == "ortho":
            ags = [[u.atoms[a], u.atoms[b]] for a, b in mk.site_indices()]
            for nb in n_blocks_list[:2]:
                s = InterRDF_s(u, ags, nbins=nbins, range=rng).run(n_blocks=nb)
                out["sites_nb%d" % nb] = np.stack([c[0] for c in s.count])


def _check(out, n_blocks_list):
    for name in sorted(mk.CASES):
        for nb in n_blocks_list:
            np.testing.assert_array_equal(out["%s_nb%d" % (name, nb)],
                                          GOLD[name + "__total"])
        np.testing.assert_array_equal(out["%s_frame0" % name],
                                      GOLD[name + "__frame0"])
    for nb in n_blocks_list[:2]:
        np.testing.assert_array_equal(out["sites_nb%d" % nb],
                                      GOLD["sites__total"])


@pytest.mark.gpu
def test_one_gpu_any_number_of_blocks():
    from pmda_b200 import engine
    engine._set_test_backend(None)
    out = {}
    _run_all([1, 2, 3, 4], out)
    _check(out, [1, 2, 3, 4])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _rank_worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    # one GPU per rank when the lease has them; otherwise the ranks share
    # cuda:0 (NCCL refuses two ranks on one device: gloo carries the
    # all-reduce, the kernels are the real ones either way)
    ngpu = torch.cuda.device_count()
    torch.cuda.set_device(rank % ngpu)
    backend = "nccl" if ngpu >= world else "gloo"
    dist.init_process_group(backend, rank=rank, world_size=world)
    from pmda_b200 import _cabi, engine
    engine._set_test_backend(None)
    l0 = _cabi.launch_count()
    out = {}
    _run_all([world, 2 * world + 1], out)
    out["launches"] = np.array(_cabi.launch_count() - l0)
    out["backend"] = np.frombuffer(backend.encode(), np.uint8)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), **out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 3])
def test_ranks_give_the_fixture_integers_with_the_real_kernels(tmp_path, world):
    import torch.multiprocessing as mp
    mp.spawn(_rank_worker, args=(world, _free_port(), str(tmp_path)),
             nprocs=world, join=True)
    ranks = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
             for r in range(world)]
    for z in ranks:
        _check(z, [world, 2 * world + 1])
        assert int(z["launches"]) > 0          # every rank ran kernels
