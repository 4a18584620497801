"""world_size-2 `gloo` run of the frame sharding + all-reduce path on CPU
(kernels replaced by the oracle-backed stand-in)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _make_universe():
    from pmda_b200.universe import Universe
    rng = np.random.default_rng(77)
    L = 22.0
    pos = (rng.random((9, 200, 3)) * L).astype(np.float32)
    return Universe(pos, [L, L, L, 90, 90, 90])


def _make_water():
    from pmda_b200 import synthetic
    from pmda_b200.universe import Universe
    rng = np.random.default_rng(78)
    dims = [15.0, 15.0, 15.0, 90, 90, 90]
    pos, names = synthetic.water_box(5, 45, dims, rng)
    u = Universe(pos, dims, names=names)
    o = 3 * np.arange(45)
    return u, np.repeat(o, 2), (o[:, None] + np.array([1, 2])).reshape(-1), o


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pmda_b200 import engine
    from pmda_b200.rdf import InterRDF, InterRDF_s
    from fake_backend import OracleBackend
    be = OracleBackend()
    engine._set_test_backend(be)
    u = _make_universe()
    g1, g2 = u.atoms[:80], u.atoms[80:]
    res = {}
    for nb in (1, 2, 3):
        r = InterRDF(g1, g2).run(n_blocks=nb)
        res["count_nb%d" % nb] = r.count
        res["rdf_nb%d" % nb] = r.rdf
        res["vol_nb%d" % nb] = np.float64(r.volume)
        res["blocks_nb%d" % nb] = np.stack(list(r._results[:, 0]))
        res["io_len_nb%d" % nb] = len(r.timing.io)
    ags = [[u.atoms[[3]], u.atoms[10:14]], [u.atoms[[5, 6]], u.atoms[20:23]]]
    s = InterRDF_s(u, ags).run(n_blocks=2)
    res["site0"], res["site1"] = s.count[0], s.count[1]
    from pmda_b200.hbond_analysis import HydrogenBondAnalysis
    uw, don, hyd, acc = _make_water()
    for nb in (1, 3):
        hb = HydrogenBondAnalysis(uw, acceptors_sel=acc, pairs=(don, hyd))
        hb.run(n_blocks=nb)
        res["hbonds_nb%d" % nb] = hb.hbonds
    res["calls"] = be.calls
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), **res)
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_gloo(tmp_path):
    sys.path.insert(0, HERE)
    from helpers import oracle_interrdf, oracle_interrdf_s
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world,
             join=True)
    u = _make_universe()
    g1, g2 = u.atoms[:80], u.atoms[80:]
    want, vol, nf = oracle_interrdf(u, g1, g2)
    ranks = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
             for r in range(world)]
    for z in ranks:
        for nb in (1, 2, 3):
            assert np.array_equal(z["count_nb%d" % nb], want)
            assert float(z["vol_nb%d" % nb]) == pytest.approx(vol, rel=1e-14)
            assert z["blocks_nb%d" % nb].shape == (nb, 75)
            assert np.array_equal(z["blocks_nb%d" % nb].sum(0), want)
            assert int(z["io_len_nb%d" % nb]) == nf
    # both ranks hold identical results, bit for bit
    for key in ranks[0].files:
        if key != "calls":
            assert np.array_equal(ranks[0][key], ranks[1][key]), key
    # per-block results keep the reference meaning (block b = its frames)
    from pmda_b200.util import make_balanced_slices
    blocks = make_balanced_slices(9, 3, 0, 9, 1)
    for b, sl in enumerate(blocks):
        c, _, _ = oracle_interrdf(u, g1, g2, frames=range(sl.start, sl.stop))
        assert np.array_equal(ranks[0]["blocks_nb3"][b], c)
    ags = [[u.atoms[[3]], u.atoms[10:14]], [u.atoms[[5, 6]], u.atoms[20:23]]]
    counts, _, _ = oracle_interrdf_s(u, ags)
    assert np.array_equal(ranks[1]["site0"], counts[0])
    assert np.array_equal(ranks[1]["site1"], counts[1])
    # ragged hydrogen-bond tables: gathered from both ranks, in frame order
    from oracle import numpy_oracle
    uw, don, hyd, acc = _make_water()
    rows = np.vstack([numpy_oracle.hbonds_single_frame(
        i, uw.trajectory[i].positions, don, hyd, acc,
        uw.trajectory[i].dimensions) for i in range(5)])
    assert len(rows) > 3
    for z in ranks:
        assert np.array_equal(z["hbonds_nb1"], rows)
        assert np.array_equal(z["hbonds_nb3"], rows)
    # every rank did part of the work
    assert all(int(z["calls"]) > 0 for z in ranks)


def test_frame_shards_cover_everything_once():
    from pmda_b200.engine import _frame_shards
    blocks = [range(0, 4), range(4, 7), range(7, 10)]
    for parts in (1, 2, 3, 4, 8, 16):
        shards = _frame_shards(blocks, parts)
        seen = []
        for p in shards:
            for b, pos0, frames in p:
                assert list(frames) == list(range(pos0, pos0 + len(frames)))
                assert set(frames) <= set(blocks[b])
                seen += list(frames)
        assert sorted(seen) == list(range(10))
        sizes = [sum(len(f) for _, _, f in p) for p in shards]
        used = [s for s in sizes if s]
        assert max(used) - min(used) <= 1
    # strided blocks keep their stride
    blocks = [range(0, 12, 3), range(12, 20, 3)]
    shards = _frame_shards(blocks, 2)
    got = sorted(f for p in shards for _, _, fr in p for f in fr)
    assert got == [0, 3, 6, 9, 12, 15, 18]
