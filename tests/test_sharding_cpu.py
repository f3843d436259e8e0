"""N > 1 host logic on CPU (world_size 2, gloo): query slicing and result gathering of redis_hnsw_b200/sharding.py.
The per-rank engine here is the CPU oracle standing in for a replicated device index (no GPU in this test)."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nq, out_dir):
    import torch.distributed as dist

    import oracle
    from redis_hnsw_b200 import data, sharding

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x, q = data.uniform(1500, 32, seed=3, n_queries=nq)
    levels = data.draw_levels(1500, 5, seed=9)
    orc = oracle.Oracle(32, 5, 40)   # every rank holds the same replica
    orc.add_batch(x, levels)

    def engine(qs):
        ids, sims, counts, _, _ = orc.search_batch(qs, 10, ef=40, stats=False)
        return ids, sims, counts

    ids, sims, counts = sharding.sharded_search(engine, q, 10, rank, world)
    full = engine(q)
    ok = all(np.array_equal(a, b) for a, b in zip((ids, sims, counts), full))
    lo, hi = sharding.query_slice(nq, rank, world)
    np.save(os.path.join(out_dir, "r%d.npy" % rank), np.array([int(ok), lo, hi]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("nq", [64, 37])
def test_sharded_search_gloo_world2(tmp_path, nq):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, nq, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (np.load(tmp_path / ("r%d.npy" % r)) for r in range(2))
    assert r0[0] == 1 and r1[0] == 1
    assert r0[1] == 0 and r0[2] == r1[1] and r1[2] == nq   # contiguous, complete, disjoint slices


def test_sharded_search_gloo_world3_with_an_empty_rank(tmp_path):
    """Fewer queries than ranks: a rank with an empty slice still takes part in the gather."""
    port = _free_port()
    mp.spawn(_worker, args=(3, port, 2, str(tmp_path)), nprocs=3, join=True)
    res = [np.load(tmp_path / ("r%d.npy" % r)) for r in range(3)]
    assert all(r[0] == 1 for r in res)
    assert [int(r[2] - r[1]) for r in res] == [1, 1, 0]


def test_query_slice_partitions():
    from redis_hnsw_b200 import sharding

    for nq in (0, 1, 7, 8, 10000, 10001):
        for world in (1, 2, 3, 4, 8):
            cuts = [sharding.query_slice(nq, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == nq
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1
