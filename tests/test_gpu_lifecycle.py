"""Several indexes alive at once (the reference keeps a global map of indexes, lib.rs:32-35), interleaved commands,
destroy/recreate cycles without leaking device memory."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle  # noqa: E402
from redis_hnsw_b200 import data  # noqa: E402


def test_interleaved_indexes_are_independent():
    import redis_hnsw_b200 as r

    specs = [(32, 5, 40), (128, 16, 64), (20, 6, 32), (96, 8, 32)]
    n = 400
    pairs = []
    for dim, m, efc in specs:
        x, q = data.uniform(n, dim, seed=dim, n_queries=50)
        lv = data.draw_levels(n, m, seed=dim + 1)
        pairs.append((r.DeviceIndex(dim, m, efc), oracle.Oracle(dim, m, efc), x, q, lv))
    for i in range(n):                                   # round-robin NODE.ADD across the four indexes
        for dev, orc, x, q, lv in pairs:
            assert dev.add(x[i], int(lv[i])) == orc.add(x[i], int(lv[i]))
    for dev, orc, x, q, lv in pairs:
        ids, sims, cnt = dev.search_batch(q, 5, ef=32)
        oids, osims, ocnt, ost, _ = orc.search_batch(q, 5, ef=32)
        ok = ost[:, 3] == 0
        assert np.array_equal(ids[ok], oids[ok]) and np.array_equal(sims[ok].view(np.uint32), osims[ok].view(np.uint32))
    for dev, orc, x, q, lv in pairs:                     # and deletes
        for v in (1, 17, 200):
            dev.delete(v)
            orc.delete(v)
        gd, go = dev.export_graph(), orc.export_graph()
        assert np.array_equal(gd["nbrs"], go["nbrs"]) and np.array_equal(gd["row_offs"], go["row_offs"])
        dev.close()


def test_create_destroy_cycles_do_not_leak():
    import torch

    import redis_hnsw_b200 as r

    x, q = data.lowrank(20000, 128, seed=3, n_queries=100)
    lv = data.draw_levels(20000, 16, seed=4)

    def cycle():
        dev = r.DeviceIndex(128, 16, 100)
        dev.add_batch(x, lv, mode=r.BUILD_FAST)
        dev.search_batch(q, 10, ef=64)
        dev.search_batch(q, 10, ef=64, stats=True)
        dev.add(q[0], -1)
        dev.delete(5)
        dev.close()

    cycle()
    torch.cuda.synchronize()
    free0, _ = torch.cuda.mem_get_info()
    for _ in range(5):
        cycle()
    torch.cuda.synchronize()
    free1, _ = torch.cuda.mem_get_info()
    assert free0 - free1 < 64 << 20, "device memory leaked across create/destroy cycles: %d MB" % ((free0 - free1) >> 20)
