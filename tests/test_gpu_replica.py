"""Index replication: the source index exposes its device buffers, a replica created with the same
parameters adopts a byte-for-byte copy of them and answers identically.  bench.py does the copy with one ncclBroadcast
per buffer across GPUs; here the copy is a device-to-device memcpy on ONE GPU, which exercises the same
layout / prepare / adopt protocol."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from redis_hnsw_b200 import data  # noqa: E402
from redis_hnsw_b200.sharding import _DevPtr, query_slice  # noqa: E402


def test_replica_adopts_buffers_and_answers_identically():
    import torch

    import redis_hnsw_b200 as r

    n, dim, m, efc = 30000, 128, 16, 100
    x, q = data.lowrank(n, dim, seed=9, n_queries=2000)
    lv = data.draw_levels(n, m, seed=10)
    src = r.DeviceIndex(dim, m, efc)
    src.add_batch(x, lv, mode=r.BUILD_FAST)
    for v in (3, 500, 29999):
        src.delete(v)                                   # replicas carry deletions too
    rep = r.DeviceIndex(dim, m, efc)
    rep.prepare_replica(src.replica_layout())
    sb, rb = src.device_buffers(), rep.device_buffers()
    assert len(sb) == len(rb) == 9 and [b for _, b in sb] == [b for _, b in rb]
    for (sp, nbytes), (rp, _) in zip(sb, rb):
        if nbytes:
            torch.as_tensor(_DevPtr(rp, nbytes), device="cuda").copy_(torch.as_tensor(_DevPtr(sp, nbytes), device="cuda"))
    torch.cuda.synchronize()
    rep.adopt_replica()
    ps, pr = src.params(), rep.params()
    for key in ("node_count", "n_ids", "max_layer", "enterpoint", "m_max_0"):
        assert ps[key] == pr[key], key
    ids, sims, cnt = src.search_batch(q, 10, ef=64)
    # each "rank" answers its slice of the batch on its own replica; together they reproduce the single-index answer
    for rank, idx in enumerate((src, rep)):
        lo, hi = query_slice(len(q), rank, 2)
        i2, s2, c2 = idx.search_batch(q[lo:hi], 10, ef=64)
        assert np.array_equal(i2, ids[lo:hi]) and np.array_equal(s2.view(np.uint32), sims[lo:hi].view(np.uint32))
        assert np.array_equal(c2, cnt[lo:hi])
    gs, gr = src.export_graph(), rep.export_graph()
    assert np.array_equal(gs["nbrs"], gr["nbrs"]) and np.array_equal(gs["levels"], gr["levels"])
    # the replica is a full index: it keeps accepting mutations
    assert rep.add(q[0], 0) == src.add(q[0], 0)
