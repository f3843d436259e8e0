"""BASELINE.json configs[1] at FULL size (1M x 128-d, M=16, efCon=200, ef=64, k=10) on the GPU, checked through
size-independent properties plus a bounded oracle comparison (the oracle cannot build 1M nodes in test time, so the graph
is the device's FAST build, exported; SURVEY.md §8c tier 2)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N, DIM, M, EFC, EF, K = 1_000_000, 128, 16, 200, 64, 10


@pytest.fixture(scope="module")
def big():
    import redis_hnsw_b200 as r
    from redis_hnsw_b200 import data

    x, q = data.lowrank(N, DIM, r=16, seed=123, n_queries=20_000)
    levels = data.draw_levels(N, M, seed=42)
    dev = r.DeviceIndex(DIM, M, EFC)
    dev.reserve(N)
    dev.add_batch(x, levels, mode=r.BUILD_FAST)
    return dict(dev=dev, x=x, q=q, levels=levels)


def test_params_and_structure(big):
    dev, levels = big["dev"], big["levels"]
    p = dev.params()
    assert p["node_count"] == N and p["n_ids"] == N and p["m_max_0"] == 2 * M
    st = dev.build_stats()                                         # the FAST builder dropped nothing on the headline workload
    assert (st["fast_worklist_dropped"], st["fast_reprunes_skipped"], st["fast_edges_refused"]) == (0, 0, 0), st
    top = int(levels[1:].max())                                    # the first node is forced to level 0 (core.rs:393-405)
    assert p["max_layer"] == top and dev.node_level(p["enterpoint"]) == top   # core.rs:587-593
    rng = np.random.default_rng(0)
    over = 0
    for i in rng.integers(0, N, 300):                              # symmetric, duplicate-free, no self loops (fact #5)
        i = int(i)
        for lv in range(dev.node_level(i) + 1):
            nb = dev.node_neighbors(i, lv)
            assert len(set(nb.tolist())) == len(nb) and i not in nb
            over += len(nb) > (2 * M if lv == 0 else M)
            for j in nb[:4]:
                assert i in dev.node_neighbors(int(j), lv)
    assert over < 60                                               # over-full rows exist (core.rs:793-795) but are rare


def test_search_properties_at_full_size(big):
    import redis_hnsw_b200 as r
    from redis_hnsw_b200 import data

    dev, x, q = big["dev"], big["x"], big["q"]
    ids, sims, counts = dev.search_batch(q, K, ef=EF)              # staged kernel
    assert (counts == K).all()
    assert (np.diff(sims, axis=1) <= 0).all()                      # nearest first (core.rs:878-891)
    assert all(len(set(row.tolist())) == K for row in ids[:2000])
    # every returned sim is the reference metric of (query, stored vector), bit for bit
    flat = ids[:3000].reshape(-1).astype(np.int64)
    again = r.l2_batch(np.repeat(q[:3000], K, axis=0), x[flat])
    assert np.array_equal(again.view(np.uint32), sims[:3000].reshape(-1).view(np.uint32))
    # idempotent, and both kernels agree (exact visited set vs lossy tag table)
    ids2, sims2, _ = dev.search_batch(q, K, ef=EF)
    assert np.array_equal(ids, ids2) and np.array_equal(sims.view(np.uint32), sims2.view(np.uint32))
    ids3, sims3, _, st = dev.search_batch(q[:5000], K, ef=EF, stats=True)
    assert np.array_equal(ids3, ids[:5000]) and np.array_equal(sims3.view(np.uint32), sims[:5000].view(np.uint32))
    # recall@10 >= 0.95 at the headline operating point
    gt = data.brute_force_topk(x, q[:2000], K, device="cuda")
    assert data.recall_at_k(ids[:2000], gt) >= 0.95
    # a larger ef can only help, k > ef truncates (core.rs:879)
    ids4, _, c4 = dev.search_batch(q[:2000], K, ef=128)
    assert data.recall_at_k(ids4, gt) >= data.recall_at_k(ids[:2000], gt)
    _, _, c5 = dev.search_batch(q[:100], K, ef=4)
    assert (c5 == 4).all()


def test_oracle_on_the_exported_graph_matches(big):
    import oracle

    dev, x, q = big["dev"], big["x"], big["q"]
    orc = oracle.Oracle(DIM, M, EFC)
    orc.import_graph(x, dev.export_graph())
    nq = 3000
    ids, sims, counts, st = dev.search_batch(q[:nq], K, ef=EF, stats=True)
    oids, osims, ocounts, ost, _ = orc.search_batch(q[:nq], K, ef=EF, threads=8)
    ok = ost[:, 3] == 0
    assert ok.mean() > 0.99
    assert np.array_equal(ids[ok], oids[ok]) and np.array_equal(sims[ok].view(np.uint32), osims[ok].view(np.uint32))
    assert np.array_equal(st[ok, :3].astype(np.uint64), ost[ok, :3])
    # the reference's own ef (= ef_construction, core.rs:485)
    ids2, sims2, _ = dev.search_batch(q[:500], K)
    oids2, osims2, _, ost2, _ = orc.search_batch(q[:500], K, threads=8)
    ok2 = ost2[:, 3] == 0
    assert np.array_equal(ids2[ok2], oids2[ok2]) and np.array_equal(sims2[ok2].view(np.uint32), osims2[ok2].view(np.uint32))


def test_exact_and_spec_inserts_at_full_size(big):
    """NODE.ADD on the 1M-node graph (core.rs:489-599): one command at a time through the one-warp EXACT kernel, then a
    stream through the SPEC builder — both must leave the lists the oracle's insert leaves, for the new nodes and for the
    old nodes they were linked to.  (Runs last: it grows the index.)"""
    import oracle
    import redis_hnsw_b200 as r
    from redis_hnsw_b200 import data

    dev, x, q = big["dev"], big["x"], big["q"]
    orc = oracle.Oracle(DIM, M, EFC)
    orc.import_graph(x, dev.export_graph())
    n_single, n_stream = 60, 400
    new = q[12_000:12_000 + n_single + n_stream]                   # same distribution as the data (same generator stream)
    lv = data.draw_levels(n_single + n_stream, M, seed=4242)
    for i in range(n_single):
        assert dev.add(new[i], int(lv[i])) == orc.add(new[i], int(lv[i])) == N + i
    orc.add_batch(new[n_single:], lv[n_single:])
    assert dev.add_batch(new[n_single:], lv[n_single:], mode=r.BUILD_SPEC) == N + n_single
    st = dev.build_stats()
    assert st["spec_rounds"] > 0 and st["spec_rounds"] < n_stream  # windows did commit more than one insert per round
    p, op = dev.params(), orc.params()
    for key in ("node_count", "max_layer", "enterpoint"):
        assert p[key] == op[key], key
    touched = set()
    for j in range(N, N + n_single + n_stream):
        assert dev.node_level(j) == int(lv[j - N])
        for level in range(int(lv[j - N]) + 1):
            nb = dev.node_neighbors(j, level)
            assert np.array_equal(nb, orc.node_neighbors(j, level)), (j, level)
            if level == 0:
                touched.update(int(v) for v in nb)
    for v in sorted(touched)[:3000]:                               # old rows that got a new neighbour (and maybe a re-selection)
        assert np.array_equal(dev.node_neighbors(v, 0), orc.node_neighbors(v, 0)), v
