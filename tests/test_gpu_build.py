"""HNSW.NODE.ADD on the GPU (core.rs:383-412, 489-599, 677-822) against the CPU oracle.

EXACT mode: same data, same injected levels, same order -> the device graph must equal the oracle's graph list by
list, including adjacency order, enterpoint and max_layer (SURVEY.md §8c tier 1).
FAST mode (batched; a labelled extension): structural invariants, recall, and tier-2 parity (the oracle searching
the exported device graph returns exactly what the device returns)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle  # noqa: E402
from gpu_fixtures import assert_search_parity, case  # noqa: E402
from redis_hnsw_b200 import data  # noqa: E402


def _lists(g):
    out = {}
    row = 0
    for i, lv in enumerate(g["levels"]):
        for l in range(int(lv) + 1):
            out[(i, l)] = g["nbrs"][int(g["row_offs"][row]):int(g["row_offs"][row + 1])]
            row += 1
    return out


def _assert_same_graph(gd, go, x=None):
    """Device graph == oracle graph, list by list and in list order.  With the vectors `x` given, ONE kind of difference is
    tolerated and counted: two entries of a list swapped whose sims to the list's node are bit-equal — the reference leaves
    the order of such a pair to BinaryHeap internals, so no fixture may pin it."""
    assert np.array_equal(gd["levels"], go["levels"])
    assert gd["entry"] == go["entry"] and gd["max_layer"] == go["max_layer"]
    if np.array_equal(gd["row_offs"], go["row_offs"]) and np.array_equal(gd["nbrs"], go["nbrs"]):
        return 0
    ld, lo = _lists(gd), _lists(go)
    bad = [k for k in lo if not np.array_equal(ld[k], lo[k])]
    tied = 0
    if x is not None:
        for (i, l) in list(bad):
            a, b = ld[(i, l)], lo[(i, l)]
            if a.size == b.size and np.array_equal(np.sort(a), np.sort(b)) and all(
                    oracle.euclidean(x[i], x[int(u)]) == oracle.euclidean(x[i], x[int(v)]) for u, v in zip(a, b) if u != v):
                bad.remove((i, l))
                tied += 1
    if bad:
        k = bad[0]
        raise AssertionError("%d of %d adjacency lists differ; first (node, level)=%r device=%r oracle=%r"
                             % (len(bad), len(lo), k, ld[k], lo[k]))
    assert tied <= 3, "%d lists differ by the order of tied entries: too many to be accidents of f32" % tied
    return tied


@pytest.mark.parametrize("name,n", [
    ("cfg1_10k_d32_m5", 4000),   # BASELINE configs[0] parameters
    ("d128_m16", 6000),          # configs[1] parameters
    ("d768_m32", 500),           # configs[2] shape
    ("d96_m8_generic", 1500),
    ("d20_m6_scalar", 1200),
    ("d64_m6_generic_v2", 1200),
    ("d256_m12_generic_v4", 900),
    ("d33_m4_scalar", 800),
])
def test_exact_build_equals_oracle_graph(name, n):
    import redis_hnsw_b200 as r

    c = case(name)
    x, levels = c["x"][:n], c["levels"][:n]
    orc = oracle.Oracle(c["dim"], c["m"], c["efc"])
    orc.add_batch(x, levels)
    dev = r.DeviceIndex(c["dim"], c["m"], c["efc"])
    first = dev.add_batch(x, levels, mode=r.BUILD_EXACT)
    assert first == 0
    _assert_same_graph(dev.export_graph(), orc.export_graph())
    p, op = dev.params(), orc.params()
    for key in ("node_count", "max_layer", "enterpoint"):
        assert p[key] == op[key], key
    assert_search_parity(dev, orc, c["q"][:300], 10, 64)
    st = dev.build_stats()
    assert st["inserts"] == n - 1 and st["reprunes"] > 0


@pytest.mark.parametrize("dim,m,efc,n,dataset", [
    (128, 16, 8, 2500, "lowrank"),     # staged kernels (insert_exact2_kernel)
    (32, 12, 5, 2000, "uniform"),      # staged, 32-d rows
    (20, 6, 3, 1200, "uniform"),       # scalar metric (insert_exact_kernel, the kernel the generic dims share)
])
def test_exact_build_with_ef_construction_below_m(dim, m, efc, n, dataset):
    """ef_construction < m: w comes back full with fewer than m entries, so select_neighbors' candidate extension
    (core.rs:698-721) is NOT redundant — nodes the search turned away only because w was full are linked until m are
    selected.  Graph identical to the oracle's, single adds included; FAST requests run the exact builder here.
    (Seeds are fixed to builds without a tie AT a selection cut: with ef_construction this small a 96-d uniform build
    (seed 5) hits sim(291, 1182) == sim(291, 1002) exactly at rank 20/21 of a re-selection, where the reference leaves
    the order to BinaryHeap internals — the parity definition excludes such ties)"""
    import redis_hnsw_b200 as r

    x, q = (data.lowrank(n + 40, dim, r=16, seed=5, n_queries=200) if dataset == "lowrank"
            else data.uniform(n + 40, dim, seed=5, n_queries=200))
    levels = data.draw_levels(n + 40, m, seed=6)
    orc = oracle.Oracle(dim, m, efc)
    orc.add_batch(x[:n], levels[:n])
    go = orc.export_graph()
    deg = np.diff(go["row_offs"])
    assert deg.max() > efc, "the case must exercise the extension (a node with more than ef_construction links)"
    for mode in (r.BUILD_EXACT, r.BUILD_FAST):
        dev = r.DeviceIndex(dim, m, efc)
        dev.add_batch(x[:n], levels[:n], mode=mode)
        _assert_same_graph(dev.export_graph(), go)
    longest = 0
    for i in range(n, n + 40):          # NODE.ADD one at a time on the last index
        assert dev.add(x[i], int(levels[i])) == orc.add(x[i], int(levels[i]))
        assert sorted(dev.touched()) == sorted(orc.touched())
        longest = max(longest, len(orc.node_neighbors(i, 0)))
    assert longest > efc                # a fresh node linked more neighbours than w held: the extension was exercised
    _assert_same_graph(dev.export_graph(), orc.export_graph())
    assert_search_parity(dev, orc, q, 10, 32)


def test_exact_build_in_pieces_and_single_adds_report_touched_nodes():
    """add_batch in several calls, then NODE.ADD one at a time: same graph, and the touched set equals what the
    reference reports through update_fn (core.rs:522,535-537,570-572,580-584)."""
    import redis_hnsw_b200 as r

    c = case("cfg1_10k_d32_m5")
    n0, n1 = 1500, 1700
    x, levels = c["x"][:n1], c["levels"][:n1]
    orc = oracle.Oracle(c["dim"], c["m"], c["efc"])
    dev = r.DeviceIndex(c["dim"], c["m"], c["efc"])
    for a, b in ((0, 1), (1, 2), (2, 700), (700, n0)):
        orc.add_batch(x[a:b], levels[a:b])
        assert dev.add_batch(x[a:b], levels[a:b], mode=r.BUILD_EXACT) == a
    _assert_same_graph(dev.export_graph(), orc.export_graph())
    for i in range(n0, n1):
        orc.add(x[i], int(levels[i]))
        assert dev.add(x[i], int(levels[i])) == i
        assert np.array_equal(dev.touched(), orc.touched()), i
    _assert_same_graph(dev.export_graph(), orc.export_graph())


def test_reference_core_kat_through_the_device():
    """src/hnsw/core_tests.rs:7-53 driven through the host mirror of Index: 100 nodes [i;4], m=5, efCon=16;
    search [10;4], k=5 -> sims 0,-4,-4,-16,-16 and the top hit is node10."""
    import redis_hnsw_b200 as r

    idx = r.Index("foo", 4, 5, 16)
    assert (idx.name, idx.mfunc_kind, idx.data_dim, idx.m, idx.m_max, idx.m_max_0, idx.ef_construction) == \
        ("foo", "Euclidean", 4, 5, 5, 10, 16)                      # core_tests.rs:12-19
    seen = []
    for i in range(100):
        idx.add_node("node%d" % i, np.full(4, float(i), np.float32), lambda name, node: seen.append(name))
    assert idx.node_count == 100                                   # :41
    with pytest.raises(r.HNSWError, match="already exists"):
        idx.add_node("node7", np.zeros(4, np.float32))             # core.rs:407-409
    with pytest.raises(r.HNSWError, match="data dimension: 3 does not match Index"):
        idx.add_node("x", np.zeros(3, np.float32))                 # core.rs:389-391
    res = idx.search_knn(np.full(4, 10.0, np.float32), 5)          # :45
    assert len(res) == 5
    assert [x.sim for x in res] == [0.0, -4.0, -4.0, -16.0, -16.0]  # :48-53
    assert res[0].name == "node10"
    assert np.array_equal(res[0].data, np.full(4, 10.0, np.float32))
    assert len(seen) > 100


def _check_invariants(g, m):
    lists = _lists(g)
    cap_exceeded = 0
    for (i, l), nb in lists.items():
        assert len(set(nb.tolist())) == len(nb), "duplicate neighbour in (%d,%d)" % (i, l)
        assert i not in nb, "self loop at (%d,%d)" % (i, l)
        cap = 2 * m if l == 0 else m
        cap_exceeded += len(nb) > cap
        for j in nb:
            assert g["levels"][j] >= l
            assert i in lists[(int(j), l)], "edge %d->%d on level %d is not mirrored" % (i, j, l)
    return cap_exceeded / max(1, len(lists))


@pytest.mark.parametrize("name,n,batch", [("d128_m16", 6000, 256), ("cfg1_10k_d32_m5", 10000, 0), ("d20_m6_scalar", 2000, 64)])
def test_fast_build_invariants_recall_and_tier2_parity(name, n, batch):
    import redis_hnsw_b200 as r

    c = case(name)
    x, levels, q = c["x"][:n], c["levels"][:n], c["q"][:500]
    dev = r.DeviceIndex(c["dim"], c["m"], c["efc"])
    if batch:
        dev.set_option("build_batch", batch)
    dev.add_batch(x, levels, mode=r.BUILD_FAST)
    st = dev.build_stats()
    # the batched builder's lossy valves (worklist full, re-selection skipped, hub row full) must stay shut here
    assert (st["fast_worklist_dropped"], st["fast_reprunes_skipped"], st["fast_edges_refused"]) == (0, 0, 0), st
    g = dev.export_graph()
    assert np.array_equal(g["levels"], levels.clip(min=0) * (np.arange(n) > 0))  # first node sits on level 0
    over = _check_invariants(g, c["m"])
    assert over < 0.08   # the reference leaves 1-3 % of the rows over their cap as well
    # tier-2 parity: the oracle searching the exported graph returns exactly the device's answers
    orc = oracle.Oracle(c["dim"], c["m"], c["efc"])
    orc.import_graph(x, g)
    ids, _, _, _ = assert_search_parity(dev, orc, q, 10, 64)
    # recall of the batched graph is on par with the sequential (oracle-built) graph at the same ef
    gt = data.brute_force_topk(x, q, 10)
    rec_fast = data.recall_at_k(ids, gt)
    seq = oracle.Oracle(c["dim"], c["m"], c["efc"])
    seq.add_batch(x, levels)
    rec_seq = data.recall_at_k(seq.search_batch(q, 10, ef=64)[0], gt)
    assert rec_fast >= rec_seq - 0.02, (rec_fast, rec_seq)


def test_fast_then_exact_and_mixed_calls():
    """Modes can be mixed on one index: a FAST bulk load followed by exact single NODE.ADDs."""
    import redis_hnsw_b200 as r

    c = case("d96_m8_generic")
    x, levels = c["x"], c["levels"]
    dev = r.DeviceIndex(c["dim"], c["m"], c["efc"])
    dev.add_batch(x[:2000], levels[:2000], mode=r.BUILD_FAST)
    orc = oracle.Oracle(c["dim"], c["m"], c["efc"])
    orc.import_graph(x[:2000], dev.export_graph())
    for i in range(2000, 2100):           # from the same starting graph the exact path must track the oracle
        orc.add(x[i], int(levels[i]))
        dev.add(x[i], int(levels[i]))
    _assert_same_graph(dev.export_graph(), orc.export_graph())


def test_drawn_levels_follow_the_reference_distribution():
    """level = floor(-ln(u) / ln(m)) (core.rs:601-605): P(level >= 1) = 1/m."""
    import redis_hnsw_b200 as r

    dev = r.DeviceIndex(32, 8, 32)
    dev.seed(7)
    x, _ = data.uniform(4000, 32, seed=5)
    dev.add_batch(x, None, mode=r.BUILD_FAST)
    lv = dev.export_graph()["levels"]
    assert lv[0] == 0
    assert abs((lv >= 1).mean() - 1 / 8) < 0.02 and abs((lv >= 2).mean() - 1 / 64) < 0.01
