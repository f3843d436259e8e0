"""HNSW.NODE.DEL on the GPU (core.rs:414-475 delete_node, :824-863 delete_node_from_neighbors) against the CPU oracle.

Same graph, same victims in the same order -> after every delete the device graph must equal the oracle's graph list
by list (order included), the touched set must equal what the reference reports through update_fn, and entry point /
max_layer / node_count must follow (the oracle and the device both replace a deleted enterpoint by the smallest id of
the highest non-empty layer; the reference takes an arbitrary HashSet element there, core.rs:453)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle  # noqa: E402
from gpu_fixtures import CASES, assert_search_parity  # noqa: E402
from redis_hnsw_b200 import data  # noqa: E402
from test_gpu_build import _assert_same_graph, _check_invariants  # noqa: E402


def _fresh(name, n):
    """A private oracle + device pair holding the same exact-built graph (the cached fixtures are shared: never mutate them)."""
    import redis_hnsw_b200 as r

    _, dim, m, efc, ds, nq = CASES[name]
    if ds == "uniform":
        x, q = data.uniform(n, dim, seed=7, n_queries=200)
    else:
        x, q = data.lowrank(n, dim, r=int(ds[7:]), seed=7, n_queries=200)
    levels = data.draw_levels(n, m, seed=11)
    orc = oracle.Oracle(dim, m, efc)
    orc.add_batch(x, levels)
    dev = r.DeviceIndex(dim, m, efc)
    dev.load_graph(x, orc.export_graph())
    return orc, dev, x, q, levels, m


@pytest.mark.parametrize("name,n,n_del", [
    ("cfg1_10k_d32_m5", 1500, 120),
    ("d128_m16", 1200, 60),
    ("d768_m32", 300, 12),
    ("d96_m8_generic", 800, 40),
    ("d20_m6_scalar", 800, 40),
    ("d64_m6_generic_v2", 700, 30),
    ("d256_m12_generic_v4", 500, 20),
])
def test_delete_equals_oracle(name, n, n_del):
    orc, dev, x, q, levels, m = _fresh(name, n)
    rng = np.random.default_rng(5)
    victims = rng.permutation(n)[:n_del].tolist()
    ep = orc.params()["enterpoint"]
    victims[3] = ep if ep not in victims else victims[3]          # the enterpoint goes too (core.rs:449-472)
    victims = list(dict.fromkeys(victims))
    for t, v in enumerate(victims):
        orc.delete(v)
        dev.delete(v)
        assert np.array_equal(dev.touched(), orc.touched()), "touched set differs after deleting %d" % v
        p, op = dev.params(), orc.params()
        for key in ("node_count", "max_layer", "enterpoint"):
            assert p[key] == op[key], (key, v)
        assert dev.node_level(v) == -1
        if t % 10 == 0 or t == len(victims) - 1:
            _assert_same_graph(dev.export_graph(), orc.export_graph())
    g = dev.export_graph()
    live = g["levels"] >= 0
    assert int(live.sum()) == n - len(victims)
    assert not np.isin(g["nbrs"], np.asarray(victims, np.uint32)).any()      # core_tests.rs:69-79: nothing dangles
    assert_search_parity(dev, orc, q, 10, 64)
    # NODE.ADD after NODE.DEL: the stream continues on the same graph
    x2, _ = data.uniform(40, x.shape[1], seed=99, n_queries=1) if "uniform" == CASES[name][4] else \
        data.lowrank(40, x.shape[1], r=int(CASES[name][4][7:]), seed=99, n_queries=1)
    lv2 = data.draw_levels(40, m, seed=3)
    for i in range(40):
        a = orc.add(x2[i], int(lv2[i]))
        assert dev.add(x2[i], int(lv2[i])) == a
    _assert_same_graph(dev.export_graph(), orc.export_graph())
    with pytest.raises(Exception, match="does not exist"):
        dev.delete(victims[0])                                    # core.rs:421
    with pytest.raises(Exception, match="does not exist"):
        dev.delete(10 ** 6)


def test_reference_core_kat_delete_everything():
    """src/hnsw/core_tests.rs:56-80 through the host mirror of Index: delete node0..node99 one by one; after each,
    node_count drops, the name leaves `nodes`, and no adjacency list on any level still holds the node."""
    import redis_hnsw_b200 as r

    idx = r.Index("foo", 4, 5, 16)
    for i in range(100):
        idx.add_node("node%d" % i, np.full(4, float(i), np.float32))
    dev = idx.device_index
    for i in range(100):
        seen = []
        idx.delete_node("node%d" % i, lambda name, node: seen.append(name))
        assert idx.node_count == 100 - i - 1                      # :60
        assert "node%d" % i not in idx.nodes                      # :61
        assert "node%d" % i not in seen                           # the victim is never written back (core.rs:810-813)
        g = dev.export_graph()
        assert g["levels"][i] == -1                               # :63-67 gone from its layer set
        assert not (g["nbrs"] == i).any()                         # :69-79
        if i < 99:
            _check_invariants_live(g)
            res = idx.search_knn(np.full(4, float(i + 1), np.float32), 1)
            assert res[0].sim == 0.0 and res[0].name == "node%d" % (i + 1)
    assert idx.enterpoint is None and idx.node_count == 0
    assert idx.search_knn(np.zeros(4, np.float32), 3) == []        # core.rs:481-483
    with pytest.raises(r.HNSWError, match="does not exist"):
        idx.delete_node("node5")
    # the index is usable again (core.rs:393-405: first-node path)
    idx.add_node("again", np.ones(4, np.float32))
    assert idx.node_count == 1 and idx.enterpoint == "again"
    assert idx.search_knn(np.ones(4, np.float32), 1)[0].name == "again"


def _check_invariants_live(g):
    lv = g["levels"]
    sub = dict(g)
    assert g["entry"] >= 0 and lv[g["entry"]] == g["max_layer"]
    _check_invariants(sub, 5)
