"""search_layer / search_knn on the GPU against the CPU oracle on the same graph (SURVEY.md §8c tier 1):
ids bit-exact, sims bit-exact, result counts and work counters equal."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from gpu_fixtures import CASES, assert_search_parity, case, device_index  # noqa: E402


@pytest.mark.parametrize("name,efs", [
    ("cfg1_10k_d32_m5", (1, 10, 16, 64, 100, 200, 400)),
    ("d128_m16", (10, 32, 64, 128, 200, 512)),
    ("d768_m32", (16, 128, 400)),
    ("d96_m8_generic", (16, 64, 200)),
    ("d20_m6_scalar", (16, 48, 100)),
    ("d64_m6_generic_v2", (1, 16, 64, 100)),
    ("d256_m12_generic_v4", (10, 64, 256)),
    ("d33_m4_scalar", (8, 24, 64)),
])
def test_search_parity(name, efs):
    c = case(name)
    dev = device_index(name)
    for ef in efs:
        assert_search_parity(dev, c["oracle"], c["q"], 10, ef)


@pytest.mark.parametrize("name", ["cfg1_10k_d32_m5", "d128_m16", "d768_m32"])
def test_staged_kernel_counters_and_options(name):
    """search2.cuh: hops and adjacency ids equal the oracle's; its lossy visited table may re-evaluate a node, so
    distance evaluations are >= the oracle's (and close); stage size / table size do not change results."""
    c = case(name)
    dev = device_index(name)
    q = c["q"][:300]
    oids, osims, ocounts, ost, _ = c["oracle"].search_batch(q, 10, ef=64)
    ok = ost[:, 3] == 0
    for rows, slots, tag, copy in ((0, 0, 0, 0), (4, 64, 0, 0), (8, 64, 32, 0), (16, 256, 0, 0), (32, 4096, 32, 0),
                                   (0, 0, 0, 1), (8, 64, 32, 1), (32, 256, 0, 1)):
        dev.set_option("search_impl", 2)
        dev.set_option("row_copy", copy)      # 1: cp.async row copies instead of bulk-async ones (128-d rows only)
        dev.set_option("stage_rows", rows)
        dev.set_option("recent_slots", slots)
        dev.set_option("recent_tag", tag)
        ids, sims, counts, st = dev.search_batch(q, 10, ef=64, stats=True)
        assert np.all(st[:, 3] == 4)
        assert np.array_equal(ids[ok], oids[ok]) and np.array_equal(sims[ok].view(np.uint32), osims[ok].view(np.uint32))
        assert np.array_equal(counts, ocounts)
        assert np.array_equal(st[ok, 1:3].astype(np.uint64), ost[ok, 1:3])
        assert np.all(st[ok, 0] >= ost[ok, 0])
        if slots == 0:
            assert st[ok, 0].sum() <= 1.5 * ost[ok, 0].sum()


@pytest.mark.parametrize("name,efs", [("cfg1_10k_d32_m5", (1, 16, 100)), ("d128_m16", (8, 64, 200, 512)), ("d768_m32", (16, 128))])
def test_latency_mode_parity(name, efs):
    """Calls with fewer queries than SMs take a 32-row stage (one staging round per adjacency chunk); both row-copy
    flavours: ids, sims, counts, hops and adjacency counters equal the oracle's."""
    c = case(name)
    dev = device_index(name)
    for ef in efs:
        for lo, n in ((0, 1), (1, 7), (8, 100)):
            q = c["q"][lo:lo + n]
            oids, osims, ocounts, ost, _ = c["oracle"].search_batch(q, 10, ef=ef)
            ok = ost[:, 3] == 0
            for copy in (0, 1):
                dev.set_option("row_copy", copy)
                ids, sims, counts = dev.search_batch(q, 10, ef=ef)
                assert np.array_equal(counts, ocounts)
                assert np.array_equal(ids[ok], oids[ok])
                assert np.array_equal(sims[ok].view(np.uint32), osims[ok].view(np.uint32))
                dev.set_option("search_impl", 2)      # same kernel, counters requested
                ids, sims, counts, st = dev.search_batch(q, 10, ef=ef, stats=True)
                dev.set_option("search_impl", 0)
                assert np.array_equal(ids[ok], oids[ok]) and np.array_equal(counts, ocounts)
                assert np.array_equal(st[ok, 1:3].astype(np.uint64), ost[ok, 1:3])
                assert np.all(st[ok, 0] >= ost[ok, 0])
    dev.set_option("row_copy", 0)


@pytest.mark.timeout(120)
@pytest.mark.parametrize("name,efs", [("cfg1_10k_d32_m5", (1, 16, 100)), ("d128_m16", (8, 64, 200, 512)), ("d768_m32", (16, 128))])
def test_cta_latency_kernel_parity(name, efs):
    """Option search_cta = 1: calls with at most two queries per SM run one query per CTA of 4 warps (search_knn2_cta_kernel,
    32-d / 128-d rows): same ids, sims and counts as the oracle; and the same again with the option back off."""
    c = case(name)
    dev = device_index(name)
    dev.set_option("search_cta", 1)
    for ef in efs:
        for lo, n in ((0, 1), (1, 7), (8, 100)):
            q = c["q"][lo:lo + n]
            oids, osims, ocounts, ost, _ = c["oracle"].search_batch(q, 10, ef=ef)
            ok = ost[:, 3] == 0
            ids, sims, counts = dev.search_batch(q, 10, ef=ef)
            assert np.array_equal(counts, ocounts)
            assert np.array_equal(ids[ok], oids[ok])
            assert np.array_equal(sims[ok].view(np.uint32), osims[ok].view(np.uint32))
    dev.set_option("search_cta", 0)
    q = c["q"][:40]
    oids, osims, ocounts, ost, _ = c["oracle"].search_batch(q, 10, ef=efs[-1])
    ok = ost[:, 3] == 0
    ids, sims, counts = dev.search_batch(q, 10, ef=efs[-1])
    assert np.array_equal(counts, ocounts) and np.array_equal(ids[ok], oids[ok])


@pytest.mark.parametrize("name", ["cfg1_10k_d32_m5", "d128_m16"])
def test_two_way_visited_sets_parity(name):
    """Option recent_ways = 2 (Recent<Way2>): same results, fewer or equal re-evaluations (not the default: it costs more
    than it saves, profiles/r2_experiments.md)."""
    c = case(name)
    dev = device_index(name)
    q = c["q"][:300]
    oids, osims, ocounts, ost, _ = c["oracle"].search_batch(q, 10, ef=64)
    ok = ost[:, 3] == 0
    evals = {}
    for ways in (1, 2):
        dev.set_option("search_impl", 2)
        dev.set_option("recent_ways", ways)
        ids, sims, counts, st = dev.search_batch(q, 10, ef=64, stats=True)
        assert np.array_equal(ids[ok], oids[ok]) and np.array_equal(sims[ok].view(np.uint32), osims[ok].view(np.uint32))
        assert np.array_equal(counts, ocounts)
        assert np.array_equal(st[ok, 1:3].astype(np.uint64), ost[ok, 1:3]) and np.all(st[ok, 0] >= ost[ok, 0])
        evals[ways] = int(st[ok, 0].sum())
    dev.set_option("search_impl", 0)
    dev.set_option("recent_ways", 1)
    assert evals[2] <= evals[1] * 1.01


def test_ef_up_to_1024():
    """The sixth register class (32 list registers per lane x 2) serves ef / ef_construction up to 1024 — search parity at
    ef 600 / 1024 and an exact build with ef_construction = 700."""
    import oracle
    import redis_hnsw_b200 as r

    c = case("d128_m16")
    dev = device_index("d128_m16")
    for ef in (600, 1024):
        # a list of 600 / 1024 entries holds 10-17 % of this 6000-node graph: equal sims between far-away nodes are met by
        # most queries, so the tie-free subset the ids are compared on is smaller than at the usual ef
        assert_search_parity(dev, c["oracle"], c["q"][:400], 10, ef, min_tie_free=0.05)
    n = 1500
    orc = oracle.Oracle(c["dim"], c["m"], 700)
    orc.add_batch(c["x"][:n], c["levels"][:n])
    d2 = r.DeviceIndex(c["dim"], c["m"], 700)
    d2.add_batch(c["x"][:n], c["levels"][:n], mode=r.BUILD_EXACT)
    go, gd = orc.export_graph(), d2.export_graph()
    assert np.array_equal(go["row_offs"], gd["row_offs"]) and np.array_equal(go["nbrs"], gd["nbrs"])


def test_ef_beyond_the_register_classes():
    """The reference accepts any EFCON (lib.rs:53, core.rs:322-346) and searches with it (core.rs:485).  Beyond the largest
    register class (1024 entries) the candidate list lives in memory (CandList<0>): search parity at ef 1500 / 3000, an
    exact build with ef_construction = 1500 identical to the oracle's, NODE.ADD / NODE.DEL on it, and a clean error past
    the supported maximum."""
    import oracle
    import redis_hnsw_b200 as r

    c = case("d128_m16")
    dev = device_index("d128_m16")
    for ef in (1500, 3000):
        assert_search_parity(dev, c["oracle"], c["q"][:300], 10, ef, min_tie_free=0.02)
        ids, sims = dev.search_level(c["q"][0], c["graph"]["entry"], ef, 0)
        oids, osims = c["oracle"].search_level(c["q"][0], c["graph"]["entry"], ef, 0)
        assert len(ids) == len(oids) and np.array_equal(np.sort(ids), np.sort(oids))
    n = 1300
    orc = oracle.Oracle(c["dim"], c["m"], 1500)
    orc.add_batch(c["x"][:n], c["levels"][:n])
    for mode in (r.BUILD_EXACT, r.BUILD_SPEC, r.BUILD_FAST):      # SPEC and FAST hand such an index to the EXACT stream
        d2 = r.DeviceIndex(c["dim"], c["m"], 1500)
        d2.add_batch(c["x"][:n], c["levels"][:n], mode=mode)
        go, gd = orc.export_graph(), d2.export_graph()
        assert np.array_equal(go["row_offs"], gd["row_offs"]) and np.array_equal(go["nbrs"], gd["nbrs"])
    for i in range(n, n + 10):
        assert d2.add(c["x"][i], int(c["levels"][i])) == orc.add(c["x"][i], int(c["levels"][i]))
    d2.delete(77)
    orc.delete(77)
    go, gd = orc.export_graph(), d2.export_graph()
    assert np.array_equal(go["row_offs"], gd["row_offs"]) and np.array_equal(go["nbrs"], gd["nbrs"])
    ids, sims, counts = d2.search_batch(c["q"][:50], 10)         # ef = 0 -> ef_construction = 1500 (core.rs:485)
    oids, osims, ocounts, ost, _ = orc.search_batch(c["q"][:50], 10)
    ok = ost[:, 3] == 0
    assert np.array_equal(counts, ocounts) and np.array_equal(ids[ok], oids[ok])
    r.DeviceIndex(32, 5, 65536).close()
    with pytest.raises(r.HNSWError, match="not supported"):
        r.DeviceIndex(32, 5, 65537)


def test_default_ef_is_ef_construction():
    """core.rs:485: search_knn always searches with ef = ef_construction."""
    c = case("cfg1_10k_d32_m5")
    dev = device_index("cfg1_10k_d32_m5")
    ids0, sims0, _ = dev.search_batch(c["q"][:200], 10, ef=0)
    ids1, sims1, _ = dev.search_batch(c["q"][:200], 10, ef=c["efc"])
    assert np.array_equal(ids0, ids1) and np.array_equal(sims0, sims1)


def test_k_larger_than_ef_and_single_query():
    c = case("d128_m16")
    dev = device_index("d128_m16")
    ids, sims, counts = dev.search_batch(c["q"][:64], 20, ef=8)  # core.rs:879 fewer than k results
    assert np.all(counts == 8)
    assert np.all(ids[:, 8:] == 0xFFFFFFFF) and np.all(np.isneginf(sims[:, 8:]))
    oi, osim = c["oracle"].search(c["q"][0], 20, ef=8)
    assert np.array_equal(ids[0, :8], oi) and np.array_equal(sims[0, :8], osim)
    i1, s1 = dev.search(c["q"][3], 10, ef=64)
    oi, osim = c["oracle"].search(c["q"][3], 10, ef=64)
    assert np.array_equal(i1, oi) and np.array_equal(s1.view(np.uint32), osim.view(np.uint32))


def test_search_level_parity_on_every_level():
    c = case("cfg1_10k_d32_m5")
    dev = device_index("cfg1_10k_d32_m5")
    g = c["graph"]
    top = g["max_layer"]
    assert top >= 3
    levels = g["levels"]
    for lv in range(0, top + 1):
        eps = np.nonzero(levels >= lv)[0][:3]
        for ep in eps:
            for ef in (1, 24, 100):
                ids, sims = dev.search_level(c["q"][7], int(ep), ef, lv)
                oi, osim = c["oracle"].search_level(c["q"][7], int(ep), ef, lv)
                assert np.array_equal(ids, oi) and np.array_equal(sims.view(np.uint32), osim.view(np.uint32))


def test_visited_overflow_takes_retry_pass():
    """A deliberately tiny visited table overflows; the retry pass (global-memory table) must give the same
    answers, and the stats flag says which queries went there."""
    c = case("d128_m16")
    dev = device_index("d128_m16")
    dev.set_option("visited_slots", 256)
    ids, sims, counts, st = assert_search_parity(dev, c["oracle"], c["q"][:400], 10, 64)
    assert (st[:, 3] & 1).sum() > 300


def test_block_and_cta_options_do_not_change_results():
    c = case("d128_m16")
    dev = device_index("d128_m16")
    base = dev.search_batch(c["q"][:300], 10, ef=64)
    for blk, ctas in ((32, 1), (128, 2), (256, 0)):
        dev.set_option("search_block", blk)
        dev.set_option("search_ctas_per_sm", ctas)
        got = dev.search_batch(c["q"][:300], 10, ef=64)
        assert all(np.array_equal(a, b) for a, b in zip(base, got))


def test_degree_overflow_rows_are_searched():
    """The reference does not bound the degree (core.rs:793-795); lists longer than the fixed row width live in
    chained overflow rows and must be walked in list order."""
    c = case("cfg1_10k_d32_m5")
    g = c["graph"]
    rows0 = np.concatenate([[0], np.cumsum(g["levels"] + 1)[:-1]])
    deg0 = (g["row_offs"][rows0 + 1] - g["row_offs"][rows0]).astype(int)
    assert deg0.max() > 10  # above m_max_0
    # force rows wider than W = 32 through a synthetic hub: give node 0 every node 1..99 as neighbour (mirrored)
    import copy

    import oracle as orc_mod
    import redis_hnsw_b200 as r

    orc = c["oracle"]
    lists = []
    row = 0
    for i in range(c["n"]):
        for lv in range(g["levels"][i] + 1):
            lists.append(list(g["nbrs"][int(g["row_offs"][row]):int(g["row_offs"][row + 1])]))
            row += 1
    hub = 0
    for j in range(1, 100):
        if j not in lists[rows0[hub]]:
            lists[rows0[hub]].append(np.uint32(j))
            lists[rows0[j]].append(np.uint32(hub))
    offs = np.zeros(len(lists) + 1, np.uint64)
    offs[1:] = np.cumsum([len(x) for x in lists])
    g2 = dict(g, row_offs=offs, nbrs=np.array([v for x in lists for v in x], np.uint32))
    o2 = orc_mod.Oracle(c["dim"], c["m"], c["efc"])
    o2.import_graph(c["x"], g2)
    dev = r.DeviceIndex(c["dim"], c["m"], c["efc"])
    dev.load_graph(c["x"], g2)
    assert np.array_equal(dev.node_neighbors(hub, 0), np.array(lists[rows0[hub]], np.uint32))
    assert len(dev.node_neighbors(hub, 0)) >= 99
    q = np.concatenate([c["x"][:50] + 0.01, c["q"][:200]]).astype(np.float32)  # queries near the hub's neighbourhood
    assert_search_parity(dev, o2, q, 10, 64)
    g3 = dev.export_graph()
    assert np.array_equal(g3["row_offs"], g2["row_offs"]) and np.array_equal(g3["nbrs"], g2["nbrs"])


def test_export_roundtrip_and_getters():
    c = case("d96_m8_generic")
    dev = device_index("d96_m8_generic")
    g = dev.export_graph()
    for key in ("levels", "row_offs", "nbrs"):
        assert np.array_equal(g[key], c["graph"][key]), key
    assert g["entry"] == c["graph"]["entry"] and g["max_layer"] == c["graph"]["max_layer"]
    assert np.array_equal(dev.export_vectors(), c["x"])  # lane-permuted slab maps back to natural order
    p = dev.params()
    op = c["oracle"].params()
    for key in ("data_dim", "m", "m_max", "m_max_0", "ef_construction", "node_count", "max_layer", "enterpoint"):
        assert p[key] == op[key], key
    assert p["level_mult"] == op["level_mult"]
    for i in (0, 5, 77):
        assert np.array_equal(dev.node_vector(i), c["x"][i])
        assert dev.node_level(i) == c["oracle"].node_level(i)
        for lv in range(dev.node_level(i) + 1):
            assert np.array_equal(dev.node_neighbors(i, lv), c["oracle"].node_neighbors(i, lv))


def test_errors_and_empty_index():
    import redis_hnsw_b200 as r

    dev = r.DeviceIndex(32, 5, 100)
    ids, sims = dev.search(np.zeros(32, np.float32), 5)  # core.rs:481-483
    assert len(ids) == 0
    with pytest.raises(r.HNSWError, match="data dimension: 31 does not match Index"):
        dev.search(np.zeros(31, np.float32), 5)
    ids, sims, counts = dev.search_batch(np.zeros((4, 32), np.float32), 5)
    assert np.all(counts == 0) and np.all(ids == 0xFFFFFFFF)
    with pytest.raises(r.HNSWError, match="not supported"):
        dev.search_batch(np.zeros((4, 32), np.float32), 5, ef=70000)
    with pytest.raises(r.HNSWError):
        r.DeviceIndex(0, 5, 100)
