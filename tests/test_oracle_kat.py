"""Pins the CPU oracle against every known-answer test the reference holds for the hot path
(reference: src/hnsw/metrics_tests.rs:4-33, src/hnsw/core_tests.rs:7-81)."""
import numpy as np
import pytest

EPS = np.finfo(np.float32).eps  # f32::EPSILON


# ---- metrics_tests.rs
def test_diff_is_zero(oracle_mod):  # metrics_tests.rs:4-9
    v1 = np.full(512, 1.0, np.float32)
    v2 = np.full(512, 1.0, np.float32)
    assert abs(oracle_mod.sim_avx(v1, v2) - 0.0) < EPS
    assert abs(oracle_mod.sim_avx(v1, v2, portable=True) - 0.0) < EPS
    assert abs(oracle_mod.sim_scalar(v1, v2) - 0.0) < EPS


def test_diff_is_512(oracle_mod):  # metrics_tests.rs:12-17
    v1 = np.zeros(512, np.float32)
    v2 = np.ones(512, np.float32)
    assert abs(oracle_mod.sim_avx(v1, v2) - -512.0) < EPS
    assert abs(oracle_mod.sim_avx(v1, v2, portable=True) - -512.0) < EPS
    assert abs(oracle_mod.sim_scalar(v1, v2) - -512.0) < EPS


def test_diff_is_512_2_x512(oracle_mod):  # metrics_tests.rs:20-25
    v1 = np.zeros(512, np.float32)
    v2 = np.full(512, 512.0, np.float32)
    assert abs(oracle_mod.sim_avx(v1, v2) - -134217728.0) < EPS
    assert abs(oracle_mod.sim_scalar(v1, v2) - -134217728.0) < EPS


def test_diff_non_x32(oracle_mod):  # metrics_tests.rs:28-33
    v1 = np.zeros(33, np.float32)
    v2 = np.ones(33, np.float32)
    assert abs(oracle_mod.sim_scalar(v1, v2) - -33.0) < EPS
    assert abs(oracle_mod.euclidean(v1, v2) - -33.0) < EPS  # dispatch: 33 % 32 != 0 -> scalar path


def test_avx_intrinsic_equals_lanewise_restatement(oracle_mod):
    """The AVX2 intrinsic transcription and the lane-by-lane fmaf restatement agree bit for bit."""
    rng = np.random.default_rng(7)
    for dim in (32, 64, 128, 768, 1024):
        a = rng.standard_normal((200, dim)).astype(np.float32)
        b = rng.standard_normal((200, dim)).astype(np.float32)
        for i in range(200):
            x = np.float32(oracle_mod.sim_avx(a[i], b[i]))
            y = np.float32(oracle_mod.sim_avx(a[i], b[i], portable=True))
            assert x.tobytes() == y.tobytes()


def test_identical_vectors_give_negative_zero(oracle_mod):
    v = np.arange(128, dtype=np.float32)
    s = np.float32(oracle_mod.euclidean(v, v))
    assert s == 0.0 and np.signbit(s)  # `-res` of +0.0 (metrics.rs:75)


# ---- core_tests.rs::hnsw_test
@pytest.mark.parametrize("seed", [0, 1, 2, 3, 42, 1234])
def test_hnsw_test(oracle_mod, seed):
    n, dim = 100, 4
    idx = oracle_mod.Oracle(dim, 5, 16)  # core_tests.rs:12
    p = idx.params()
    assert p["data_dim"] == dim and p["m"] == 5 and p["ef_construction"] == 16  # :13-16
    assert p["node_count"] == 0 and p["max_layer"] == 0 and p["enterpoint"] == -1  # :17-19
    assert p["m_max"] == 5 and p["m_max_0"] == 10  # core.rs:335-336

    levels = oracle_mod.draw_levels(n, 5, seed)
    names = {}
    for i in range(n):  # :24-28
        names[idx.add(np.full(dim, float(i), np.float32), levels[i])] = "node%d" % i
    assert idx.node_count == n  # :41
    assert idx.params()["enterpoint"] >= 0  # :42

    ids, sims = idx.search(np.full(4, 10.0, np.float32), 5)  # :45-46 (ef = ef_construction)
    assert len(ids) == 5  # :47
    assert abs(sims[0] - 0.0) < EPS and names[int(ids[0])] == "node10"  # :48-49
    assert abs(sims[1] - -4.0) < EPS and abs(sims[2] - -4.0) < EPS  # :50-51
    assert abs(sims[3] - -16.0) < EPS and abs(sims[4] - -16.0) < EPS  # :52-53

    # delete every node, checking that nothing dangles (:56-80)
    for i in range(n):
        idx.delete(i)
        assert idx.node_count == n - i - 1
        assert idx.node_level(i) == -1
        for j in range(i + 1, n):
            for lv in range(idx.node_n_levels(j)):
                assert i not in idx.node_neighbors(j, lv)
    assert idx.params()["enterpoint"] == -1


def test_errors_and_empty(oracle_mod):
    idx = oracle_mod.Oracle(4, 5, 16)
    ids, sims = idx.search(np.zeros(4, np.float32), 5)  # core.rs:481-483 empty index -> Ok(vec![])
    assert len(ids) == 0
    with pytest.raises(oracle_mod.HNSWError, match="data dimension: 3 does not match Index"):
        idx.add(np.zeros(3, np.float32), 0)
    with pytest.raises(oracle_mod.HNSWError, match="data dimension: 5 does not match Index"):
        idx.search(np.zeros(5, np.float32), 1)
    with pytest.raises(oracle_mod.HNSWError, match="does not exist"):
        idx.delete(0)


def test_symmetric_graph_and_overflow_degree(oracle_mod):
    """Facts the survey measured on the reference semantics: the graph is strictly undirected and the
    level-0 degree is NOT bounded by m_max_0 (core.rs:793-795 adds without a cap check)."""
    rng = np.random.default_rng(123)
    n, dim, m = 3000, 32, 5
    x = rng.random((n, dim), dtype=np.float32)
    idx = oracle_mod.Oracle(dim, m, 100)
    idx.add_batch(x, oracle_mod.draw_levels(n, m, 42))
    g = idx.export_graph()
    offs, nbrs, levels = g["row_offs"], g["nbrs"], g["levels"]
    row = 0
    adj = {}
    for i in range(n):
        for lv in range(levels[i] + 1):
            adj[(i, lv)] = nbrs[int(offs[row]):int(offs[row + 1])]
            row += 1
    maxdeg0 = 0
    for (i, lv), lst in adj.items():
        assert len(set(lst.tolist())) == len(lst)  # no duplicate edges
        for j in lst:
            assert i in adj[(int(j), lv)]  # mirrored
        if lv == 0:
            maxdeg0 = max(maxdeg0, len(lst))
    assert maxdeg0 > 2 * m


def test_k_larger_than_ef_returns_ef(oracle_mod):
    rng = np.random.default_rng(5)
    x = rng.random((500, 32), dtype=np.float32)
    idx = oracle_mod.Oracle(32, 5, 50)
    idx.add_batch(x, oracle_mod.draw_levels(500, 5, 1))
    ids, _ = idx.search(x[0], 20, ef=8)  # core.rs:879: fewer than k results when k > ef
    assert len(ids) == 8


def test_export_import_roundtrip(oracle_mod):
    rng = np.random.default_rng(9)
    x = rng.random((800, 32), dtype=np.float32)
    a = oracle_mod.Oracle(32, 6, 40)
    a.add_batch(x, oracle_mod.draw_levels(800, 6, 3))
    g = a.export_graph()
    b = oracle_mod.Oracle(32, 6, 40)
    b.import_graph(x, g)
    q = rng.random((50, 32), dtype=np.float32)
    ia, sa, ca, _, _ = a.search_batch(q, 10, ef=30)
    ib, sb, cb, _, _ = b.search_batch(q, 10, ef=30, threads=3)
    assert np.array_equal(ia, ib) and np.array_equal(sa.view(np.uint32), sb.view(np.uint32)) and np.array_equal(ca, cb)
