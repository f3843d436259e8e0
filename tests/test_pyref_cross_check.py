"""Two independent restatements of the reference must agree: oracle/hnsw_oracle.cpp (the checker of the GPU tests) and
tests/pyref_hnsw.py (a literal Python transliteration of src/hnsw/core.rs with Rust's BinaryHeap sift rules).
Graphs list by list IN ORDER, touched sets, deletes, enterpoint bookkeeping and search results — on continuous data and
on grid data where most comparisons are ties (which is where heap internals decide the outcome)."""
import numpy as np
import pytest

import oracle
from pyref_hnsw import PyRefIndex, RustHeap, sim_func_avx_euc, sim_func_euc
from redis_hnsw_b200 import data


def _levels(n, m, seed):
    return data.draw_levels(n, m, seed=seed)


def _same_graph(orc, ref, n_ids):
    p = orc.params()
    assert p["max_layer"] == ref.max_layer
    assert p["node_count"] == ref.node_count
    ep = p["enterpoint"]
    assert (None if ep in (-1, 0xFFFFFFFF, None) else ep) == ref.enterpoint
    g = orc.export_graph()
    row = 0
    for i in range(n_ids):
        lv = int(g["levels"][i])
        if i not in ref.nodes:
            assert lv == -1
            continue
        top = ref.levels[i] if ref.levels[i] <= ref.max_layer or i == ref.enterpoint else ref.levels[i]
        assert lv == top, (i, lv, top)
        for l in range(lv + 1):
            got = g["nbrs"][int(g["row_offs"][row]):int(g["row_offs"][row + 1])].tolist()
            assert got == ref.adjacency(i, l), "node %d level %d: oracle %r pyref %r" % (i, l, got, ref.adjacency(i, l))
            row += 1


def _dataset(kind, n, dim, seed):
    if kind == "uniform":
        return data.uniform(n, dim, seed=seed, n_queries=40)
    rng = np.random.default_rng(seed)              # grid: small integer coordinates -> masses of exactly equal sims
    x = rng.integers(0, 4, size=(n, dim)).astype(np.float32)
    q = rng.integers(0, 4, size=(40, dim)).astype(np.float32)
    return x, q


@pytest.mark.parametrize("kind,n,dim,m,efc", [
    ("uniform", 400, 4, 5, 16),       # the reference KAT's parameters (core_tests.rs:7-53)
    ("uniform", 300, 20, 6, 12),
    ("uniform", 300, 8, 8, 4),        # ef_construction < m: the extension of select_neighbors is not redundant
    ("grid", 300, 4, 5, 16),          # ties everywhere
    ("grid", 250, 3, 4, 6),
])
def test_oracle_and_python_transliteration_agree(kind, n, dim, m, efc):
    x, q = _dataset(kind, n, dim, seed=11)
    levels = _levels(n, m, seed=12)
    orc = oracle.Oracle(dim, m, efc)
    ref = PyRefIndex(dim, m, efc)
    for i in range(n):
        a = orc.add(x[i], int(levels[i]))
        b = ref.add_node(x[i], int(levels[i]))
        assert a == b == i
        assert sorted(int(t) for t in orc.touched()) == sorted(ref.last_updated), "update_fn set after insert %d" % i
        if i in (1, 2, 5, 50, n // 2):
            _same_graph(orc, ref, i + 1)
    _same_graph(orc, ref, n)
    for ef in (1, 3, efc, 40):
        for j in range(len(q)):
            oi, osim = orc.search(q[j], 7, ef=ef)
            ri, rsim = ref.search_knn(q[j], 7, ef=ef)
            assert oi.tolist() == ri, (ef, j, oi.tolist(), ri)
            assert np.array_equal(np.asarray(osim, np.float32).view(np.uint32), np.asarray(rsim, np.float32).view(np.uint32))
    rng = np.random.default_rng(5)
    victims = rng.permutation(n)[:60].tolist()
    if ref.enterpoint not in victims:
        victims[7] = ref.enterpoint               # exercise the enterpoint replacement (core.rs:449-472)
    for v in victims:
        orc.delete(int(v))
        ref.delete_node(int(v))
        assert sorted(int(t) for t in orc.touched()) == sorted(ref.last_updated), "update_fn set after delete %d" % v
    _same_graph(orc, ref, n)
    for j in range(len(q)):
        oi, osim = orc.search(q[j], 5, ef=efc)
        ri, rsim = ref.search_knn(q[j], 5, ef=efc)
        assert oi.tolist() == ri
    # inserts after deletes (ids keep growing; a freed enterpoint/top layer is handled like core.rs:587-596)
    x2, _ = _dataset(kind, 30, dim, seed=13)
    lv2 = _levels(30, m, seed=14)
    for i in range(30):
        assert orc.add(x2[i], int(lv2[i])) == ref.add_node(x2[i], int(lv2[i]))
    _same_graph(orc, ref, n + 30)


def test_rust_binary_heap_model_against_known_traces():
    """The heap model on hand-checked cases: push keeps earlier equal elements above later ones; pop walks the hole to
    the bottom preferring the RIGHT child on ties and sifts back up."""
    h = RustHeap()
    for item in [(1.0, "a"), (1.0, "b"), (1.0, "c"), (2.0, "d"), (1.0, "e")]:
        h.push(item)
    assert [n for _, n in h.into_vec()] == ["d", "a", "c", "b", "e"]
    assert h.pop()[1] == "d"
    # after removing d: last element e goes to the root, hole walks to the right child on the tie (a <= c), e lands at a leaf
    assert [n for _, n in h.into_vec()] == ["c", "a", "e", "b"]
    assert [h.pop()[1] for _ in range(4)] == ["c", "e", "a", "b"]
    r = RustHeap(reverse=True)
    for item in [(3.0, 1), (1.0, 2), (2.0, 3), (1.0, 4)]:
        r.push(item)
    assert r.peek() == (1.0, 2)
    assert r.pop() == (1.0, 2) and r.pop() == (1.0, 4) and r.pop() == (2.0, 3)


@pytest.mark.parametrize("dim,rows", [(32, 60), (128, 40), (768, 8)])
def test_avx_fma_metric_order_against_exact_rational_emulation(dim, rows):
    """metrics.rs:48-77 restated with exact rational arithmetic (every FMA rounded once, adds in the reference's tree)
    must give the oracle's bits — the same bits the CUDA kernels are tested against (tests/test_gpu_metric.py)."""
    rng = np.random.default_rng(dim)
    a = (rng.standard_normal((rows, dim)) * 3).astype(np.float32)
    b = (rng.standard_normal((rows, dim)) * 3).astype(np.float32)
    got = oracle.euclidean_batch(a, b)
    want = np.array([sim_func_avx_euc(a[i], b[i]) for i in range(rows)], dtype=np.float32)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    # and the scalar path on a non-x32 dimension (metrics.rs:79-84)
    a2, b2 = a[:, :dim - 1], b[:, :dim - 1]
    got2 = oracle.euclidean_batch(np.ascontiguousarray(a2), np.ascontiguousarray(b2))
    want2 = np.array([sim_func_euc(a2[i], b2[i]) for i in range(rows)], dtype=np.float32)
    assert np.array_equal(got2.view(np.uint32), want2.view(np.uint32))


@pytest.mark.parametrize("level_seed", [0, 1, 2, 3, 4, 5, 6, 7])
def test_python_transliteration_passes_the_reference_kat(level_seed):
    """The reference's own test (core_tests.rs:7-81) replayed on the transliteration, for several level draws (the
    reference seeds its RNG from entropy, so the KAT must hold for any): 100 nodes [i; 4], m = 5, efCon = 16; query
    [10; 4], k = 5 -> sims 0, -4, -4, -16, -16 with node10 first; then every node is deleted in insertion order and no
    layer set or adjacency list may still mention it."""
    n, dim = 100, 4
    ref = PyRefIndex(dim, 5, 16)
    orc = oracle.Oracle(dim, 5, 16)
    levels = data.draw_levels(n, 5, seed=level_seed)
    for i in range(n):
        v = np.full(dim, float(i), np.float32)
        assert ref.add_node(v, int(levels[i])) == orc.add(v, int(levels[i])) == i
    assert ref.node_count == n and ref.enterpoint is not None
    ids, sims = ref.search_knn(np.full(dim, 10.0, np.float32), 5)
    assert len(ids) == 5 and ids[0] == 10
    assert [float(s) for s in sims] == [0.0, -4.0, -4.0, -16.0, -16.0]
    oi, osim = orc.search(np.full(dim, 10.0, np.float32), 5)
    assert oi.tolist() == ids and [float(s) for s in osim] == [float(s) for s in sims]
    for i in range(n):
        ref.delete_node(i)
        orc.delete(i)
        assert ref.node_count == n - i - 1 and i not in ref.nodes
        assert all(i not in layer for layer in ref.layers)
        assert all(i not in lst for node in ref.nodes.values() for lst in node.neighbors)
        assert orc.params()["node_count"] == ref.node_count
    assert ref.enterpoint is None
