"""The CPU oracle replays the committed golden vectors (tests/golden/): reference KATs by value, metric bits, the
NODE.ADD stream (graph list by list), HNSW.SEARCH results and work counters, and the NODE.DEL stream."""
import numpy as np
import pytest

import golden_util as G

EPS = np.finfo(np.float32).eps


def test_reference_metric_kats_from_json(oracle_mod):
    for kat in G.kats()["metric"]:
        a = np.full(kat["dim"], kat["a_fill"], np.float32)
        b = np.full(kat["dim"], kat["b_fill"], np.float32)
        assert abs(oracle_mod.euclidean(a, b) - kat["expect"]) < EPS, kat["ref"]


def test_reference_core_kat_from_json(oracle_mod):
    kat = G.kats()["core"]
    idx = oracle_mod.Oracle(kat["dim"], kat["m"], kat["ef_construction"])
    lv = oracle_mod.draw_levels(kat["n"], kat["m"], 5)
    for i in range(kat["n"]):
        idx.add(np.full(kat["dim"], float(i), np.float32), lv[i])
    ids, sims = idx.search(np.full(kat["dim"], kat["query_fill"], np.float32), kat["k"])
    assert [float(s) for s in sims] == kat["expect_sims"]
    assert "node%d" % ids[0] == kat["expect_top_name"]


@pytest.mark.parametrize("dim", G.METRIC_DIMS)
def test_metric_bits(oracle_mod, dim):
    z = G.load()
    got = oracle_mod.euclidean_batch(z["metric_a_%d" % dim], z["metric_b_%d" % dim]).view(np.uint32)
    assert np.array_equal(got, z["metric_bits_%d" % dim])


@pytest.mark.parametrize("name", G.GRAPHS)
def test_add_search_delete_streams(oracle_mod, name):
    z = G.load()
    n, dim, m, efc, ef, k = (int(v) for v in z[name + "_params"])
    x, q = z[name + "_x"], z[name + "_q"]
    idx = oracle_mod.Oracle(dim, m, efc)
    idx.add_batch(x, z[name + "_levels"].astype(np.int32))
    assert G.same_graph(idx.export_graph(), G.graph(z, name))
    ids, sims, counts, st, _ = idx.search_batch(q, k, ef=ef)
    assert np.array_equal(ids.astype(np.uint16), z[name + "_ids"])
    assert np.array_equal(sims.view(np.uint32), z[name + "_sim_bits"])
    assert np.array_equal(counts, z[name + "_counts"])
    assert np.array_equal(st.astype(np.uint32), z[name + "_stats"])
    for v in z[name + "_victims"]:
        idx.delete(int(v))
    assert G.same_graph(idx.export_graph(), G.graph(z, name, "del_"))
