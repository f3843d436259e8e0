"""The CUDA path through the C ABI against the committed golden vectors (tests/golden/) — no oracle in the loop:
metric bits (both reference reduction orders), EXACT NODE.ADD stream -> the golden graph, HNSW.SEARCH ids / sim bits
/ counts / work counters on it, NODE.DEL stream -> the golden graph after deletion."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import golden_util as G  # noqa: E402


def test_reference_metric_kats_from_json():
    import redis_hnsw_b200 as r

    for kat in G.kats()["metric"]:
        a = np.full((1, kat["dim"]), kat["a_fill"], np.float32)
        b = np.full((1, kat["dim"]), kat["b_fill"], np.float32)
        assert float(r.l2_batch(a, b)[0]) == kat["expect"], kat["ref"]


@pytest.mark.parametrize("dim", G.METRIC_DIMS)
def test_metric_bits(dim):
    import redis_hnsw_b200 as r

    z = G.load()
    got = r.l2_batch(z["metric_a_%d" % dim], z["metric_b_%d" % dim]).view(np.uint32)
    assert np.array_equal(got, z["metric_bits_%d" % dim])


@pytest.mark.parametrize("name", G.GRAPHS)
def test_add_search_delete_streams(name):
    import redis_hnsw_b200 as r

    z = G.load()
    n, dim, m, efc, ef, k = (int(v) for v in z[name + "_params"])
    x, q = z[name + "_x"], z[name + "_q"]
    dev = r.DeviceIndex(dim, m, efc)
    dev.add_batch(x, z[name + "_levels"].astype(np.int32), mode=r.BUILD_EXACT)
    got = dev.export_graph()
    assert G.same_graph(got, G.graph(z, name)), "EXACT NODE.ADD stream does not reproduce the golden graph"
    tie_free = z[name + "_stats"][:, 3] == 0
    assert tie_free.mean() > 0.9
    for stats in (True, False):                                   # exact-visited kernel / TMA-staged kernel
        res = dev.search_batch(q, k, ef=ef, stats=stats)
        ids, sims, counts = res[0], res[1], res[2]
        assert np.array_equal(ids[tie_free].astype(np.uint16), z[name + "_ids"][tie_free])
        assert np.array_equal(sims[tie_free].view(np.uint32), z[name + "_sim_bits"][tie_free])
        assert np.array_equal(counts, z[name + "_counts"])
        if stats:
            assert np.array_equal(res[3][tie_free, :3], z[name + "_stats"][tie_free, :3])
    for v in z[name + "_victims"]:
        dev.delete(int(v))
    assert G.same_graph(dev.export_graph(), G.graph(z, name, "del_")), "NODE.DEL stream does not reproduce the golden graph"
