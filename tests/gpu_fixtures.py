"""Oracle-built graphs shared by the GPU parity tests (built once per session on the host CPU)."""
import functools

import numpy as np

import oracle
from redis_hnsw_b200 import data

# name -> (n, dim, m, ef_construction, dataset, n_queries)
CASES = {
    "cfg1_10k_d32_m5": (10000, 32, 5, 100, "uniform", 2000),      # BASELINE configs[0]
    "d128_m16": (6000, 128, 16, 200, "lowrank16", 1500),          # configs[1] parameters, small N
    "d768_m32": (1200, 768, 32, 120, "lowrank32", 300),           # configs[2] shape, small N / efCon
    "cfg3_5k_d768_m32_efc400": (5000, 768, 32, 400, "lowrank32", 300),   # configs[2] parameters in full (M=32, efCon=400)
    "d96_m8_generic": (3000, 96, 8, 64, "uniform", 500),          # dim % 32 == 0 without a specialised kernel
    "d20_m6_scalar": (2000, 20, 6, 48, "uniform", 500),           # dim % 32 != 0 -> reference scalar path
    "d64_m6_generic_v2": (2000, 64, 6, 48, "uniform", 400),       # generic kind, 2-wide row loads (dim/32 % 4 == 2)
    "d256_m12_generic_v4": (1500, 256, 12, 64, "lowrank16", 300),  # generic kind, 4-wide row loads
    "d33_m4_scalar": (800, 33, 4, 24, "uniform", 200),            # the reference's own non-x32 KAT dimension (metrics_tests.rs:28)
}


@functools.lru_cache(maxsize=None)
def case(name):
    n, dim, m, efc, ds, nq = CASES[name]
    if ds == "uniform":
        x, q = data.uniform(n, dim, seed=123, n_queries=nq)
    else:
        x, q = data.lowrank(n, dim, r=int(ds[7:]), seed=123, n_queries=nq)
    levels = data.draw_levels(n, m, seed=42)
    orc = oracle.Oracle(dim, m, efc)
    orc.add_batch(x, levels)
    return dict(x=x, q=q, levels=levels, oracle=orc, graph=orc.export_graph(), dim=dim, m=m, efc=efc, n=n)


def device_index(name):
    import redis_hnsw_b200 as r

    c = case(name)
    idx = r.DeviceIndex(c["dim"], c["m"], c["efc"])
    idx.load_graph(c["x"], c["graph"])
    return idx


def assert_search_parity(dev, orc, q, k, ef, check_stats=True, min_tie_free=0.9):
    """IDs and result counts bit-exact, sims bit-exact, work counters equal — on every query where the oracle
    met no tie between different nodes (the reference leaves tie order to BinaryHeap internals)."""
    ids, sims, counts, st = dev.search_batch(q, k, ef=ef, stats=True)
    oids, osims, ocounts, ost, _ = orc.search_batch(q, k, ef=ef)
    tie_free = ost[:, 3] == 0
    assert tie_free.mean() > min_tie_free
    assert np.array_equal(counts[tie_free], ocounts[tie_free])
    assert np.array_equal(ids[tie_free], oids[tie_free])
    assert np.array_equal(sims[tie_free].view(np.uint32), osims[tie_free].view(np.uint32))
    if check_stats:
        assert np.array_equal(st[tie_free, :3].astype(np.uint64), ost[tie_free, :3])
    # queries with ties: same result count and the same sims (ids may be permuted among equals)
    t = ~tie_free
    if t.any():
        assert np.array_equal(counts[t], ocounts[t])
    # the default (no work counters) call takes the TMA-staged kernel where one exists for the dimension
    ids2, sims2, counts2 = dev.search_batch(q, k, ef=ef)
    assert np.array_equal(counts2[tie_free], ocounts[tie_free])
    assert np.array_equal(ids2[tie_free], oids[tie_free])
    assert np.array_equal(sims2[tie_free].view(np.uint32), osims[tie_free].view(np.uint32))
    return ids, sims, counts, st
