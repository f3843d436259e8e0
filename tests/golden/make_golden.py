"""Generates tests/golden/hnsw_golden.npz — committed golden vectors for the hot path.

The reference (Rust) cannot run in this image, so the vectors come from the CPU oracle (oracle/hnsw_oracle.cpp), which
is itself pinned on the reference's own known-answer tests (src/hnsw/metrics_tests.rs:4-33, src/hnsw/core_tests.rs:7-81;
tests/test_oracle_kat.py).  The fixtures freeze its output so that (i) an accidental change of the oracle is caught on
CPU (tests/test_golden_cpu.py) and (ii) the CUDA path is checked against committed numbers, not only against a live
oracle (tests/test_gpu_golden.py).

    python tests/golden/make_golden.py          # rewrites hnsw_golden.npz next to this file
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from redis_hnsw_b200 import data  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hnsw_golden.npz")

# name -> (n, dim, m, ef_construction, dataset, n_queries, ef, k, n_delete)
GRAPHS = {
    "d32_m5": (700, 32, 5, 100, "uniform", 48, 100, 10, 25),      # BASELINE configs[0] parameters (AVX-order metric)
    "d128_m16": (400, 128, 16, 200, "lowrank16", 32, 64, 10, 10),  # configs[1] parameters
    "d20_m6": (400, 20, 6, 48, "uniform", 32, 48, 5, 10),          # dim % 32 != 0 -> scalar metric (metrics.rs:79-84)
}


def main():
    out = {}
    # metric: random pairs, bits of euclidean() per dimension (metrics.rs:14-84)
    rng = np.random.default_rng(2024)
    for dim in (32, 128, 768, 20, 33):
        rows = 16 if dim > 128 else 64
        a = rng.standard_normal((rows, dim)).astype(np.float32)
        b = rng.standard_normal((rows, dim)).astype(np.float32)
        out["metric_a_%d" % dim] = a
        out["metric_b_%d" % dim] = b
        out["metric_bits_%d" % dim] = oracle.euclidean_batch(a, b).view(np.uint32)
    for name, (n, dim, m, efc, ds, nq, ef, k, n_del) in GRAPHS.items():
        if ds == "uniform":
            x, q = data.uniform(n, dim, seed=321, n_queries=nq)
        else:
            x, q = data.lowrank(n, dim, r=int(ds[7:]), seed=321, n_queries=nq)
        levels = data.draw_levels(n, m, seed=17)
        orc = oracle.Oracle(dim, m, efc)
        orc.add_batch(x, levels)
        g = orc.export_graph()
        ids, sims, counts, st, _ = orc.search_batch(q, k, ef=ef)
        out[name + "_params"] = np.asarray([n, dim, m, efc, ef, k], np.int64)
        out[name + "_x"] = x
        out[name + "_q"] = q
        out[name + "_levels"] = levels.astype(np.int8)
        out[name + "_row_offs"] = g["row_offs"].astype(np.uint32)
        out[name + "_nbrs"] = g["nbrs"].astype(np.uint16)
        out[name + "_entry"] = np.asarray([g["entry"], g["max_layer"]], np.int64)
        out[name + "_ids"] = ids.astype(np.uint16)
        out[name + "_sim_bits"] = sims.view(np.uint32)
        out[name + "_counts"] = counts.astype(np.uint8)
        out[name + "_stats"] = st.astype(np.uint32)          # n_dist, n_adj, n_hops, n_ties per query
        # NODE.DEL stream on the same graph, graph after it
        victims = np.random.default_rng(9).permutation(n)[:n_del].astype(np.uint32)
        victims[0] = g["entry"]
        for v in victims:
            orc.delete(int(v))
        g2 = orc.export_graph()
        out[name + "_victims"] = victims.astype(np.uint16)
        out[name + "_del_levels"] = g2["levels"].astype(np.int8)
        out[name + "_del_row_offs"] = g2["row_offs"].astype(np.uint32)
        out[name + "_del_nbrs"] = g2["nbrs"].astype(np.uint16)
        out[name + "_del_entry"] = np.asarray([g2["entry"], g2["max_layer"]], np.int64)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
