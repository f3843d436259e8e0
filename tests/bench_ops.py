#!/usr/bin/env python
"""Secondary measurements on one GPU (not the headline bench): per-command latencies of the reference's three mutating /
querying operators on a big index, GPU vs the CPU oracle on the SAME graph.

  HNSW.SEARCH   one query per call (hnsw_index_search): the reference's command granularity (lib.rs:462-496)
  HNSW.NODE.ADD one node per call (hnsw_index_add, EXACT) and as one EXACT stream (hnsw_index_add_batch)
  HNSW.NODE.DEL one node per call (hnsw_index_delete)

    python tests/bench_ops.py [--workload 1Mx128_M16_efc200] > gpurun_out/ops.json
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (workload table + data)


def pct(a, p):
    return float(np.percentile(np.asarray(a) * 1e6, p))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="1Mx128_M16_efc200")
    ap.add_argument("--ef", type=int, default=64)
    ap.add_argument("--n-search", type=int, default=2000)
    ap.add_argument("--n-add", type=int, default=1000)
    ap.add_argument("--n-del", type=int, default=300)
    ap.add_argument("--cpu-add", type=int, default=200)
    ap.add_argument("--cpu-del", type=int, default=100)
    ap.add_argument("--only", default="all", choices=["all", "search"])
    ap.add_argument("--option", action="append", default=[], help="name=value library option set after the build")
    args = ap.parse_args()
    import oracle
    import redis_hnsw_b200 as r
    from redis_hnsw_b200 import data

    wl = args.workload
    n, dim, m, efc, ds, rr = bench.WORKLOADS[wl]
    extra_n = 2 * args.n_add + args.cpu_add
    x, q_all, levels = bench.make_data(wl, args.n_search + extra_n)   # new points: same distribution as the data
    q, extra = q_all[:args.n_search], q_all[args.n_search:]
    lv_extra = data.draw_levels(extra_n, m, seed=4242)
    dev = r.DeviceIndex(dim, m, efc)
    dev.reserve(n + extra_n)
    t0 = time.perf_counter()
    dev.add_batch(x, levels, mode=r.BUILD_FAST)
    build_s = time.perf_counter() - t0
    out = {"workload": wl, "n": n, "dim": dim, "M": m, "ef_construction": efc, "ef_search": args.ef,
           "fast_build_s": build_s}

    for opt in args.option:
        name, val = opt.split("=")
        dev.set_option(name, int(val))
    out["options"] = args.option
    orc = oracle.Oracle(dim, m, efc)
    orc.import_graph(x, dev.export_graph())

    # ---- HNSW.SEARCH, one query per call
    def single(n_calls):
        for i in range(20):
            dev.search(q[i], 10, ef=args.ef)
        lat, res = [], []
        for i in range(n_calls):
            t = time.perf_counter()
            ids, sims = dev.search(q[i], 10, ef=args.ef)
            lat.append(time.perf_counter() - t)
            res.append(ids)
        return lat, res

    lat, res = single(args.n_search)                 # default: cp.async row staging (search2.cuh, COPY = 1)
    dev.set_option("row_copy", 0)                    # bulk-async (TMA) row copies
    lat_bulk, res_bulk = single(args.n_search)
    dev.set_option("row_copy", 1)
    clat, cres = [], []
    for i in range(args.n_search):
        t = time.perf_counter()
        oids, osims = orc.search(q[i], 10, ef=args.ef)
        clat.append(time.perf_counter() - t)
        cres.append(oids)
    same_spec = all(np.array_equal(a, b) for a, b in zip(res, res_bulk))
    same_orc = float(np.mean([np.array_equal(a, b) for a, b in zip(res, cres)]))
    out["search_single"] = {"gpu_us_p50": pct(lat, 50), "gpu_us_p99": pct(lat, 99), "gpu_bulk_copy_us_p50": pct(lat_bulk, 50),
                            "cpu_oracle_us_p50": pct(clat, 50), "cpu_oracle_us_p99": pct(clat, 99), "n": args.n_search,
                            "ids_equal_both_row_copy_modes": bool(same_spec), "ids_equal_oracle_fraction": same_orc,
                            "note": "python ctypes call overhead (~5 us) included on both sides"}
    # ---- small batches through hnsw_index_search_batch (host buffers): latency of one call and the rate it gives
    sweep = {}
    for nb in (1, 8, 32, 148, 1250, 10000):
        if nb > args.n_search:
            break
        qb = np.ascontiguousarray(q[:nb])
        for _ in range(3):
            dev.search_batch(qb, 10, ef=args.ef)
        ts = []
        for _ in range(30):
            t = time.perf_counter()
            dev.search_batch(qb, 10, ef=args.ef)
            ts.append(time.perf_counter() - t)
        sweep[nb] = {"call_us_p50": pct(ts, 50), "qps": nb / float(np.median(ts))}
    out["search_batch_sweep"] = sweep
    if args.only == "search":
        print(json.dumps(out), flush=True)
        return

    # ---- HNSW.NODE.ADD: CPU oracle first (on its own copy of the graph), then the device, same vectors and levels
    t = time.perf_counter()
    for i in range(args.cpu_add):
        orc.add(extra[i], int(lv_extra[i]))
    cpu_add_s = (time.perf_counter() - t) / args.cpu_add
    alat = []
    same = True
    for i in range(args.n_add):
        t = time.perf_counter()
        dev.add(extra[i], int(lv_extra[i]))
        alat.append(time.perf_counter() - t)
        if i == args.cpu_add - 1:   # both sides have now applied the same inserts to the same graph
            same = all(np.array_equal(dev.node_neighbors(n + j, 0), orc.node_neighbors(n + j, 0)) for j in range(args.cpu_add))
    st0 = dev.build_stats()
    t = time.perf_counter()
    dev.add_batch(extra[args.n_add:2 * args.n_add], lv_extra[args.n_add:2 * args.n_add], mode=r.BUILD_EXACT)
    stream_s = (time.perf_counter() - t) / args.n_add
    st1 = dev.build_stats()
    out["node_add_exact"] = {"gpu_single_us_p50": pct(alat, 50), "gpu_single_us_p99": pct(alat, 99),
                             "gpu_stream_us_per_insert": stream_s * 1e6, "gpu_stream_inserts_per_s": 1.0 / stream_s,
                             "cpu_oracle_us_per_insert": cpu_add_s * 1e6, "cpu_oracle_inserts_per_s": 1.0 / cpu_add_s,
                             "dist_evals_per_insert_stream": (st1["dist_evals"] - st0["dist_evals"]) / args.n_add,
                             "n_gpu": args.n_add, "n_cpu": args.cpu_add}
    out["node_add_exact"]["lists_equal_oracle_on_shared_prefix"] = bool(same)

    # ---- HNSW.NODE.DEL
    rng = np.random.default_rng(1)
    victims = rng.permutation(n)[:args.n_del]
    t = time.perf_counter()
    for v in victims[:args.cpu_del]:
        orc.delete(int(v))
    cpu_del_s = (time.perf_counter() - t) / args.cpu_del
    dlat = []
    for v in victims:
        t = time.perf_counter()
        dev.delete(int(v))
        dlat.append(time.perf_counter() - t)
    out["node_del"] = {"gpu_us_p50": pct(dlat, 50), "gpu_us_p99": pct(dlat, 99), "cpu_oracle_us_per_delete": cpu_del_s * 1e6,
                       "n_gpu": args.n_del, "n_cpu": args.cpu_del}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
