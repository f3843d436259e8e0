#!/usr/bin/env python
"""Throughput of the SPEC (speculative-exact) builder as the graph grows: the NODE.ADD stream of a workload is fed in
pieces and every piece is timed.  One JSON line per piece: nodes so far, inserts/s of the piece, commit rounds, inserts
per round, executions per insert, distance evaluations per insert.

    python tools/spec_probe.py --workload 1Mx128_M16_efc200 --limit 300000 --piece 20000 [--option spec_window=64]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="1Mx128_M16_efc200")
    ap.add_argument("--limit", type=int, default=0)
    ap.add_argument("--piece", type=int, default=20000)
    ap.add_argument("--seconds", type=float, default=1e9, help="stop feeding pieces after this much build time")
    ap.add_argument("--option", action="append", default=[])
    ap.add_argument("--mode", default="spec", choices=["spec", "exact", "fast"])
    args = ap.parse_args()
    import redis_hnsw_b200 as r

    n, dim, m, efc, _, _ = bench.WORKLOADS[args.workload]
    x, _, levels = bench.make_data(args.workload, 0)
    if args.limit:
        n = min(n, args.limit)
    dev = r.DeviceIndex(dim, m, efc)
    dev.reserve(n)
    for opt in args.option:
        k, v = opt.split("=")
        dev.set_option(k, int(v))
    mode = {"spec": r.BUILD_SPEC, "exact": r.BUILD_EXACT, "fast": r.BUILD_FAST}[args.mode]
    done, total_s, prev = 0, 0.0, dev.build_stats()
    while done < n and total_s < args.seconds:
        k = min(args.piece, n - done)
        t0 = time.perf_counter()
        dev.add_batch(x[done:done + k], levels[done:done + k], mode=mode)
        dt = time.perf_counter() - t0
        total_s += dt
        done += k
        st = dev.build_stats()
        d = {key: st[key] - prev[key] for key in st if key != "spec_max_window"}
        prev = st
        rounds = max(1, d["spec_rounds"])
        print(json.dumps({"nodes": done, "piece_s": round(dt, 3), "inserts_per_s": round(k / dt, 1),
                          "rounds": d["spec_rounds"], "inserts_per_round": round(d["inserts"] / rounds, 2),
                          "ms_per_round": round(1e3 * dt / rounds, 3),
                          "executions_per_insert": round(d["spec_executions"] / max(1, d["inserts"]), 3),
                          "dist_evals_per_insert": round(d["dist_evals"] / max(1, d["inserts"]), 1),
                          "wasted_evals_per_insert": round(d["spec_dist_evals_wasted"] / max(1, d["inserts"]), 1),
                          "exact_fallbacks": d["spec_exact_fallbacks"], "max_window": st["spec_max_window"]}), flush=True)
    print(json.dumps({"total_nodes": done, "total_s": round(total_s, 2), "inserts_per_s": round(done / total_s, 1),
                      "mode": args.mode, "options": args.option}), flush=True)


if __name__ == "__main__":
    main()
