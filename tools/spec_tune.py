#!/usr/bin/env python
"""SPEC window policy at full scale without paying for a full exact build: the workload's graph is built by the FAST
builder up to `--base` nodes, then the NODE.ADD stream continues in SPEC mode, one piece per option setting.

    python tools/spec_tune.py --workload 1Mx128_M16_efc200 --base 940000 --piece 10000 --grid "spec_mult=15,20,30,40,60,80"
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
    ap.add_argument("--base", type=int, default=940000)
    ap.add_argument("--piece", type=int, default=10000)
    ap.add_argument("--grid", default="spec_mult=15,20,30,40,60,80")
    args = ap.parse_args()
    import redis_hnsw_b200 as r

    n, dim, m, efc, _, _ = bench.WORKLOADS[args.workload]
    x, _, levels = bench.make_data(args.workload, 0)
    dev = r.DeviceIndex(dim, m, efc)
    dev.reserve(n)
    dev.add_batch(x[:args.base], levels[:args.base], mode=r.BUILD_FAST)
    name, vals = args.grid.split("=")
    done = args.base
    prev = dev.build_stats()
    for v in vals.split(","):
        if done + args.piece > n:
            break
        dev.set_option(name, int(v))
        t0 = time.perf_counter()
        dev.add_batch(x[done:done + args.piece], levels[done:done + args.piece], mode=r.BUILD_SPEC)
        dt = time.perf_counter() - t0
        done += args.piece
        st = dev.build_stats()
        d = {k: st[k] - prev[k] for k in st if k != "spec_max_window"}
        prev = st
        rounds = max(1, d["spec_rounds"])
        print(json.dumps({name: int(v), "nodes": done, "inserts_per_s": round(args.piece / dt, 1), "inserts_per_round": round(args.piece / rounds, 2),
                          "ms_per_round": round(1e3 * dt / rounds, 3), "executions_per_insert": round(d["spec_executions"] / args.piece, 3)}), flush=True)


if __name__ == "__main__":
    main()
