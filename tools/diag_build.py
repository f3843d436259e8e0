#!/usr/bin/env python
"""Diagnostic: FAST build in chunks, printing builder counters and degree statistics after each chunk."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import redis_hnsw_b200 as r
from redis_hnsw_b200 import data

n, dim, m, efc, rr = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
chunk = int(sys.argv[6]) if len(sys.argv) > 6 else 20000
x, q = data.lowrank(n, dim, r=rr, seed=123, n_queries=10)
levels = data.draw_levels(n, m, seed=42)
dev = r.DeviceIndex(dim, m, efc)
dev.reserve(n)
pos = 0
while pos < n:
    e = min(n, pos + chunk)
    t0 = time.perf_counter()
    try:
        dev.add_batch(x[pos:e], levels[pos:e], mode=r.BUILD_FAST)
        err = None
    except Exception as ex:
        err = str(ex)
    dt = time.perf_counter() - t0
    g = dev.export_graph()
    deg = np.diff(g["row_offs"].astype(np.int64))
    lv = g["levels"]
    # row index of level-0 rows
    first = np.concatenate([[0], np.cumsum(lv[:-1] + 1)]) if len(lv) else np.zeros(0, int)
    d0 = deg[first[lv >= 0]] if len(lv) else deg
    print("n=%d dt=%.1fs stats=%s deg0 mean=%.1f max=%d over_cap=%.4f overall max=%d err=%s" % (
        e, dt, dev.build_stats(), d0.mean(), d0.max(), (d0 > 2 * m).mean(), deg.max(), err), flush=True)
    if err:
        break
    pos = e
