#!/usr/bin/env python
"""Diagnostic: wall time of FAST builds in one process (first build pays one-time costs), per batch-size option."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import redis_hnsw_b200 as r
from redis_hnsw_b200 import data

n, dim, m, efc, rr = 1000000, 128, 16, 200, 16
x, q = data.lowrank(n, dim, r=rr, seed=123, n_queries=10)
levels = data.draw_levels(n, m, seed=42)
for b in [4096, 4096, 8192, 16384, 4096]:
    dev = r.DeviceIndex(dim, m, efc)
    dev.set_option("build_batch", b)
    dev.reserve(n)
    t0 = time.perf_counter()
    marks = []
    for s in range(0, n, 250000):
        dev.add_batch(x[s:s + 250000], levels[s:s + 250000], mode=r.BUILD_FAST)
        marks.append(round(time.perf_counter() - t0, 2))
    print("batch", b, "cumulative s per 250K:", marks, dev.build_stats(), flush=True)
    dev.close()
