#!/usr/bin/env python
"""Small end-to-end run of every kernel family (metric, both search kernels, search_level, EXACT and FAST insert, delete)
meant to be run under compute-sanitizer:  compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import redis_hnsw_b200 as r
from redis_hnsw_b200 import data

for dim, m, efc in ((128, 16, 64), (32, 5, 40), (20, 6, 32), (96, 8, 32)):
    n = 1500
    x, q = data.uniform(n, dim, seed=1, n_queries=64)
    lv = data.draw_levels(n, m, seed=2)
    a = r.l2_batch(x[:256], x[256:512])
    dev = r.DeviceIndex(dim, m, efc)
    dev.add_batch(x[:400], lv[:400], mode=r.BUILD_EXACT)
    dev.add_batch(x[400:600], lv[400:600], mode=r.BUILD_SPEC)      # r2: speculative-exact windows (spec_exec / spec_commit)
    dev.add_batch(x[600:], lv[600:], mode=r.BUILD_FAST)
    for i in range(5):
        dev.add(q[i], -1)
    ids, sims, cnt, st = dev.search_batch(q, 10, ef=48, stats=True)
    ids2, sims2, cnt2 = dev.search_batch(q, 10, ef=48)
    assert np.array_equal(ids, ids2)
    if dim == 128:   # host batches >= 8192 queries take the two-stream pipelined path
        big = np.tile(q, (128, 1))
        ib, sb, cb = dev.search_batch(big, 10, ef=48)
        assert np.array_equal(ib[:64], ids) and np.array_equal(ib[-64:], ids)
    dev.search(q[0], 5)
    if dim in (128, 32):                                            # r2: one query per CTA, lookahead kernel (options)
        # HNSW_SMOKE_CTA=1 adds the CTA-per-query kernel: synccheck reports its named-barrier handshake (owner and worker
        # warps meet at bar.sync 1 / 2 from different program points) as divergent, so it is kept out of the default run
        for opt in (("search_cta", "lookahead") if os.environ.get("HNSW_SMOKE_CTA") else ("lookahead",)):
            dev.set_option(opt, 1)
            i3, s3, c3 = dev.search_batch(q[:8], 10, ef=48)
            assert np.array_equal(i3, ids[:8])
            dev.set_option(opt, 0)
    dev.search_level(q[0], int(dev.params()["enterpoint"]), 8, 0)
    for v in (3, 700, int(dev.params()["enterpoint"]), 1499):
        dev.delete(v)
    g = dev.export_graph()
    dev.close()
    print("ok", dim, m, efc, int(cnt.sum()), flush=True)
