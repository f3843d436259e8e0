#!/usr/bin/env python
"""The SPEC builder under compute-sanitizer: every policy path of spec_exec_kernel / spec_commit_kernel (dependency- and
row-level validation, checkpoints ahead of the window, suspended executions, operations applied at commit) on small graphs.
    compute-sanitizer --tool memcheck python tools/sanitize_spec.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import redis_hnsw_b200 as r
from redis_hnsw_b200 import data

FAST = os.environ.get("SANITIZE_SPEC_FAST")      # only the suspension path (time budget), two dimensions
STRESS = os.environ.get("SPEC_STRESS")            # no sanitizer: larger graphs, many budgets
CONFIGS = ((128, 16, 64), (32, 5, 40)) if (FAST or STRESS) else ((128, 16, 64), (32, 5, 40), (768, 32, 48))
POLICIES = ({}, {"spec_budget_us": 150}, {"spec_validation": 1}, {"spec_ahead": -1, "spec_window": 24})
if FAST:
    POLICIES = ({}, {"spec_budget_us": 150}, {"spec_budget_us": 60})
if STRESS:
    POLICIES = ({},) + tuple({"spec_budget_us": b} for b in (40, 80, 120, 160, 250, 400)) + ({"spec_budget_us": 100, "spec_window": 96},)
for dim, m, efc in CONFIGS:
    n = (6000 if STRESS else 1300) if dim < 768 else 500
    x, q = data.uniform(n, dim, seed=1, n_queries=16)
    lv = data.draw_levels(n, m, seed=2)
    ref = None
    for opts in POLICIES:
        dev = r.DeviceIndex(dim, m, efc)
        for k, v in opts.items():
            dev.set_option(k, v)
        dev.add_batch(x, lv, mode=r.BUILD_SPEC)
        g = dev.export_graph()
        st = dev.build_stats()
        if ref is None:
            ref = g
        else:
            assert np.array_equal(g["nbrs"], ref["nbrs"]) and np.array_equal(g["row_offs"], ref["row_offs"]), opts
        dev.close()
        print("ok", dim, m, efc, opts, st["spec_rounds"], st["spec_executions"], st["spec_rows_as_operations"], flush=True)
