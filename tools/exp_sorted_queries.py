#!/usr/bin/env python
"""Experiment: does processing spatially sorted queries raise the L2 hit rate enough to matter?  Sort the query batch on
the host by the nearest of 4096 sampled data points (several tie-break variants), time the device-resident search."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import torch
from redis_hnsw_b200 import data

wl = "1Mx128_M16_efc200"
n, dim, m, efc, ds, r_lat = bench.WORKLOADS[wl]
nq = 100_000
x, q, levels = bench.make_data(wl, nq)
dev, _ = bench.build_index(wl, x, levels, 0, 0, 1)
xt = torch.from_numpy(x[np.random.default_rng(0).choice(n, 4096, replace=False)]).cuda()
qt = torch.from_numpy(q).cuda()
d = (qt * qt).sum(1, keepdim=True) - 2 * qt @ xt.T + (xt * xt).sum(1)[None, :]
lab = d.argmin(1).cpu().numpy()
# order labels themselves along a 1-D projection so that neighbouring labels are spatial neighbours
proj = (xt @ torch.randn(dim, 1, device="cuda")).squeeze(1).cpu().numpy()
orders = {"original": np.arange(nq), "by_label": np.argsort(lab, kind="stable"), "by_label_projected": np.argsort(proj[lab], kind="stable"),
          "random": np.random.default_rng(1).permutation(nq)}
d_ids = torch.empty((nq, 10), dtype=torch.int32, device="cuda")
d_sims = torch.empty((nq, 10), dtype=torch.float32, device="cuda")
d_cnt = torch.empty(nq, dtype=torch.int32, device="cuda")
ts = torch.cuda.Stream()
torch.cuda.set_stream(ts)
for name, o in orders.items():
    dq = torch.from_numpy(np.ascontiguousarray(q[o])).cuda()
    def step():
        dev.search_batch_device(nq, dq.data_ptr(), 10, 64, d_ids.data_ptr(), d_sims.data_ptr(), d_cnt.data_ptr(), 0, ts.cuda_stream)
    for _ in range(3):
        step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("%-20s %.3f ms/batch  %.2f M QPS" % (name, ms, nq / ms / 1e3), flush=True)
