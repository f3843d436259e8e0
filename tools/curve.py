#!/usr/bin/env python
"""recall@10 - ef - QPS curve of HNSW.SEARCH on one GPU (always publish the curve next to the operating
point).  Device-resident queries, CUDA-event timing, same workloads as bench.py.

    python tools/curve.py [--workload 1Mx128_M16_efc200] > gpurun_out/curve.json
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="1Mx128_M16_efc200")
    ap.add_argument("--nq", type=int, default=100_000)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--efs", default="8,16,24,32,48,64,96,128,200,256,400,512")
    args = ap.parse_args()
    import torch

    from redis_hnsw_b200 import data

    wl = args.workload
    n, dim, m, efc, ds, r_lat = bench.WORKLOADS[wl]
    x, q, levels = bench.make_data(wl, args.nq)
    dev, _binfo = bench.build_index(wl, x, levels, 0, 0, 1)
    build_s = _binfo["build_seconds"]
    gt = data.brute_force_topk(x, q[:2000], 10, device="cuda")
    d_q = torch.from_numpy(q).cuda()
    nq = args.nq
    d_ids = torch.empty((nq, 10), dtype=torch.int32, device="cuda")
    d_sims = torch.empty((nq, 10), dtype=torch.float32, device="cuda")
    d_cnt = torch.empty(nq, dtype=torch.int32, device="cuda")
    d_st = torch.empty((nq, 4), dtype=torch.int32, device="cuda")
    ts = torch.cuda.Stream()
    torch.cuda.set_stream(ts)
    rows = []
    for ef in (int(e) for e in args.efs.split(",")):
        def step(stats=False):
            dev.search_batch_device(nq, d_q.data_ptr(), 10, ef, d_ids.data_ptr(), d_sims.data_ptr(), d_cnt.data_ptr(),
                                    d_st.data_ptr() if stats else 0, ts.cuda_stream)
        step(stats=True)                       # exact-visited kernel: the reference's own work counters
        torch.cuda.synchronize()
        st = d_st.cpu().numpy().astype(np.int64)
        alg = int(st[:, 0].sum()) * 4 * dim + int(st[:, 1].sum()) * 4 + nq * (4 * dim + 80)
        for _ in range(2):
            step()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        rec = data.recall_at_k(d_ids[:2000].cpu().numpy().view(np.uint32), gt)
        rows.append({"ef": ef, "recall_at_10": round(rec, 4), "qps": nq / (ms / 1e3), "ms_per_batch": ms,
                     "dist_evals_per_query": float(st[:, 0].mean()), "hops_per_query": float(st[:, 2].mean()),
                     "alg_GBps": alg / (ms / 1e3) / 1e9})
        print("[curve]", rows[-1], file=sys.stderr, flush=True)
    print(json.dumps({"workload": wl, "n": n, "dim": dim, "M": m, "ef_construction": efc, "k": 10, "queries_per_batch": nq,
                      "build_seconds": build_s, "curve": rows}))


if __name__ == "__main__":
    main()
