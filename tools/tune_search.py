#!/usr/bin/env python
"""Tuning sweep of the search kernels on one GPU: builds the workload's index once, then times search_batch_device
for every combination of the library options given on the command line.

    python tools/tune_search.py --workload 1Mx128_M16_efc200 --ef 64 --grid "search_impl=2;stage_rows=8,16,32;recent_slots=512,1024,2048;search_block=64,128,256"
"""
import argparse
import itertools
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
    ap.add_argument("--ef", type=int, default=64)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--grid", default="search_impl=1,2")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import torch

    wl = args.workload
    n, dim, m, efc, _, _ = bench.WORKLOADS[wl]
    x, q, levels = bench.make_data(wl, args.nq)
    dev, _binfo = bench.build_index(wl, x, levels, 0, 0, 1)
    build_s = _binfo["build_seconds"]
    nq, k = args.nq, 10
    d_q = torch.from_numpy(q).cuda()
    d_ids = torch.empty((nq, k), dtype=torch.int32, device="cuda")
    d_sims = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    d_cnt = torch.empty(nq, dtype=torch.int32, device="cuda")
    d_stats = torch.empty((nq, 4), dtype=torch.int32, device="cuda")
    ts = torch.cuda.Stream()
    torch.cuda.set_stream(ts)

    def run(stats=False):
        dev.search_batch_device(nq, d_q.data_ptr(), k, args.ef, d_ids.data_ptr(), d_sims.data_ptr(), d_cnt.data_ptr(),
                                d_stats.data_ptr() if stats else 0, ts.cuda_stream)

    dev.set_option("search_impl", 1)
    run(stats=True)
    torch.cuda.synchronize()
    st = d_stats.cpu().numpy().astype(np.int64)
    ref_ids = d_ids.cpu().numpy().copy()
    alg = int(st[:, 0].sum()) * 4 * dim + int(st[:, 1].sum()) * 4 + nq * (4 * dim + 8 * k)
    print("alg bytes/query %.0f  evals/query %.1f" % (alg / nq, st[:, 0].mean()), flush=True)
    axes = []
    for part in args.grid.split(";"):
        name, vals = part.split("=")
        axes.append([(name, int(v)) for v in vals.split(",")])
    rows = []
    for combo in itertools.product(*axes):
        try:
            for name, v in combo:
                dev.set_option(name, v)
            impl = dict(combo).get("search_impl", 0)
            run(stats=(impl == 2))
            torch.cuda.synchronize()
            evals = float(d_stats[:, 0].float().mean().item()) if impl == 2 else float(st[:, 0].mean())
            same = bool(np.array_equal(d_ids.cpu().numpy(), ref_ids))
            for _ in range(2):
                run()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(args.steps):
                run()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            row = dict(combo)
            row.update(ms=round(ms, 3), mqps=round(nq / ms / 1e3, 3), gbs=round(alg / ms / 1e6, 1), evals=round(evals, 1), same_ids=same)
        except Exception as ex:  # an option combination the kernel cannot run (shared memory, ...)
            row = dict(combo)
            row.update(error=str(ex)[:80])
        rows.append(row)
        print(json.dumps(row), flush=True)
    if args.out:
        json.dump(rows, open(args.out, "w"), indent=0)


if __name__ == "__main__":
    main()
