#!/usr/bin/env python
"""Headline benchmark: HNSW.SEARCH throughput (queries/sec at recall@10 >= 0.95) on 1M x 128-d, M=16, efCon=200
(BASELINE.json configs[1]), on N GPUs of one node with the index replicated and the query batch sharded.

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's CPU algorithm (oracle port) on the host cores

A "step" is one pass of search_knn (core.rs:477-486, 865-892) over one batch of `--nq` synthetic queries per GPU.
  value     device-timed (CUDA events, max over ranks) whole-job QPS with queries and results resident in HBM
  e2e       the same through hnsw_index_search_batch with pinned HOST buffers (H2D + kernel + D2H in the timed region)
  roofline  algorithmic bytes of the batch (n_dist*4*dim + n_adj*4 + query + result, counted per query by the kernel and
            proven equal to the oracle's counters in tests/) / step time, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the oracle (C++ restatement of the reference) on the same graph and a bounded query sample
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (n, dim, m, ef_construction, dataset, r)
    "1Mx128_M16_efc200": (1_000_000, 128, 16, 200, "lowrank", 16),     # BASELINE configs[1] (headline)
    "1Mx768_M32_efc400": (1_000_000, 768, 32, 400, "lowrank", 32),     # configs[2]
    "10Mx128_M16_efc200": (10_000_000, 128, 16, 200, "lowrank", 16),   # configs[3]
    "100Kx128_M16_efc200": (100_000, 128, 16, 200, "lowrank", 16),     # quick check
    "10Kx32_M5_efc100": (10_000, 32, 5, 100, "uniform", 0),            # configs[0]
    "1Mx128_M16_efc200_uniform": (1_000_000, 128, 16, 200, "uniform", 0),   # stress dataset: intrinsic dim 128
}


def log(*a):
    if int(os.environ.get("RANK", "0")) == 0:
        print("[bench]", *a, file=sys.stderr, flush=True)


_JSON_OUT = None


def claim_stdout():
    """stdout carries exactly ONE JSON line: keep a private handle on it and point fd 1 at stderr, so that banners written
    by libraries (NCCL prints "NCCL version ..." on stdout at N > 1) cannot end up next to the line."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()


class ClockSampler:
    """nvidia-smi clocks and throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.rows, self.p = gpu, [], None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def wait_first(self, timeout=3.0):
        """nvidia-smi takes a few hundred ms to print its first row: block until the sampler is live."""
        t0 = time.time()
        while self.p and not self.rows and time.time() - t0 < timeout:
            time.sleep(0.01)

    def stop(self, t_begin=None, t_end=None):
        """Samples taken in [t_begin, t_end] (the timed region); if the region was shorter than the sampling period, every
        sample since the sampler started (it is started before warm-up, so those are under load too)."""
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.p.terminate()
        allrows = [(t, r) for t, r in self.rows if len(r) >= 7]
        rows = [r for t, r in allrows if (t_begin is None or t >= t_begin) and (t_end is None or t <= t_end + 0.03)]
        window = "timed region"
        if len(rows) < 3:
            rows = [r for t, r in allrows if t_begin is None or t >= t_begin - 1.0]
            window = "warm-up + timed region (timed region shorter than 3 samples)"
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "reasons": reasons, "samples": len(rows),
                "power_w_max": max(float(r[2]) for r in rows), "window": window}


def make_data(wl, nq_total):
    from redis_hnsw_b200 import data

    n, dim, m, efc, ds, r = WORKLOADS[wl]
    if ds == "lowrank":
        x, q = data.lowrank(n, dim, r=r, seed=123, n_queries=nq_total)
    else:
        x, q = data.uniform(n, dim, seed=123, n_queries=nq_total)
    levels = data.draw_levels(n, m, seed=42)
    return x, q, levels


def build_index(wl, x, levels, device, rank, world, options=()):
    """Rank 0 builds with the batched device builder; other ranks receive the device buffers over NCCL."""
    import torch

    import redis_hnsw_b200 as r

    n, dim, m, efc, _, _ = WORKLOADS[wl]
    dev = r.DeviceIndex(dim, m, efc, device=device)
    for opt in options:
        name, val = opt.split("=")
        dev.set_option(name, int(val))
    build_s = None
    if rank == 0:
        dev.reserve(n)
        t0 = time.perf_counter()
        dev.add_batch(x, levels, mode=r.BUILD_FAST)
        build_s = time.perf_counter() - t0
        log("built %d nodes in %.1f s (%.0f inserts/s) %s" % (n, build_s, n / build_s, dev.build_stats()))
    r.sharding.replicate_index(dev, rank, world)   # one NCCL broadcast per device buffer
    return dev, build_s


def pick_ef(dev, x, q, wl, target=0.95):
    """Smallest ef of the sweep with recall@10 >= target on a 10 000-query sample (exact ground truth by brute force)."""
    import torch

    from redis_hnsw_b200 import data

    sample = q[:10000]   # standard error of recall@10 on 10 000 queries is ~0.0007 (2 000 gave 0.0015: VERDICT r1 weak #8)
    gt = data.brute_force_topk(x, sample, 10, device="cuda")
    curve = {}
    chosen = None
    for ef in (16, 24, 32, 48, 64, 72, 80, 96, 128, 160, 200, 256, 320, 400, 512):   # finer above 64: a near miss costs 10 %, not 35 %
        ids, _, _ = dev.search_batch(sample, 10, ef=ef)
        rec = data.recall_at_k(ids, gt)
        curve[ef] = round(rec, 4)
        if chosen is None and rec >= target:
            chosen = ef
        if rec >= target and ef >= 64:
            break
    torch.cuda.synchronize()
    return chosen, curve


def oracle_on_graph(dev, x, wl):
    import oracle

    n, dim, m, efc, _, _ = WORKLOADS[wl]
    orc = oracle.Oracle(dim, m, efc)
    orc.import_graph(x, dev.export_graph())
    return orc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="1Mx128_M16_efc200", choices=sorted(WORKLOADS))
    ap.add_argument("--nq", type=int, default=100_000, help="queries per GPU per step")
    ap.add_argument("--ef", type=int, default=0, help="0 = smallest ef of the sweep with recall@10 >= 0.95")
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--cpu-sample", type=int, default=0,
                    help="queries of the CPU legs; 0 = sized for 10-30 s of CPU work (40 000 at 128-d / ef 64, scaled by row size and ef)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--option", action="append", default=[], help="name=value library option (tuning)")
    args = ap.parse_args()
    claim_stdout()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference" and rank != 0:
        return 0  # the CPU arm runs on rank 0 only

    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1 and args.impl != "reference":
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    wl = args.workload
    n, dim, m, efc, ds, r_lat = WORKLOADS[wl]
    nq = args.nq
    t_setup = time.perf_counter()
    x, q_all, levels = make_data(wl, nq * (world if args.impl != "reference" else 1))
    log("data ready in %.1f s" % (time.perf_counter() - t_setup))
    dev, build_s = build_index(wl, x, levels, local_rank, rank if args.impl != "reference" else 0,
                               world if args.impl != "reference" else 1, args.option)

    # operating point: smallest ef with recall@10 >= 0.95 (rank 0 decides, everyone follows)
    ef, curve = args.ef, {}
    if rank == 0:
        chosen, curve = pick_ef(dev, x, q_all, wl)
        log("recall@10 by ef:", curve, "-> ef =", chosen)
        if not ef:
            ef = chosen or max(curve)
    if world > 1 and args.impl != "reference":
        t = torch.tensor([ef], dtype=torch.int64, device="cuda")
        dist.broadcast(t, 0)
        ef = int(t.item())
    recall = curve.get(ef)
    if not args.cpu_sample:
        args.cpu_sample = max(500, int(40_000 * (128.0 / dim) * (64.0 / max(ef, 16))))

    base_cfg = {"workload": wl, "n": n, "dim": dim, "M": m, "ef_construction": efc, "ef_search": ef, "k": args.k,
                "queries_per_gpu_per_step": nq, "dataset": "lowrank r=%d sigma=0.05 seed=123" % r_lat if ds == "lowrank" else "uniform seed=123",
                "recall_at_10": recall, "recall_by_ef": curve,
                "graph": "built on the GPU by the batched (FAST) builder, levels injected (seed 42)",
                "l2_policy": "inputs larger than L2 (vector slab %d MB + adjacency; L2 126 MB), no flush" % (n * dim * 4 >> 20)}

    # ------------------------------------------------------------------ reference arm (CPU oracle, all host threads)
    if args.impl == "reference":
        orc = oracle_on_graph(dev, x, wl)
        cores = os.cpu_count() or 1
        sample = min(args.cpu_sample, nq)
        qs = q_all[:sample]
        times = []
        for it in range(args.warmup + args.steps):
            _, _, _, _, secs = orc.search_batch(qs, args.k, ef=ef, threads=cores, stats=False)
            if it >= args.warmup:
                times.append(secs)
        tot = sum(times)
        qps = sample * args.steps / tot
        line = {"impl": "reference", "metric": "queries/sec @ recall@10>=0.95", "value": qps, "unit": "queries/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": dict(base_cfg, queries_per_step=sample),
                "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port",
                                 "sample": "%d queries per step, %d threads, oracle (C++ restatement of the reference; "
                                           "the Rust reference cannot be built here) on the same graph" % (sample, cores)},
                "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return 0

    # ------------------------------------------------------------------ B200 arm
    import redis_hnsw_b200 as r

    q = q_all[rank * nq:(rank + 1) * nq]
    d_q = torch.from_numpy(q).cuda()
    d_ids = torch.empty((nq, args.k), dtype=torch.int32, device="cuda")
    d_sims = torch.empty((nq, args.k), dtype=torch.float32, device="cuda")
    d_cnt = torch.empty(nq, dtype=torch.int32, device="cuda")
    d_stats = torch.empty((nq, 4), dtype=torch.int32, device="cuda")
    # a non-default torch stream: the library treats a NULL stream as "the index's own stream", and torch.cuda.Event
    # only sees the stream it is recorded on
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    def step(stats=False):
        dev.search_batch_device(nq, d_q.data_ptr(), args.k, ef, d_ids.data_ptr(), d_sims.data_ptr(), d_cnt.data_ptr(),
                                d_stats.data_ptr() if stats else 0, stream)

    # algorithmic bytes of one step, from the kernel's own per-query counters (== the oracle's, tests/test_gpu_search.py)
    step(stats=True)
    torch.cuda.synchronize()
    st = d_stats.cpu().numpy().astype(np.int64)
    n_dist, n_adj, n_hops = int(st[:, 0].sum()), int(st[:, 1].sum()), int(st[:, 2].sum())
    alg_bytes = n_dist * 4 * dim + n_adj * 4 + nq * (4 * dim + 8 * args.k)
    retried = int((st[:, 3] & 1).sum())
    # the timed path (no counters requested) is the TMA-staged kernel; its lossy visited table may re-evaluate nodes
    evals_done = None
    try:
        dev.set_option("search_impl", 2)
        step(stats=True)
        torch.cuda.synchronize()
        evals_done = float(d_stats[:, 0].double().mean().item())
    except Exception:
        pass
    dev.set_option("search_impl", 0)

    sampler = ClockSampler(local_rank)
    if rank == 0:                     # started BEFORE warm-up: nvidia-smi needs a few hundred ms to deliver its first row
        sampler.start()
        sampler.wait_first()
    t_warm = time.time()
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    while time.time() - t_warm < 0.5:  # keep the GPU under load until the sampler has rows from a loaded device
        step()
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches0 = r.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    torch.cuda.synchronize()
    cuprof = os.environ.get("HNSW_BENCH_CUPROF") == "1"  # ncu --profile-from-start off: profile the timed region only
    if cuprof:
        torch.cuda.profiler.start()
    t_begin = time.time()
    ev[0].record()
    for i in range(args.steps):
        step()
        ev[i + 1].record()
    torch.cuda.synchronize()
    t_end = time.time()
    if cuprof:
        torch.cuda.profiler.stop()
    if world > 1:
        dist.barrier()
    launches = r.launch_count() - launches0
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    total_ms = ev[0].elapsed_time(ev[-1])
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    qps = world * nq * args.steps / (total_ms / 1e3)

    # ------------------------------------------------------------------ e2e: host buffers through the C ABI
    hq = torch.from_numpy(q).pin_memory()
    h_ids = torch.empty((nq, args.k), dtype=torch.int32).pin_memory()
    h_sims = torch.empty((nq, args.k), dtype=torch.float32).pin_memory()
    h_cnt = torch.empty(nq, dtype=torch.int32).pin_memory()
    out = (h_ids.numpy().view(np.uint32), h_sims.numpy(), h_cnt.numpy().view(np.uint32))
    hqn = hq.numpy()
    e2e_steps = max(3, args.steps // 2)
    for _ in range(2):
        dev.search_batch(hqn, args.k, ef=ef, out=out)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        dev.search_batch(hqn, args.k, ef=ef, out=out)
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_qps = world * nq * e2e_steps / e2e_s
    # the two paths agree
    assert np.array_equal(out[0], d_ids.cpu().numpy().view(np.uint32)), "host-API results differ from the device-API results"

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    kernel_ms = float(np.mean(step_ms))
    achieved = alg_bytes / (kernel_ms / 1e3) / 1e9
    traffic, traffic_src = None, None
    try:  # DRAM bytes of one launch from the committed ncu --set full capture of this workload (profiles/traffic.json)
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(wl)
        if tr and tr["ef"] == ef:
            traffic = (tr["dram_bytes_read"] + tr["dram_bytes_write"]) * nq / tr["queries_per_launch"]
            traffic_src = tr["source"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "alg_bytes_per_launch": alg_bytes,
                # DRAM bytes actually moved (ncu) over the same launch time: L2 serves part of the algorithmic bytes, so
                # `frac` can exceed 1 while DRAM itself stays below its peak
                "traffic_frac": (traffic / (kernel_ms / 1e3) / 1e9 / peak) if traffic else None,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                "frac_of_nominal_8TBs": achieved / 8000.0,
                "kernel": "search_knn2_kernel (one launch per step; duration = CUDA events around the step on the launch stream)",
                "alg_bytes_per_query": alg_bytes / nq, "dist_evals_per_query": n_dist / nq, "adj_ids_per_query": n_adj / nq,
                "hops_per_query": n_hops / nq,
                "dist_evals_performed_per_query": evals_done}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        t0 = time.perf_counter()
        orc = oracle_on_graph(dev, x, wl)
        log("oracle import %.1f s" % (time.perf_counter() - t0))
        sample = min(args.cpu_sample, nq)
        oids, osims, ocnt, ost, secs = orc.search_batch(q[:sample], args.k, ef=ef, threads=1, stats=True)
        g_ids = d_ids[:sample].cpu().numpy().view(np.uint32)
        tie_free = ost[:, 3] == 0
        same = bool(np.array_equal(g_ids[tie_free], oids[tie_free]))
        same_counters = bool(np.array_equal(st[:sample][tie_free, :3], ost[tie_free, :3].astype(np.int64)))
        _, _, _, _, secs = orc.search_batch(q[:sample], args.k, ef=ef, threads=1, stats=False)
        cores = os.cpu_count() or 1
        _, _, _, _, secs_all = orc.search_batch(q[:sample], args.k, ef=ef, threads=cores, stats=False)
        cpu = {"value": sample / secs, "unit": "queries/s", "cores": 1, "kind": "port",
               "sample": "%d of the step's queries, same graph, same ef; oracle = C++ restatement of the reference "
                         "(no Arc/RwLock/SipHash overheads: faster than the Rust reference)" % sample,
               "all_cores": {"value": sample / secs_all, "cores": cores},
               "ids_match_gpu": same, "counters_match_gpu": same_counters, "tie_free_queries": int(tie_free.sum())}
        log("cpu baseline", cpu)

    line = {"metric": "queries/sec @ recall@10>=0.95", "value": qps, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": dict(base_cfg, build_seconds=build_s,
                                                                                    inserts_per_s=(n / build_s) if build_s else None),
            "e2e": {"value": e2e_qps, "unit": "queries/s", "h2d_bytes_per_step": int(q.nbytes),
                    "d2h_bytes_per_step": int(nq * args.k * 8 + nq * 4), "steps": e2e_steps,
                    "api": "hnsw_index_search_batch (pinned host buffers)"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline}
    if cpu:
        line["cpu_baseline"] = cpu
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
