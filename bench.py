#!/usr/bin/env python
"""Benchmarks of the HNSW hot path on N GPUs of one node (index replicated, queries sharded).

    python bench.py --gpus 1 --steps 20 --warmup 3                      # headline: BASELINE configs[1], weak scaling
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...                                 # the reference's CPU algorithm (oracle port)
    torchrun ... bench.py --gpus 8 --scaling strong --workload 10Mx128_M16_efc200 --nq 10000   # BASELINE configs[3], literal
    python bench.py --bench build --workload 1Mx128_M16_efc200           # BASELINE configs[4]: NODE.ADD stream throughput

Search bench.  A "step" is one pass of search_knn (core.rs:477-486, 865-892) over one batch of synthetic queries:
  weak    `--nq` queries per GPU per step (the batch grows with N)
  strong  ONE batch of `--nq` queries per step, sliced over the GPUs, results all-gathered inside the timed region
  value     device-timed (CUDA events, max over ranks) whole-job QPS with queries and results resident in HBM
  e2e       the same through hnsw_index_search_batch with pinned HOST buffers (H2D + kernel + D2H in the timed region)
  roofline  algorithmic bytes of rank 0's launch (n_dist*4*dim + n_adj*4 + query + result, counted per query by the kernel
            and proven equal to the oracle's counters in tests/) / its duration, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the oracle (C++ restatement of the reference) on the same graph and a bounded query sample
Build bench.  A step is one piece of the NODE.ADD stream (core.rs:383-412, 489-599) through hnsw_index_add_batch; the
  first `--warmup` pieces are the untimed start of the stream; value = inserts/s over the timed pieces.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
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
LITERAL_EF = {"1Mx128_M16_efc200": 64, "1Mx768_M32_efc400": 128}       # the efSearch BASELINE.json quotes


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
        """Samples taken in [t_begin, t_end] (the timed region); if the region was shorter than three sampling periods,
        every sample since one second before it (the sampler starts before warm-up, so those are under load too)."""
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


GRAPH_LABEL = {
    "fast": "built on the GPU by the batched (FAST) builder — a labelled extension: nodes of one batch do not see each other, "
            "so the graph differs from the reference's; levels injected (seed 42)",
    "spec": "built on the GPU by the speculative-exact (SPEC) builder: list for list the graph the reference's sequential "
            "NODE.ADD stream builds (tests/test_gpu_spec_build.py); levels injected (seed 42)",
    "exact": "built on the GPU by the one-warp EXACT stream: the reference's sequential graph; levels injected (seed 42)",
}


def build_index(wl, x, levels, device, rank, world, options=(), graph="fast"):
    """Rank 0 builds on its GPU; the other ranks receive the device buffers over NCCL (one broadcast per buffer)."""
    import redis_hnsw_b200 as r

    n, dim, m, efc, _, _ = WORKLOADS[wl]
    dev = r.DeviceIndex(dim, m, efc, device=device)
    for opt in options:
        name, val = opt.split("=")
        dev.set_option(name, int(val))
    info = {"build_seconds": None, "build_stats": None}
    if rank == 0:
        dev.reserve(n)
        mode = {"spec": r.BUILD_SPEC, "fast": r.BUILD_FAST, "exact": r.BUILD_EXACT}[graph]
        t0 = time.perf_counter()
        dev.add_batch(x, levels, mode=mode)
        info["build_seconds"] = time.perf_counter() - t0
        info["build_stats"] = dev.build_stats()
        log("built %d nodes (%s) in %.1f s (%.0f inserts/s) %s" % (n, graph, info["build_seconds"], n / info["build_seconds"],
                                                                  info["build_stats"]))
    nbytes, secs = r.sharding.replicate_index(dev, rank, world)
    if world > 1:
        info["replicate"] = {"bytes": int(nbytes), "seconds": secs, "gbs_per_receiver": nbytes / secs / 1e9 if secs else None,
                             "how": "one ncclBroadcast per device buffer (9 buffers), wall clock around the broadcasts on rank 0"}
        log("replicated %.2f GB to %d ranks in %.3f s (%.1f GB/s per receiver)" % (nbytes / 1e9, world - 1, secs, nbytes / secs / 1e9))
    return dev, info


def pick_ef(dev, x, q, wl, target=0.95, force=()):
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
    for ef in force:
        if ef and ef not in curve:
            ids, _, _ = dev.search_batch(sample, 10, ef=ef)
            curve[ef] = round(data.recall_at_k(ids, gt), 4)
    torch.cuda.synchronize()
    return chosen, curve


def oracle_on_graph(dev, x, wl):
    import oracle

    n, dim, m, efc, _, _ = WORKLOADS[wl]
    orc = oracle.Oracle(dim, m, efc)
    orc.import_graph(x, dev.export_graph())
    return orc


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def dataset_label(ds, r_lat):
    return "lowrank r=%d sigma=0.05 seed=123" % r_lat if ds == "lowrank" else "uniform seed=123"


# ====================================================================================================== build bench

def bench_build(args, rank, world, local_rank):
    """BASELINE configs[4]: throughput of the NODE.ADD stream.  "replicas only" (DESIGN.md §5): the sequential stream does
    not shard, so with N GPUs every rank would build the same graph; the bench runs on rank 0's GPU and says so."""
    import torch

    import oracle
    import redis_hnsw_b200 as r
    from redis_hnsw_b200 import data

    if rank != 0:
        return 0
    wl = args.workload
    n, dim, m, efc, ds, r_lat = WORKLOADS[wl]
    x, _, levels = make_data(wl, 0)
    mode = {"spec": r.BUILD_SPEC, "fast": r.BUILD_FAST, "exact": r.BUILD_EXACT}[args.graph]
    config = {"workload": wl + "_build", "n": n, "dim": dim, "M": m, "ef_construction": efc,
              "builder": args.graph, "graph": GRAPH_LABEL[args.graph], "dataset": dataset_label(ds, r_lat),
              "l2_policy": "inputs larger than L2 (vector slab %d MB), no flush" % (n * dim * 4 >> 20)}
    if args.impl == "reference":
        # The reference's own insert (core.rs:489-599; sequential, one core) at the END of the stream, where the metric
        # is quoted: the graph of the first n - sample nodes is built on the GPU (FAST: seconds) and handed to the oracle,
        # which then applies the rest of the stream, `steps` timed pieces after `warmup` untimed ones.
        per_step = max(20, min(100, int(100 * 128 / dim)))
        sample = per_step * (args.warmup + args.steps)
        base = n - sample
        dev = r.DeviceIndex(dim, m, efc, device=local_rank)
        dev.reserve(base)
        dev.add_batch(x[:base], levels[:base], mode=r.BUILD_FAST)
        orc = oracle.Oracle(dim, m, efc)
        orc.import_graph(x[:base], dev.export_graph())
        dev.close()
        times = []
        for i in range(args.warmup + args.steps):
            lo = base + i * per_step
            t0 = time.perf_counter()
            for j in range(lo, lo + per_step):
                orc.add(x[j], int(levels[j]))
            if i >= args.warmup:
                times.append(time.perf_counter() - t0)
        ips = per_step * args.steps / sum(times)
        emit({"impl": "reference", "metric": "NODE.ADD stream throughput (bulk index build)", "value": ips, "unit": "inserts/s",
              "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / args.steps,
              "higher_is_better": True, "scaling": "replicas only (the sequential stream does not shard)", "vs_baseline": None,
              "dtype": "f32", "data": "synthetic", "config": config,
              "details": {"nodes_per_step": per_step, "base_nodes": base,
                          "base_graph": "first %d nodes built on the GPU by the FAST builder and imported by the oracle" % base},
              "cpu_baseline": {"value": ips, "unit": "inserts/s", "cores": 1, "kind": "port",
                               "sample": "%d NODE.ADDs per step at the end of the stream, oracle (C++ restatement of the reference; "
                                         "the Rust reference cannot be built here); the reference's insert is sequential: "
                                         "1 core" % per_step},
              "e2e": {"value": ips, "unit": "inserts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        return 0
    dev = r.DeviceIndex(dim, m, efc, device=local_rank)
    for opt in args.option:
        name, val = opt.split("=")
        dev.set_option(name, int(val))
    dev.reserve(n)
    # the last `extra` nodes of the stream are held back: the oracle and the device both apply them to the finished graph
    # (CPU baseline on the same points, same distribution, same box — and one more list-for-list comparison)
    extra = 0 if args.no_cpu_baseline else max(50, min(400, int(400 * 128 / dim)))
    n_total, n = n, n - extra
    pieces = args.warmup + args.steps
    piece = n // pieces
    sampler = ClockSampler(local_rank)
    sampler.start()
    sampler.wait_first()
    done, launches0, st_begin, t_begin = 0, 0, None, None
    step_s, all_s = [], []
    for i in range(pieces):
        k = piece if i + 1 < pieces else n - done
        if i == args.warmup:
            torch.cuda.synchronize()
            st_begin, launches0, t_begin = dev.build_stats(), r.launch_count(), time.time()
        t0 = time.perf_counter()
        dev.add_batch(x[done:done + k], levels[done:done + k], mode=mode)     # host buffers in, graph rows on the device out
        dt = time.perf_counter() - t0
        all_s.append(dt)
        if i >= args.warmup:
            step_s.append(dt)
        done += k
        log("piece %d/%d: %d nodes in %.2f s (%.0f inserts/s)" % (i + 1, pieces, k, dt, k / dt))
    t_end = time.time()
    clocks = sampler.stop(t_begin, t_end)
    st_end = dev.build_stats()
    timed_nodes = n - args.warmup * piece
    total_s = sum(step_s)
    ips = timed_nodes / total_s
    d_evals = st_end["dist_evals"] - st_begin["dist_evals"]
    d_waste = st_end["spec_dist_evals_wasted"] - st_begin["spec_dist_evals_wasted"]
    g = dev.export_graph()
    edges = int(g["nbrs"].size)
    # algorithmic bytes of the committed inserts: n_dist * 4 * dim for the vector rows their searches and
    # re-selections evaluate + 4 bytes per adjacency word they leave behind (the rows they read sit in n_dist's shadow)
    alg_bytes = d_evals * 4 * dim + edges * 4 * timed_nodes // n
    peaks = load_peaks()
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = alg_bytes / total_s / 1e9
    cpu = None
    if not args.no_cpu_baseline:
        # the oracle continues the SAME stream from the same graph: the held-back tail, at full size
        xe, le = x[n:n_total], levels[n:n_total]
        orc = oracle.Oracle(dim, m, efc)
        orc.import_graph(x[:n], g)
        t0 = time.perf_counter()
        for i in range(extra):
            orc.add(xe[i], int(le[i]))
        cpu_s = time.perf_counter() - t0
        t0 = time.perf_counter()
        dev.add_batch(xe, le, mode=mode)
        gpu_tail_s = time.perf_counter() - t0
        same = None
        if args.graph != "fast":
            same = all(np.array_equal(dev.node_neighbors(n + j, 0), orc.node_neighbors(n + j, 0)) for j in range(extra))
        cpu = {"value": extra / cpu_s, "unit": "inserts/s", "cores": 1, "kind": "port",
               "sample": "the last %d NODE.ADDs of the stream applied by the oracle to the exported %d-node graph (the reference's "
                         "insert is sequential: 1 core); the device applied the same %d inserts in %.3f s" % (extra, n, extra, gpu_tail_s),
               "device_same_inserts_per_s": extra / gpu_tail_s,
               "lists_equal_gpu": same}
        log("cpu baseline", cpu)
    line = {"metric": "NODE.ADD stream throughput (bulk index build)", "value": ips, "unit": "inserts/s", "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_s / args.steps, "higher_is_better": True,
            "scaling": "replicas only (the sequential stream does not shard)", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config,
            "details": {"nodes_per_step": piece, "warmup_is": "the first %d pieces of the stream (untimed)" % args.warmup,
                        "whole_build_seconds": sum(all_s), "whole_build_inserts_per_s": n / sum(all_s),
                        "build_stats": st_end, "edges": edges, "max_layer": g["max_layer"]},
            "e2e": {"value": ips, "unit": "inserts/s", "h2d_bytes_per_step": int(piece * dim * 4 + piece * 4), "d2h_bytes_per_step": 64,
                    "api": "hnsw_index_add_batch (host vectors in; the timed call IS the public API, so value == e2e)"},
            "gpu_launches": int(r.launch_count() - launches0), "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                         "alg_bytes_per_insert": alg_bytes / timed_nodes,
                         "dist_evals_per_insert": d_evals / timed_nodes, "dist_evals_wasted_per_insert": d_waste / timed_nodes,
                         "kernel": "spec_exec_kernel over all rounds of the timed pieces (the builder is bound by dependent-hop latency "
                                   "and by the dependency chains between inserts, not by bandwidth: DESIGN.md §3.4b)",
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"}}
    if cpu:
        line["cpu_baseline"] = cpu
    emit(line)
    return 0


# ====================================================================================================== search bench

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--bench", default="search", choices=["search", "build"])
    ap.add_argument("--workload", default="1Mx128_M16_efc200", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--graph", default=None, choices=["fast", "spec", "exact"],
                    help="builder of the graph that is searched (search bench, default fast) or that is timed (build bench, default spec)")
    ap.add_argument("--nq", type=int, default=0, help="weak: queries per GPU per step (default 100000); strong: queries of the ONE batch "
                                                      "that is sliced over the GPUs (default 10000, BASELINE configs[3])")
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
    if args.bench == "build":
        args.graph = args.graph or "spec"   # the build bench times the reference-exact stream unless told otherwise
        return bench_build(args, rank, world, local_rank)
    args.graph = args.graph or "fast"
    ref_arm = args.impl == "reference"
    dist = None
    if world > 1 and not ref_arm:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    eff_world = 1 if ref_arm else world

    import redis_hnsw_b200 as r

    wl = args.workload
    n, dim, m, efc, ds, r_lat = WORKLOADS[wl]
    strong = args.scaling == "strong"
    if not args.nq:
        args.nq = 10_000 if strong else 100_000
    nq_total = args.nq if strong else args.nq * eff_world
    n_probe = 2000
    t_setup = time.perf_counter()
    # the query stream is generated for the N the command names, in both arms: the reference arm (rank 0 only) then sees
    # the queries — and reports the recall and `config` — of the arm it is compared with
    n_q = max(args.nq if strong else args.nq * world, 10_000) + n_probe
    if rank == 0 or ref_arm:
        x, q_all, levels = make_data(wl, n_q)
    else:                                 # only the builder needs the vectors: the other ranks receive the index over NCCL
        x, levels, q_all = None, None, np.empty((n_q, dim), np.float32)
    if dist is not None:                  # ... and the queries from rank 0 (they come out of the same generator stream as x)
        tq = torch.from_numpy(q_all).cuda()
        dist.broadcast(tq, 0)
        q_all = tq.cpu().numpy()
        del tq
    q_probe = np.ascontiguousarray(q_all[-n_probe:])
    log("data ready in %.1f s" % (time.perf_counter() - t_setup))
    dev, binfo = build_index(wl, x, levels, local_rank, 0 if ref_arm else rank, eff_world, args.option, graph=args.graph)

    # operating point: smallest ef with recall@10 >= 0.95 (rank 0 decides, everyone follows)
    ef, curve = args.ef, {}
    if rank == 0:
        chosen, curve = pick_ef(dev, x, q_all, wl, force=(args.ef, LITERAL_EF.get(wl, 0)))
        log("recall@10 by ef:", curve, "-> ef =", chosen)
        if not ef:
            ef = chosen or max(curve)
    if dist is not None:
        t = torch.tensor([ef], dtype=torch.int64, device="cuda")
        dist.broadcast(t, 0)
        ef = int(t.item())
    recall = curve.get(ef)
    if not args.cpu_sample:
        args.cpu_sample = max(500, int(40_000 * (128.0 / dim) * (64.0 / max(ef, 16))))

    if ref_arm:
        lo, hi = 0, nq_total
    elif strong:
        lo, hi = r.sharding.query_slice(nq_total, rank, world)
    else:
        lo, hi = rank * args.nq, (rank + 1) * args.nq
    nq = hi - lo
    base_cfg = {"workload": wl, "n": n, "dim": dim, "M": m, "ef_construction": efc, "ef_search": ef, "k": args.k,
                "dataset": dataset_label(ds, r_lat),
                "recall_at_10": recall, "recall_by_ef": curve, "recall_sample": 10000,
                "graph": GRAPH_LABEL[args.graph], "graph_builder": args.graph,
                "l2_policy": "inputs larger than L2 (vector slab %d MB + adjacency; L2 126 MB), no flush" % (n * dim * 4 >> 20)}
    if strong:
        base_cfg.update(batch_queries=nq_total, queries_per_gpu=nq_total // world,
                        step_is="one %d-query batch sliced over the GPUs + all-gather of the result slices" % nq_total)
    else:
        base_cfg.update(queries_per_gpu_per_step=args.nq)

    # ------------------------------------------------------------------ reference arm (CPU oracle, all host threads)
    if ref_arm:
        orc = oracle_on_graph(dev, x, wl)
        cores = os.cpu_count() or 1
        sample = min(args.cpu_sample, nq_total)
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
                "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": dict(base_cfg), "details": {"queries_per_step": sample},
                "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port",
                                 "sample": "%d queries per step, %d threads, oracle (C++ restatement of the reference; "
                                           "the Rust reference cannot be built here) on the same graph" % (sample, cores)},
                "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return 0

    # ------------------------------------------------------------------ B200 arm
    q = np.ascontiguousarray(q_all[lo:hi])
    d_q = torch.from_numpy(q).cuda()
    d_ids = torch.empty((nq, args.k), dtype=torch.int32, device="cuda")
    d_sims = torch.empty((nq, args.k), dtype=torch.float32, device="cuda")
    d_cnt = torch.empty(nq, dtype=torch.int32, device="cuda")
    d_stats = torch.empty((nq, 4), dtype=torch.int32, device="cuda")
    # strong scaling: every rank ends the step holding the WHOLE result (ids, sims and counts packed: one all-gather)
    gather = strong and dist is not None
    width = max(r.sharding.query_slice(nq_total, rr, world)[1] - r.sharding.query_slice(nq_total, rr, world)[0]
                for rr in range(world)) if strong else nq
    d_pack = torch.zeros((width, 2 * args.k + 1), dtype=torch.int32, device="cuda") if gather else None
    d_full = torch.empty((world * width, 2 * args.k + 1), dtype=torch.int32, device="cuda") if gather else None
    # a non-default torch stream: the library treats a NULL stream as "the index's own stream", and torch.cuda.Event
    # only sees the stream it is recorded on
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    def search(stats=False):
        dev.search_batch_device(nq, d_q.data_ptr(), args.k, ef, d_ids.data_ptr(), d_sims.data_ptr(), d_cnt.data_ptr(),
                                d_stats.data_ptr() if stats else 0, stream)

    def exchange():
        if gather:
            d_pack[:nq, :args.k] = d_ids
            d_pack[:nq, args.k:2 * args.k] = d_sims.view(torch.int32)
            d_pack[:nq, 2 * args.k] = d_cnt
            dist.all_gather_into_tensor(d_full, d_pack)

    # algorithmic bytes of one step, from the kernel's own per-query counters (== the oracle's, tests/test_gpu_search.py)
    search(stats=True)
    torch.cuda.synchronize()
    st = d_stats.cpu().numpy().astype(np.int64)
    n_dist, n_adj, n_hops = int(st[:, 0].sum()), int(st[:, 1].sum()), int(st[:, 2].sum())
    alg_bytes = n_dist * 4 * dim + n_adj * 4 + nq * (4 * dim + 8 * args.k)
    # the timed path (no counters requested) is the staged kernel; its lossy visited table may re-evaluate nodes
    evals_done = None
    try:
        dev.set_option("search_impl", 2)
        search(stats=True)
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
    n_warm = 0
    while n_warm < max(args.warmup, 3) or (rank == 0 and dist is None and time.time() - t_warm < 0.5):
        search()                      # (single GPU: stay under load until the sampler has rows from a loaded device)
        exchange()
        n_warm += 1
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    launches0 = r.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    torch.cuda.synchronize()
    cuprof = os.environ.get("HNSW_BENCH_CUPROF") == "1"  # ncu --profile-from-start off: profile the timed region only
    if cuprof:
        torch.cuda.profiler.start()
    t_begin = time.time()
    ev[0].record()
    for i in range(args.steps):
        kev[i][0].record()
        search()
        kev[i][1].record()
        exchange()
        ev[i + 1].record()
    torch.cuda.synchronize()
    t_end = time.time()
    if cuprof:
        torch.cuda.profiler.stop()
    if dist is not None:
        dist.barrier()
    launches = r.launch_count() - launches0
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    total_ms = ev[0].elapsed_time(ev[-1])
    if dist is not None:
        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    qps = nq_total * args.steps / (total_ms / 1e3)
    if gather:   # the gathered block holds every rank's slice: this rank's own part must be what it computed
        own = d_full[rank * width:rank * width + nq, :args.k]
        assert torch.equal(own, d_ids), "all-gathered results differ from the rank's own slice"

    # ------------------------------------------------------------------ e2e: host buffers through the C ABI
    hq = torch.from_numpy(q).pin_memory()
    h_ids = torch.empty((nq, args.k), dtype=torch.int32).pin_memory()
    h_sims = torch.empty((nq, args.k), dtype=torch.float32).pin_memory()
    h_cnt = torch.empty(nq, dtype=torch.int32).pin_memory()
    out = (h_ids.numpy().view(np.uint32), h_sims.numpy(), h_cnt.numpy().view(np.uint32))
    hqn = hq.numpy()
    e2e_steps = max(3, args.steps // 2)

    def e2e_step():
        dev.search_batch(hqn, args.k, ef=ef, out=out)
        if gather:   # the caller of a sharded batch gets the whole answer back
            r.sharding.gather_results(out[0], out[1], out[2], nq_total, rank, world, device="cuda")

    for _ in range(2):
        e2e_step()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_qps = nq_total * e2e_steps / e2e_s
    # the two paths agree
    assert np.array_equal(out[0], d_ids.cpu().numpy().view(np.uint32)), "host-API results differ from the device-API results"

    # ------------------------------------------------------------------ every rank answers a shared probe identically
    p_ids, p_sims, p_cnt = dev.search_batch(q_probe, args.k, ef=ef)
    agree, sums = r.sharding.ranks_agree(r.sharding.result_checksum(p_ids, p_sims), world, device="cuda")
    rank_parity = {"probe_queries": n_probe, "all_ranks_equal_rank0": bool(agree), "checksums": ["%016x" % v for v in sums]}

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    peaks = load_peaks()
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = alg_bytes / (kernel_ms / 1e3) / 1e9
    traffic, traffic_src = None, None
    try:  # DRAM bytes of one launch from the committed ncu --set full capture of this workload (profiles/traffic.json)
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(wl)
        if tr and tr["ef"] == ef and tr.get("graph", "fast") == args.graph and not strong:
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
                "kernel": "search_knn2_kernel (one launch per step on rank 0; duration = CUDA events around the launch on its stream)",
                "kernel_ms": kernel_ms,
                "alg_bytes_per_query": alg_bytes / nq, "dist_evals_per_query": n_dist / nq, "adj_ids_per_query": n_adj / nq,
                "hops_per_query": n_hops / nq,
                "dist_evals_performed_per_query": evals_done}

    cpu, orc = None, None
    if not args.no_cpu_baseline:
        t0 = time.perf_counter()
        orc = oracle_on_graph(dev, x, wl)
        log("oracle import %.1f s" % (time.perf_counter() - t0))
        # rank 0's answers to the shared probe against the oracle (every other rank equals rank 0: the checksums above)
        oids, osims, ocnt, ost, _ = orc.search_batch(q_probe, args.k, ef=ef, threads=os.cpu_count() or 1, stats=True)
        tf = ost[:, 3] == 0
        rank_parity["rank0_equals_oracle"] = bool(np.array_equal(p_ids[tf], oids[tf]) and
                                                  np.array_equal(p_sims[tf].view(np.uint32), osims[tf].view(np.uint32)))
        rank_parity["tie_free_probe_queries"] = int(tf.sum())
    if orc is not None and world == 1:
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

    # `config` names the workload and is the SAME object in this line and in the `--impl reference` line; what was measured
    # on the way (build, per-rank parity probe, replication, batch latency) goes to `details`
    details = dict(build_seconds=binfo["build_seconds"],
                   inserts_per_s=(n / binfo["build_seconds"]) if binfo["build_seconds"] else None, build_stats=binfo["build_stats"],
                   rank_parity=rank_parity)
    if "replicate" in binfo:
        details["replicate"] = binfo["replicate"]
    if strong:
        details["batch_latency_ms"] = total_ms / args.steps
    line = {"metric": "queries/sec @ recall@10>=0.95", "value": qps, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": dict(base_cfg), "details": details,
            "e2e": {"value": e2e_qps, "unit": "queries/s", "h2d_bytes_per_step": int(q.nbytes),
                    "d2h_bytes_per_step": int(nq * args.k * 8 + nq * 4), "steps": e2e_steps,
                    "api": "hnsw_index_search_batch (pinned host buffers)" + (" + all-gather of the result slices" if gather else "")},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline}
    if cpu:
        line["cpu_baseline"] = cpu
    emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
