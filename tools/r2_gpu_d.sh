#!/bin/bash
# round 2, GPU call D: K1 time against the window size, with per-warp placement and run times
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for w in 1 2 4 8 16; do
  HNSW_BUILD_TRACE=1 timeout 120 python tools/spec_probe.py --workload 100Kx128_M16_efc200 --limit 8192 --piece 8192 --option spec_window=$w > gpurun_out/d1_w$w.jsonl 2> gpurun_out/d1_w$w.err
  echo "window $w"; tail -4 gpurun_out/d1_w$w.err; tail -1 gpurun_out/d1_w$w.jsonl
done
