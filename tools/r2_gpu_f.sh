#!/bin/bash
# round 2, GPU call F: whole GPU suite (incl. the 100k oracle fingerprint), smoke, default bench on both graphs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/f1_all_tests.log 2>&1
echo "all tests rc=$?" | tee -a gpurun_out/f1_all_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/f1_smoke.log 2>&1
echo "smoke rc=$?" | tee -a gpurun_out/f1_smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/f1_bench_fast.json 2> gpurun_out/f1_bench_fast.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/f1_bench_ref.json 2> gpurun_out/f1_bench_ref.err
tail -6 gpurun_out/f1_all_tests.log; tail -2 gpurun_out/f1_smoke.log; cut -c1-600 gpurun_out/f1_bench_fast.json; tail -3 gpurun_out/f1_bench_fast.err
