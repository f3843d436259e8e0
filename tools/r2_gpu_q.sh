#!/bin/bash
# round 2, call Q: batched sweeps + 204-register K1 + head fix: parity, budget A/B, then the 1M build bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_spec_build.py -x -q -m gpu > $O/q_spec_tests.log 2>&1
echo "spec tests rc=$?"
tail -5 $O/q_spec_tests.log
timeout 600 python tools/spec_tune.py --base 900000 --piece 8000 --grid "spec_budget_us=1300,1150,1000,-1,1300,1150,1000" > $O/q_ab.jsonl 2> $O/q_ab.err
echo "ab rc=$?"
cat $O/q_ab.jsonl; tail -3 $O/q_ab.err
timeout 900 python bench.py --bench build --steps 20 --warmup 3 > $O/q_bench_build.json 2> $O/q_bench_build.err
echo "bench build rc=$?"
cut -c1-1200 $O/q_bench_build.json; tail -3 $O/q_bench_build.err
