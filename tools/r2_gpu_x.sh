#!/bin/bash
# round 2, call X: the 1M build bench line with the CPU baseline on the held-back tail of the stream, and its reference arm
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python bench.py --bench build --steps 20 --warmup 3 > $O/x_bench_build.json 2> $O/x_bench_build.err
echo "bench build rc=$?"; cut -c1-400 $O/x_bench_build.json; tail -3 $O/x_bench_build.err | cut -c1-400
timeout 600 python bench.py --bench build --impl reference --steps 10 --warmup 2 > $O/x_build_ref.json 2> $O/x_build_ref.err
echo "build ref rc=$?"; cut -c1-300 $O/x_build_ref.json
