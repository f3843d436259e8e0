#!/bin/bash
# round 2, call S: SPEC builder under compute-sanitizer, then the 1M build bench line with the final defaults
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_spec.py > $O/s_$tool.log 2>&1
  echo "$tool rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^ok" $O/s_$tool.log | tail -14
done
timeout 900 python bench.py --bench build --steps 20 --warmup 3 > $O/s_bench_build.json 2> $O/s_bench_build.err
echo "bench build rc=$?"
cut -c1-700 $O/s_bench_build.json; tail -3 $O/s_bench_build.err
