#!/bin/bash
# round 2, call L: dependency-level validation of the SPEC builder — parity tests, then A/B against row-level validation
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_spec_build.py -x -q -m gpu > $O/l_spec_tests.log 2>&1
echo "spec tests rc=$?"
tail -15 $O/l_spec_tests.log
timeout 600 python tools/spec_tune.py --base 940000 --piece 8000 --grid "spec_validation=1,2,1,2" > $O/l_ab.jsonl 2> $O/l_ab.err
echo "ab rc=$?"
cat $O/l_ab.jsonl; tail -3 $O/l_ab.err
HNSW_BUILD_TRACE=1 timeout 600 python tools/spec_tune.py --base 940000 --piece 12000 --grid "spec_validation=2" > $O/l_trace.jsonl 2> $O/l_trace.err
echo "trace rc=$?"
cat $O/l_trace.jsonl; grep "last 1024" $O/l_trace.err | tail -5 | cut -c1-300
