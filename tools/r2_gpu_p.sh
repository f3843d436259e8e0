#!/bin/bash
# round 2, call P: time budget (suspended executions) + K2 stale floor
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_spec_build.py -x -q -m gpu > $O/p_spec_tests.log 2>&1
echo "spec tests rc=$?"
tail -15 $O/p_spec_tests.log
timeout 600 python tools/spec_tune.py --base 900000 --piece 8000 --grid "spec_budget_us=-1,1300,-1,1300,1000,1600,1150" > $O/p_ab.jsonl 2> $O/p_ab.err
echo "ab rc=$?"
cat $O/p_ab.jsonl; tail -3 $O/p_ab.err
HNSW_BUILD_TRACE=1 timeout 600 python tools/spec_tune.py --base 940000 --piece 12000 --grid "spec_ahead=0" > $O/p_trace.jsonl 2> $O/p_trace.err
echo "trace rc=$?"
cat $O/p_trace.jsonl; grep "last 128" $O/p_trace.err | tail -4 | cut -c1-300
