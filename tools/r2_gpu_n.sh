#!/bin/bash
# round 2, call N: checkpointed upper levels (prepare ahead of the window) + 16-warp commit kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_spec_build.py -x -q -m gpu > $O/n_spec_tests.log 2>&1
echo "spec tests rc=$?"
tail -15 $O/n_spec_tests.log
timeout 600 python tools/spec_tune.py --base 900000 --piece 8000 --grid "spec_ahead=-1,0,-1,0,64,400" > $O/n_ab.jsonl 2> $O/n_ab.err
echo "ab rc=$?"
cat $O/n_ab.jsonl; tail -3 $O/n_ab.err
HNSW_BUILD_TRACE=1 timeout 600 python tools/spec_tune.py --base 940000 --piece 12000 --grid "spec_ahead=0" > $O/n_trace.jsonl 2> $O/n_trace.err
echo "trace rc=$?"
cat $O/n_trace.jsonl; grep "last 128" $O/n_trace.err | tail -4 | cut -c1-300
