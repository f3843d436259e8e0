#!/bin/bash
# round 2, call O: where an execution spends its time (K1 diagnostics: level, us in searches, us in re-selections)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
HNSW_BUILD_TRACE=1 timeout 600 python tools/spec_tune.py --base 940000 --piece 12000 --grid "spec_ahead=0" > $O/o_trace.jsonl 2> $O/o_trace.err
echo "trace rc=$?"
cat $O/o_trace.jsonl; grep "last 128" $O/o_trace.err | tail -4 | cut -c1-300
