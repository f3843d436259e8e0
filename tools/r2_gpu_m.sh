#!/bin/bash
# round 2, call M: window policy of the SPEC builder under dependency-level validation + K1/K2 split
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python tools/spec_tune.py --base 900000 --piece 8000 --grid "spec_mult=15,20,30,40,60,90,10" > $O/m_mult.jsonl 2> $O/m_mult.err
echo "mult rc=$?"
cat $O/m_mult.jsonl; tail -3 $O/m_mult.err
HNSW_BUILD_TRACE=1 timeout 600 python tools/spec_tune.py --base 940000 --piece 12000 --grid "spec_validation=2" > $O/m_trace.jsonl 2> $O/m_trace.err
echo "trace rc=$?"
cat $O/m_trace.jsonl; grep "last 128" $O/m_trace.err | tail -4 | cut -c1-300
