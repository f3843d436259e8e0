#!/bin/bash
# round 2, GPU call E: SPEC throughput curve to 1M after the warp-divergence fix + the whole GPU suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
HNSW_BUILD_TRACE=1 timeout 420 python tools/spec_probe.py --workload 1Mx128_M16_efc200 --piece 50000 --seconds 300 > gpurun_out/e1_spec_probe.jsonl 2> gpurun_out/e1_spec_probe.err
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/e1_all_tests.log 2>&1
echo "all tests rc=$?" | tee -a gpurun_out/e1_all_tests.log
tail -4 gpurun_out/e1_all_tests.log; cat gpurun_out/e1_spec_probe.jsonl | cut -c1-200
