#!/bin/bash
# round 2, GPU call H: wide (unsorted) candidate lists: parity suite, recall-ef-QPS curve, SPEC K1 time, op latencies
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_spec_build.py::test_spec_build_100k_matches_the_oracle_fingerprint > gpurun_out/h1_tests.log 2>&1
echo "tests rc=$?" | tee -a gpurun_out/h1_tests.log
timeout 600 python tools/curve.py --efs 32,64,96,128,200,256,400,512 > gpurun_out/h1_curve.json 2> gpurun_out/h1_curve.err
HNSW_BUILD_TRACE=1 timeout 300 python tools/spec_probe.py --workload 100Kx128_M16_efc200 --piece 25000 > gpurun_out/h1_spec_probe.jsonl 2> gpurun_out/h1_spec_probe.err
timeout 400 python tests/bench_ops.py --n-search 1000 --n-add 500 --n-del 200 > gpurun_out/h1_ops.json 2> gpurun_out/h1_ops.err
tail -5 gpurun_out/h1_tests.log; grep curve gpurun_out/h1_curve.err | cut -c1-250; cut -c1-300 gpurun_out/h1_spec_probe.jsonl; cut -c1-1500 gpurun_out/h1_ops.json
