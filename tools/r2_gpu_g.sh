#!/bin/bash
# round 2, GPU call G: lookahead parity + latency, SPEC-vs-EXACT-vs-oracle diagnosis between 30k and 60k nodes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_search.py tests/test_gpu_build.py tests/test_gpu_delete.py tests/test_gpu_spec_build.py tests/test_gpu_redis_module.py -m gpu -q --deselect tests/test_gpu_spec_build.py::test_spec_build_100k_matches_the_oracle_fingerprint > gpurun_out/g1_tests.log 2>&1
echo "tests rc=$?" | tee -a gpurun_out/g1_tests.log
timeout 600 python tools/diag_spec_100k.py 30000 60000 5000 > gpurun_out/g1_diag.log 2>&1
timeout 300 python tests/bench_ops.py --only search --n-search 1500 > gpurun_out/g1_ops_la.json 2> gpurun_out/g1_ops_la.err
timeout 300 python tests/bench_ops.py --only search --n-search 1500 --option lookahead=0 > gpurun_out/g1_ops_nola.json 2> gpurun_out/g1_ops_nola.err
HNSW_BUILD_TRACE=1 timeout 300 python tools/spec_probe.py --workload 100Kx128_M16_efc200 --piece 25000 > gpurun_out/g1_spec_probe.jsonl 2> gpurun_out/g1_spec_probe.err
tail -5 gpurun_out/g1_tests.log; cat gpurun_out/g1_diag.log | cut -c1-250; cut -c1-400 gpurun_out/g1_spec_probe.jsonl
