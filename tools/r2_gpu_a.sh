#!/bin/bash
# round 2, GPU call A: the three drafts of r2-cta-draft on hardware for the first time
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "cta or two_way or ef_up_to_1024" > gpurun_out/a1_draft_tests.log 2>&1
echo "draft tests rc=$?" | tee -a gpurun_out/a1_draft_tests.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_search.py::test_cta_latency_kernel_parity > gpurun_out/a1_all_tests.log 2>&1
echo "all tests rc=$?" | tee -a gpurun_out/a1_all_tests.log
timeout 600 python tools/tune_search.py --workload 1Mx128_M16_efc200 --ef 64 --steps 10 --grid "search_impl=2;recent_ways=1,2,1,2" --out gpurun_out/a1_way2_ef64.json > gpurun_out/a1_way2.log 2>&1
timeout 600 python tests/bench_ops.py --only search --n-search 1500 > gpurun_out/a1_ops_base.json 2> gpurun_out/a1_ops_base.err
timeout 600 python tests/bench_ops.py --only search --n-search 1500 --option search_cta=1 > gpurun_out/a1_ops_cta.json 2> gpurun_out/a1_ops_cta.err
tail -3 gpurun_out/a1_draft_tests.log gpurun_out/a1_all_tests.log
cat gpurun_out/a1_way2.log | tail -6
