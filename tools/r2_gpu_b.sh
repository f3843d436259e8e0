#!/bin/bash
# round 2, GPU call B: first hardware run of the SPEC builder + the three drafts of r2-cta-draft
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_spec_build.py -m gpu -q -x --deselect tests/test_gpu_spec_build.py::test_spec_build_100k_matches_the_oracle_fingerprint > gpurun_out/b1_spec_tests.log 2>&1
echo "spec tests rc=$?" | tee -a gpurun_out/b1_spec_tests.log
HNSW_BUILD_TRACE=1 timeout 400 python tools/spec_probe.py --workload 1Mx128_M16_efc200 --limit 400000 --piece 20000 --seconds 120 > gpurun_out/b1_spec_probe.jsonl 2> gpurun_out/b1_spec_probe.err
timeout 600 python -m pytest tests -m gpu -q -k "cta or two_way or ef_up_to_1024" > gpurun_out/b1_draft_tests.log 2>&1
echo "draft tests rc=$?" | tee -a gpurun_out/b1_draft_tests.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_search.py::test_cta_latency_kernel_parity --ignore tests/test_gpu_spec_build.py > gpurun_out/b1_all_tests.log 2>&1
echo "all tests rc=$?" | tee -a gpurun_out/b1_all_tests.log
timeout 600 python tools/tune_search.py --workload 1Mx128_M16_efc200 --ef 64 --steps 10 --grid "search_impl=2;recent_ways=1,2,1,2" --out gpurun_out/b1_way2_ef64.json > gpurun_out/b1_way2.log 2>&1
timeout 600 python tests/bench_ops.py --only search --n-search 1500 > gpurun_out/b1_ops_base.json 2> gpurun_out/b1_ops_base.err
timeout 600 python tests/bench_ops.py --only search --n-search 1500 --option search_cta=1 > gpurun_out/b1_ops_cta.json 2> gpurun_out/b1_ops_cta.err
tail -3 gpurun_out/b1_spec_tests.log gpurun_out/b1_draft_tests.log gpurun_out/b1_all_tests.log
tail -4 gpurun_out/b1_spec_probe.jsonl
tail -5 gpurun_out/b1_way2.log
