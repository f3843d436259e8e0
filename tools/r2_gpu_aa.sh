#!/bin/bash
# round 2, call AA: Redis module tests after HNSW.NODE.MADD's default became the exact builder
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_redis_module.py -x -q -m gpu > gpurun_out/aa_module_tests.log 2>&1
echo "module tests rc=$?"; tail -4 gpurun_out/aa_module_tests.log
