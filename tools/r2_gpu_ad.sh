#!/bin/bash
# round 2, call AD: FAST builder tests with the drop counters asserted
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_build.py tests/test_gpu_fullsize.py -x -q -m gpu -k "fast or params" > gpurun_out/ad_fast.log 2>&1
echo "fast tests rc=$?"; tail -6 gpurun_out/ad_fast.log | cut -c1-300
