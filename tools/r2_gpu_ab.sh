#!/bin/bash
# round 2, call AB: full-size (1M) GPU tests incl. the new exact / SPEC insert check
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu > gpurun_out/ab_fullsize.log 2>&1
echo "fullsize rc=$?"; tail -15 gpurun_out/ab_fullsize.log | cut -c1-300
