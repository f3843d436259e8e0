#!/bin/bash
# round 2, call T: the suspension path after the relink fix — parity tests, a budget stress, synccheck
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_spec_build.py -x -q -m gpu > $O/t_spec_tests.log 2>&1
echo "spec tests rc=$?"; tail -4 $O/t_spec_tests.log
SPEC_STRESS=1 timeout 300 python tools/sanitize_spec.py > $O/t_stress.log 2>&1
echo "stress rc=$?"; tail -18 $O/t_stress.log | cut -c1-200
SANITIZE_SPEC_FAST=1 timeout 400 compute-sanitizer --tool synccheck --error-exitcode 9 python tools/sanitize_spec.py > $O/t_synccheck.log 2>&1
echo "synccheck rc=$?"; grep -E "ERROR SUMMARY|^ok|Error" $O/t_synccheck.log | tail -10
