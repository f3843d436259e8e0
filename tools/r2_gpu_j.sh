#!/bin/bash
# round 2, GPU call J: validation of the last changes (memory-backed list class, config-3 parity case, CTA kernel on by
# default, hand-derived RDB fixture through the module) + synccheck again + op latencies
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > $O/j1_tests.log 2>&1
echo "tests rc=$?" | tee -a $O/j1_tests.log
timeout 300 python __graft_entry__.py smoke > $O/j1_smoke.log 2>&1
echo "smoke rc=$?" | tee -a $O/j1_smoke.log
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python tools/sanitize_smoke.py > $O/j1_synccheck.log 2>&1
echo "synccheck rc=$?" | tee -a $O/j1_synccheck.log
timeout 400 python tests/bench_ops.py --n-search 1500 --n-add 600 --n-del 200 > $O/j1_ops.json 2> $O/j1_ops.err
tail -8 $O/j1_tests.log
tail -2 $O/j1_smoke.log
tail -3 $O/j1_synccheck.log
cut -c1-1800 $O/j1_ops.json
