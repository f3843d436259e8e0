#!/bin/bash
# round 2, call R: the whole GPU suite on the final SPEC builder, smoke, window policy once more
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > $O/r_all_tests.log 2>&1
echo "all tests rc=$?"
tail -6 $O/r_all_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r_smoke.log 2>&1
echo "smoke rc=$?"; tail -2 $O/r_smoke.log
timeout 600 python tools/spec_tune.py --base 900000 --piece 8000 --grid "spec_mult=30,40,50,30,40,50,25" > $O/r_mult.jsonl 2> $O/r_mult.err
echo "mult rc=$?"
cat $O/r_mult.jsonl; tail -3 $O/r_mult.err
