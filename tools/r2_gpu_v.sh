#!/bin/bash
# round 2, call V: final validation of the tree — whole GPU suite, smoke, default bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > $O/v_all_tests.log 2>&1
echo "all tests rc=$?"; tail -4 $O/v_all_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/v_smoke.log 2>&1
echo "smoke rc=$?"; tail -2 $O/v_smoke.log
timeout 600 python bench.py > $O/v_bench.json 2> $O/v_bench.err
echo "bench rc=$?"; cut -c1-600 $O/v_bench.json; tail -2 $O/v_bench.err | cut -c1-200
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/v_bench_ref.json 2> $O/v_bench_ref.err
echo "ref rc=$?"; cut -c1-300 $O/v_bench_ref.json
python - <<'PY'
import json
a=json.loads(open('gpurun_out/v_bench.json').read().strip().splitlines()[-1])
b=json.loads(open('gpurun_out/v_bench_ref.json').read().strip().splitlines()[-1])
print("same config:", a["config"] == b["config"])
PY
