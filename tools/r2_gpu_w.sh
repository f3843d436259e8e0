#!/bin/bash
# round 2, call W: the build bench after the config / details split, and its reference arm
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python bench.py --bench build --impl reference --steps 4 --warmup 1 > $O/w_build_ref.json 2> $O/w_build_ref.err
echo "build ref rc=$?"; cut -c1-900 $O/w_build_ref.json; tail -2 $O/w_build_ref.err | cut -c1-200
timeout 600 python bench.py --bench build --workload 100Kx128_M16_efc200 --steps 4 --warmup 1 > $O/w_build_100k.json 2> $O/w_build_100k.err
echo "build 100k rc=$?"; cut -c1-900 $O/w_build_100k.json; tail -2 $O/w_build_100k.err | cut -c1-200
python - <<'PY'
import json
a=json.loads(open('gpurun_out/w_build_ref.json').read().strip().splitlines()[-1])
print("ref line keys:", sorted(a))
PY
