#!/bin/bash
# round 2, call Z (2 GPUs): both arms print the same config at N = 2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 5 --warmup 3 > $O/z_weak2.json 2> $O/z_weak2.err
echo "weak rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/z_ref2.json 2> $O/z_ref2.err
echo "ref rc=$?"
python - <<'PY'
import json
a=json.loads(open('gpurun_out/z_weak2.json').read().strip().splitlines()[-1])
b=json.loads(open('gpurun_out/z_ref2.json').read().strip().splitlines()[-1])
print("same config:", a["config"] == b["config"], a["value"], b["value"])
PY
