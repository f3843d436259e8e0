#!/bin/bash
# round 2, call Y (2 GPUs): the multi-rank bench paths after the config / details split
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 10 --warmup 3 > $O/y_weak2.json 2> $O/y_weak2.err
echo "weak rc=$?"; cut -c1-300 $O/y_weak2.json; tail -2 $O/y_weak2.err | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --scaling strong --steps 20 --warmup 3 > $O/y_strong2.json 2> $O/y_strong2.err
echo "strong rc=$?"; cut -c1-300 $O/y_strong2.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > $O/y_ref2.json 2> $O/y_ref2.err
echo "ref rc=$?"; cut -c1-200 $O/y_ref2.json
python - <<'PY'
import json
a=json.loads(open('gpurun_out/y_weak2.json').read().strip().splitlines()[-1])
b=json.loads(open('gpurun_out/y_ref2.json').read().strip().splitlines()[-1])
print("same config:", a["config"] == b["config"], "| rank parity:", a["details"]["rank_parity"])
PY
