#!/bin/bash
# round 2, 2-GPU call: the multi-rank bench paths (weak = what the driver's SCALE run launches; strong on the 1M workload)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > $O/k1_weak2.json 2> $O/k1_weak2.err
echo "weak rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --scaling strong --nq 10000 --steps 50 --warmup 5 > $O/k1_strong2_1M.json 2> $O/k1_strong2_1M.err
echo "strong rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > $O/k1_ref2.json 2> $O/k1_ref2.err
echo "ref rc=$?"
cut -c1-500 $O/k1_weak2.json; cut -c1-500 $O/k1_strong2_1M.json; cut -c1-200 $O/k1_ref2.json; tail -3 $O/k1_weak2.err $O/k1_strong2_1M.err | cut -c1-300
