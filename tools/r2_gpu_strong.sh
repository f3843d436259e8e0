#!/bin/bash
# round 2: BASELINE configs[3] literally — 10M x 128, ONE 10 000-query batch sliced over N GPUs, results all-gathered in the
# timed region.  usage: tools/r2_gpu_strong.sh N [extra bench args]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$1; shift
if [ "$N" = "1" ]; then
  timeout 1500 python bench.py --gpus 1 --scaling strong --workload 10Mx128_M16_efc200 --nq 10000 --steps 50 --warmup 5 "$@" > gpurun_out/strong_n$N.json 2> gpurun_out/strong_n$N.err
else
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --scaling strong --workload 10Mx128_M16_efc200 --nq 10000 --steps 50 --warmup 5 "$@" > gpurun_out/strong_n$N.json 2> gpurun_out/strong_n$N.err
fi
echo "rc=$?"; cut -c1-400 gpurun_out/strong_n$N.json; grep "bench\]" gpurun_out/strong_n$N.err | tail -6 | cut -c1-300
