#!/bin/bash
# round 2, GPU call C: where does a SPEC round spend its time?
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
HNSW_BUILD_TRACE=1 timeout 200 python tools/spec_probe.py --workload 100Kx128_M16_efc200 --limit 30000 --piece 10000 > gpurun_out/c1_spec_probe.jsonl 2> gpurun_out/c1_spec_probe.err
HNSW_BUILD_TRACE=1 timeout 100 python tools/spec_probe.py --workload 100Kx128_M16_efc200 --limit 6000 --piece 3000 --option spec_window=1 > gpurun_out/c1_spec_w1.jsonl 2> gpurun_out/c1_spec_w1.err
timeout 100 python tools/spec_probe.py --workload 100Kx128_M16_efc200 --limit 6000 --piece 3000 --mode exact > gpurun_out/c1_exact.jsonl 2> gpurun_out/c1_exact.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/c1_spec_launches.csv python tools/spec_probe.py --workload 100Kx128_M16_efc200 --limit 3000 --piece 3000 > /dev/null 2>&1
tail -3 gpurun_out/c1_spec_probe.err; tail -2 gpurun_out/c1_spec_w1.err; tail -1 gpurun_out/c1_spec_w1.jsonl; tail -1 gpurun_out/c1_exact.jsonl
