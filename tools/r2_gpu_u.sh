#!/bin/bash
# round 2, call U: ncu --set full of the second version of the SPEC kernels (one launch each, 100K x 128 stream)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:spec_exec -s 1500 -c 1 -f -o $O/prof_r2b_spec_exec python tools/spec_probe.py --workload 100Kx128_M16_efc200 --limit 30000 --piece 30000 > /dev/null 2> $O/u_prof_exec.err
echo "ncu exec rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:spec_commit -s 1500 -c 1 -f -o $O/prof_r2b_spec_commit python tools/spec_probe.py --workload 100Kx128_M16_efc200 --limit 30000 --piece 30000 > /dev/null 2> $O/u_prof_commit.err
echo "ncu commit rc=$?"
ls -la $O/prof_r2b_* ; tail -2 $O/u_prof_exec.err $O/u_prof_commit.err | cut -c1-200
