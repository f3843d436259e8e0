#!/bin/bash
# round 2, call AC: smoke() with the SPEC stream in it
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ac_smoke.log 2>&1
echo "smoke rc=$?"; tail -5 gpurun_out/ac_smoke.log | cut -c1-300
