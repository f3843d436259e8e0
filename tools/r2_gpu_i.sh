#!/bin/bash
# round 2, GPU call I: the evidence run — whole GPU suite, SPEC window policy, build bench, headline on both graphs,
# config 3, recall-ef-QPS curve, ncu launch list and full captures, sanitizers
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > $O/i1_tests.log 2>&1
echo "tests rc=$?" | tee -a $O/i1_tests.log
timeout 300 python tools/spec_tune.py --base 940000 --piece 8000 --grid "spec_mult=15,20,30,40,60,80" > $O/i1_spec_tune.jsonl 2> $O/i1_spec_tune.err
timeout 900 python bench.py --bench build --steps 20 --warmup 3 > $O/i1_bench_build.json 2> $O/i1_bench_build.err
timeout 900 python bench.py --graph spec --steps 20 --warmup 3 > $O/i1_bench_specgraph.json 2> $O/i1_bench_specgraph.err
timeout 400 python bench.py --steps 20 --warmup 3 > $O/i1_bench_fast.json 2> $O/i1_bench_fast.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $O/i1_bench_ref.json 2> $O/i1_bench_ref.err
timeout 600 python bench.py --workload 1Mx768_M32_efc400 --ef 128 --steps 10 --warmup 3 > $O/i1_bench_cfg3_ef128.json 2> $O/i1_bench_cfg3_ef128.err
timeout 600 python bench.py --workload 1Mx768_M32_efc400 --steps 10 --warmup 3 > $O/i1_bench_cfg3.json 2> $O/i1_bench_cfg3.err
timeout 400 python tools/curve.py --efs 16,32,48,64,96,128,200,256,400,512,768,1024 > $O/i1_curve.json 2> $O/i1_curve.err
HNSW_BENCH_CUPROF=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file $O/i1_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2> $O/i1_launches.err
HNSW_BENCH_CUPROF=1 timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:search_knn2 -c 1 -o $O/prof_r2_search python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2> $O/i1_prof_search.err
HNSW_BENCH_CUPROF=1 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:search_knn2 -c 1 -o $O/prof_r2_search_768 python bench.py --workload 1Mx768_M32_efc400 --ef 128 --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2> $O/i1_prof_search768.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:spec_exec -s 3000 -c 1 -o $O/prof_r2_spec_exec python tools/spec_probe.py --workload 100Kx128_M16_efc200 --limit 12000 --piece 12000 > /dev/null 2> $O/i1_prof_spec.err
for tool in memcheck synccheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_smoke.py > $O/i1_$tool.log 2>&1
  echo "$tool rc=$?" | tee -a $O/i1_$tool.log
done
tail -3 $O/i1_tests.log; cat $O/i1_spec_tune.jsonl; for f in build specgraph fast cfg3_ef128 cfg3; do cut -c1-330 $O/i1_bench_$f.json; done
tail -2 $O/i1_memcheck.log $O/i1_synccheck.log $O/i1_racecheck.log
