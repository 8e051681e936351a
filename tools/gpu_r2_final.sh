#!/bin/bash
# final pass: parity suite + bench with the product library; the heat L2-prefetch build beside it
TAG=${1:-r2e}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_full_$TAG.log 2>&1
cp gpurun_out/parity_errors.json gpurun_out/parity_errors_$TAG.json 2>/dev/null
for lib in build_variants/*.so; do
  [ -e "$lib" ] || continue
  ATL_LIB_PATH=$lib timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -k "heat or cooling" > gpurun_out/pytest_heat_variant_$TAG.log 2>&1
  for s in small big; do timeout 120 python tools/prof_pv.py heat $s 9; ATL_LIB_PATH=$lib timeout 120 python tools/prof_pv.py heat $s 9; done > gpurun_out/prof_$TAG.jsonl 2>gpurun_out/prof_$TAG.err
done
timeout 1200 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "rc=$?" >> gpurun_out/bench_$TAG.err
grep -E "passed|failed" gpurun_out/pytest_full_$TAG.log gpurun_out/pytest_heat_variant_$TAG.log | tail -3; grep -E "^FAILED|^ERROR" gpurun_out/pytest_full_$TAG.log | head
cat gpurun_out/prof_$TAG.jsonl | cut -c1-220
head -c 1800 gpurun_out/bench_$TAG.json; echo; tail -3 gpurun_out/bench_$TAG.err
