#!/bin/bash
# wind kernel A/B pass: parity suite, then the shuffle-reduce wind kernel from the product
# library next to the experiment builds under build_variants/ (tools/build_variants.sh)
TAG=${1:-r2w}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_full_$TAG.log 2>&1
cp gpurun_out/parity_errors.json gpurun_out/parity_errors_$TAG.json 2>/dev/null
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv > gpurun_out/gpu_$TAG.txt
: > gpurun_out/prof_$TAG.jsonl
for rep in 1 2; do
for lib in default build_variants/*.so; do for s in small big; do
  if [ $lib = default ]; then timeout 120 python tools/prof_pv.py wind $s 9
  else ATL_LIB_PATH=$lib timeout 120 python tools/prof_pv.py wind $s 9; fi
done; done
done >> gpurun_out/prof_$TAG.jsonl 2>gpurun_out/prof_$TAG.err
for k in pv heat windsum; do timeout 120 python tools/prof_pv.py $k big 7; done >> gpurun_out/prof_$TAG.jsonl 2>>gpurun_out/prof_$TAG.err
grep -E "passed|failed" gpurun_out/pytest_full_$TAG.log | tail -3; grep -E "^FAILED|^ERROR" gpurun_out/pytest_full_$TAG.log | head
cat gpurun_out/prof_$TAG.jsonl; tail -3 gpurun_out/prof_$TAG.err
