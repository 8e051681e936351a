#!/bin/bash
# quick GPU pass: parity suite + default-kernel timings (variant 0) next to the staged variant (2)
TAG=${1:-r2u}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_full_$TAG.log 2>&1
cp gpurun_out/parity_errors.json gpurun_out/parity_errors_$TAG.json 2>/dev/null
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv > gpurun_out/gpu_$TAG.txt
for v in 0 2; do for k in pv wind heat; do for s in small big; do
  ATL_VARIANT=$v timeout 120 python tools/prof_pv.py $k $s 7
done; done; done > gpurun_out/prof_$TAG.jsonl 2>gpurun_out/prof_$TAG.err
for k in pvsum pvcube windsum; do timeout 120 python tools/prof_pv.py $k c5 7; done >> gpurun_out/prof_$TAG.jsonl 2>>gpurun_out/prof_$TAG.err
grep -E "passed|failed" gpurun_out/pytest_full_$TAG.log | tail -3; grep -E "^FAILED|^ERROR" gpurun_out/pytest_full_$TAG.log | head
cat gpurun_out/prof_$TAG.jsonl; tail -3 gpurun_out/prof_$TAG.err
