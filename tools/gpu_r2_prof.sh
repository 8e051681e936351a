#!/bin/bash
# GPU pass 2 of round 2: parity suite again + kernel A/B (variant 0 = staged reduce with the functor's
# chunk length, 1 = round-1 shuffle reduce, 2 = the other chunk length).  <= 10 GB of VRAM, small host memory.
TAG=${1:-r2p}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_full_$TAG.log 2>&1
cp gpurun_out/parity_errors.json gpurun_out/parity_errors_$TAG.json 2>/dev/null
for v in 0 1 2; do for k in wind pv heat; do for s in small big; do
  ATL_VARIANT=$v timeout 120 python tools/prof_pv.py $k $s 7
done; done; done > gpurun_out/prof_$TAG.jsonl 2>gpurun_out/prof_$TAG.err
grep -E "passed|failed" gpurun_out/pytest_full_$TAG.log | tail -3; grep -E "^FAILED|^ERROR" gpurun_out/pytest_full_$TAG.log | head
cat gpurun_out/prof_$TAG.jsonl; tail -3 gpurun_out/prof_$TAG.err
