#!/bin/bash
# GPU pass of round 2: parity suite + kernel A/B (variant 0 = staged reduce with the functor's chunk
# default (shuffle reduce for pv / wind), 2 = staged reduce with 8-step chunks, 3 = with 16-step chunks)
# + one ncu capture of the staged wind and PV kernels.  <= 10 GB of VRAM, small host memory.
TAG=${1:-r2p}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_full_$TAG.log 2>&1
cp gpurun_out/parity_errors.json gpurun_out/parity_errors_$TAG.json 2>/dev/null
for v in 0 2 3; do for k in wind pv heat; do for s in small big; do
  ATL_VARIANT=$v timeout 120 python tools/prof_pv.py $k $s 7
done; done; done > gpurun_out/prof_$TAG.jsonl 2>gpurun_out/prof_$TAG.err
for k in wind pv; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_fused_reduce -s 3 -c 1 \
    -o gpurun_out/prof_${k}_big_$TAG -f env ATL_VARIANT=2 python tools/prof_pv.py $k big 2 > gpurun_out/ncu_${k}_big_$TAG.log 2>&1
done
grep -E "passed|failed" gpurun_out/pytest_full_$TAG.log | tail -3; grep -E "^FAILED|^ERROR" gpurun_out/pytest_full_$TAG.log | head
cat gpurun_out/prof_$TAG.jsonl; tail -3 gpurun_out/prof_$TAG.err; ls -la gpurun_out/*.ncu-rep | tail -3
