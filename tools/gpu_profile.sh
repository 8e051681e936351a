#!/bin/bash
# Round-end measurement pass (run on the GPU box through gpurun):
#   gpurun --timeout 1500 -- 'bash tools/gpu_profile.sh r1b'
# Writes everything under gpurun_out/; tools/summarise_profiles.py condenses it into profiles/.
TAG=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/gpu_$TAG.txt
# 1. bench (timed, not under a profiler)
timeout 900 python bench.py > gpurun_out/bench_$TAG.log 2>gpurun_out/bench_$TAG.err
# 2. kernel-only timings
for k in pv wind heat; do for s in small big odd oddpad; do
  timeout 120 python tools/prof_pv.py $k $s 7
done; done > gpurun_out/prof_$TAG.jsonl 2>/dev/null
# 3. launch list of the bench command
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/launches_bench_$TAG.csv python bench.py --steps 2 --warmup 3 --no-extra \
  > gpurun_out/bench_under_ncu_$TAG.log 2>&1
# 4. one full capture per fused kernel (kernel-name filter, one launch each)
for ks in "pv small" "pv big" "wind small" "heat small"; do
  set -- $ks
  kn=k_fused_reduce; [ $1 = heat ] && kn=k_heat
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$kn -s 3 -c 1 \
    -o gpurun_out/prof_$1_$2_$TAG -f python tools/prof_pv.py $1 $2 2 > gpurun_out/ncu_$1_$2_$TAG.log 2>&1
done
ls -la gpurun_out | tail -20
