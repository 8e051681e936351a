#!/bin/bash
# Round-2 profiling pass (run on the GPU box through gpurun): kernel timings, ncu full captures of
# the hot kernels, the launch list of the bench command.  Everything lands in gpurun_out/.
TAG=${1:-r2b}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/gpu_$TAG.txt
for k in wind pv heat; do for s in small big; do
  timeout 120 python tools/prof_pv.py $k $s 7
done; done > gpurun_out/prof_$TAG.jsonl 2>gpurun_out/prof_$TAG.err
for ks in "wind big" "pv big" "heat big" "wind small" "pv small"; do
  set -- $ks
  kn=k_fused_reduce; [ $1 = heat ] && kn=k_heat
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$kn -s 3 -c 1 \
    -o gpurun_out/prof_$1_$2_$TAG -f python tools/prof_pv.py $1 $2 2 > gpurun_out/ncu_$1_$2_$TAG.log 2>&1
done
# launch list of the bench command (our kernels only; the step is one memset + one kernel)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 60 --csv \
  --log-file gpurun_out/launches_bench_$TAG.csv python bench.py --steps 2 --warmup 3 --no-extra \
  > gpurun_out/bench_under_ncu_$TAG.log 2>&1
ls -la gpurun_out | tail -15; cat gpurun_out/prof_$TAG.jsonl
