#!/bin/bash
# Round-2 profiling pass (run on the GPU box through gpurun): ncu full captures of the final hot kernels
# and the launch list of the bench command.  Everything lands in gpurun_out/; tools/summarise_profiles.py
# condenses it into profiles/.
TAG=${1:-r2f}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/gpu_$TAG.txt
for ks in "pv big" "wind big" "heat big" "pv small" "wind small" "spmm big"; do
  set -- $ks
  kn=k_fused_reduce; [ $1 = heat ] && kn=k_heat
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$kn -s 3 -c 1 \
    -o gpurun_out/prof_$1_$2_$TAG -f python tools/prof_pv.py $1 $2 2 > gpurun_out/ncu_$1_$2_$TAG.log 2>&1
done
# launch list of the bench command (our kernels only: inside the timed region a step is one memset + one
# kernel per resident part; torch's generators run outside it)
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --target-processes application-only \
  -k regex:k_ -c 80 --csv --log-file gpurun_out/launches_bench_$TAG.csv \
  python bench.py --steps 2 --warmup 3 --no-extra > gpurun_out/bench_under_ncu_$TAG.log 2>&1
ls -la gpurun_out | tail -12
