#!/bin/bash
# wind kernel vs slab length (432 steps ... the full year), SM clock / power sampled beside it
TAG=${1:-r2l}
mkdir -p gpurun_out
nvidia-smi --query-gpu=timestamp,clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_throttle_reasons.active --format=csv -lms 200 > gpurun_out/smi_$TAG.csv &
SMI=$!
for nt in 432 2190 8760; do for lib in default build_variants/*.so; do
  if [ $lib = default ]; then ATL_NT=$nt timeout 300 python tools/prof_pv.py wind big 15
  else ATL_NT=$nt ATL_LIB_PATH=$lib timeout 300 python tools/prof_pv.py wind big 15; fi
done; done > gpurun_out/prof_$TAG.jsonl 2>gpurun_out/prof_$TAG.err
ATL_NT=8760 ATL_TB=64 timeout 300 python tools/prof_pv.py wind big 15 >> gpurun_out/prof_$TAG.jsonl 2>>gpurun_out/prof_$TAG.err
ATL_NT=8760 ATL_TB=1024 timeout 300 python tools/prof_pv.py wind big 15 >> gpurun_out/prof_$TAG.jsonl 2>>gpurun_out/prof_$TAG.err
kill $SMI
cut -c1-260 gpurun_out/prof_$TAG.jsonl; tail -3 gpurun_out/prof_$TAG.err
awk -F, 'NR>1{print $2}' gpurun_out/smi_$TAG.csv | sort | uniq -c | sort -rn | head -8
