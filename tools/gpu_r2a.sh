#!/bin/bash
# Round-2 first GPU pass: box probe, parity suite (all failures, no -x), kernel A/B of the staged
# reduce against the round-1 kernel, the north-star bench.  Everything lands in gpurun_out/.
TAG=${1:-r2a}
mkdir -p gpurun_out
{ nvidia-smi --query-gpu=name,memory.total,memory.used,clocks.max.sm,clocks.sm,power.limit --format=csv
  nproc; free -g | head -2; lscpu | grep -i "numa\|model name\|socket\|thread"; nvidia-smi topo -m 2>/dev/null | head -20
  cat /sys/fs/cgroup/memory.max 2>/dev/null; ulimit -l; } > gpurun_out/box_$TAG.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/pytest_$TAG.log
cp gpurun_out/parity_errors.json gpurun_out/parity_errors_$TAG.json 2>/dev/null
for v in 0 1 2; do for k in wind pv heat; do for s in small big; do
  ATL_VARIANT=$v timeout 120 python tools/prof_pv.py $k $s 7
done; done; done > gpurun_out/prof_$TAG.jsonl 2>gpurun_out/prof_$TAG.err
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -5 gpurun_out/pytest_$TAG.log; cat gpurun_out/prof_$TAG.jsonl; head -c 3000 gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
