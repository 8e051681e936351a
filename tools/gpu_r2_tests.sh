#!/bin/bash
# GPU pass 1 of round 2: box probe + the GPU parity suite (all failures reported, no -x).
TAG=${1:-r2t}
mkdir -p gpurun_out
{ nvidia-smi --query-gpu=name,memory.total,memory.used,clocks.max.sm,clocks.sm,power.limit --format=csv
  nproc; python -c "import os;print('affinity',len(os.sched_getaffinity(0)))"; free -g | head -2
  lscpu | grep -i "numa\|model name\|socket\|thread"; nvidia-smi topo -m 2>/dev/null | head -20
  echo cgroup; cat /sys/fs/cgroup/memory.max /sys/fs/cgroup/cpu.max 2>/dev/null; ulimit -l; } > gpurun_out/box_$TAG.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_full_$TAG.log 2>&1
tail -80 gpurun_out/pytest_full_$TAG.log > gpurun_out/pytest_$TAG.log
cp gpurun_out/parity_errors.json gpurun_out/parity_errors_$TAG.json 2>/dev/null
cat gpurun_out/box_$TAG.txt | head -30; grep -E "passed|failed|error" gpurun_out/pytest_$TAG.log | tail -5
grep -E "^FAILED|^ERROR" gpurun_out/pytest_full_$TAG.log | head -40
