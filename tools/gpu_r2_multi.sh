#!/bin/bash
# Multi-GPU pass (gpurun --gpus N): hardware check of the time-sharded and the single-process multi-GPU
# paths, the strong-scaling bench at N = all visible GPUs, the single-process end-to-end scaling.
TAG=${1:-r2m}
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
{ nvidia-smi topo -m | head -14; free -g | head -2; cat /sys/fs/cgroup/memory.max /sys/fs/cgroup/cpu.max 2>/dev/null; } > gpurun_out/box_${TAG}_n$N.txt 2>&1
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_multi_${TAG}_n$N.log 2>&1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err
for s in small big; do timeout 120 python tools/prof_pv.py wind $s 7; timeout 120 python tools/prof_pv.py heat $s 7; done > gpurun_out/prof_${TAG}_n$N.jsonl 2>/dev/null
timeout 900 python tools/e2e_multi.py --kind pv > gpurun_out/e2e_multi_${TAG}_n$N.json 2> gpurun_out/e2e_multi_${TAG}_n$N.err
tail -3 gpurun_out/pytest_multi_${TAG}_n$N.log; head -c 2500 gpurun_out/bench_${TAG}_n$N.json; echo; tail -3 gpurun_out/bench_${TAG}_n$N.err
cat gpurun_out/prof_${TAG}_n$N.jsonl; cat gpurun_out/e2e_multi_${TAG}_n$N.json; tail -3 gpurun_out/e2e_multi_${TAG}_n$N.err
