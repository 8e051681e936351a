#!/bin/bash
# GPU pass: the north-star bench (both arms), nothing else.
TAG=${1:-r2b}
mkdir -p gpurun_out
free -g | head -2 > gpurun_out/mem_before_$TAG.txt
timeout 1200 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "rc=$?" >> gpurun_out/bench_$TAG.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
head -c 6000 gpurun_out/bench_$TAG.json; echo; tail -5 gpurun_out/bench_$TAG.err; head -c 1500 gpurun_out/bench_ref_$TAG.json
