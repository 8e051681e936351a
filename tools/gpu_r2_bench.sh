#!/bin/bash
# GPU pass: parity suite, the north-star bench (our arm; with the staged-reduce variant beside it).
TAG=${1:-r2b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_full_$TAG.log 2>&1
cp gpurun_out/parity_errors.json gpurun_out/parity_errors_$TAG.json 2>/dev/null
timeout 1200 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "rc=$?" >> gpurun_out/bench_$TAG.err
[ "$2" = nostaged ] || ATL_VARIANT=2 timeout 900 python bench.py --no-extra --steps 5 --warmup 3 > gpurun_out/bench_staged_$TAG.json 2> gpurun_out/bench_staged_$TAG.err
grep -E "passed|failed" gpurun_out/pytest_full_$TAG.log | tail -2; grep -E "^FAILED|^ERROR" gpurun_out/pytest_full_$TAG.log | head
head -c 3000 gpurun_out/bench_$TAG.json; echo; tail -3 gpurun_out/bench_$TAG.err; head -c 1500 gpurun_out/bench_staged_$TAG.json
