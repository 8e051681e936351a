#!/bin/bash
# parity suite with the product library, then an A/B pass (tools/gpu_r2_ab.sh)
TAG=${1:-r2z}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_full_$TAG.log 2>&1
cp gpurun_out/parity_errors.json gpurun_out/parity_errors_$TAG.json 2>/dev/null
grep -E "passed|failed" gpurun_out/pytest_full_$TAG.log | tail -3; grep -E "^FAILED|^ERROR" gpurun_out/pytest_full_$TAG.log | head
bash tools/gpu_r2_ab.sh $TAG "$2" "$3"
