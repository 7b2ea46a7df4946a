#!/bin/bash
# round 2, call 9 (1 GPU): mixed-issue FP64 microbenchmark, membership replay (GPU must now follow the REFERENCE branch), parity suite
cd "$(dirname "$0")/.."
TAG=${1:-r2i}
mkdir -p gpurun_out
./karamelo_b200/lib/fp64_peak > gpurun_out/fp64_peak_$TAG.txt 2>&1; cat gpurun_out/fp64_peak_$TAG.txt
python -m pytest tests/test_weight_zero_skip.py -m gpu -q -s --timeout 900 2>&1 | grep -E "either|passed|failed" | cut -c1-900 | tee gpurun_out/membership_$TAG.log
python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -4 gpurun_out/pytest_gpu_$TAG.log
