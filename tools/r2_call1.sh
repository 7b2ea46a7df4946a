#!/bin/bash
# round 2, call 1: GPU suite (new parity cases), default bench with the in-region stage timing, full-size ncu capture in the plastic regime
cd "$(dirname "$0")/.."
TAG=${1:-r2a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/smi_$TAG.txt 2>&1
python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -5 gpurun_out/pytest_gpu_$TAG.log
python bench.py > gpurun_out/bench_full_$TAG.log 2>&1; tail -1 gpurun_out/bench_full_$TAG.log
# 43 steps before the captured one: 4 matching launches per step (p2g full, g2p, p2g momentum, stress)
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_p2g_cell3|k_g2p_cell|k_stress_cell" --launch-skip 172 --launch-count 4 -o gpurun_out/prof_$TAG -f \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log | cut -c1-400
ls -la gpurun_out/prof_$TAG.ncu-rep
