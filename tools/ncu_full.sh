#!/bin/bash
# one ncu --set full capture of the step's top kernels (run on the GPU box): tools/ncu_full.sh <tag> [kernel regex]
cd "$(dirname "$0")/.."
TAG=${1:-r1}; RX=${2:-"k_p2g_cell3|k_g2p_cell|k_stress_cell"}
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$RX" --launch-skip 8 --launch-count 4 -o gpurun_out/prof_$TAG -f \
  python bench.py --steps 2 --warmup 3 --cells 96 96 96 --no-e2e --no-cpu-baseline > gpurun_out/ncu_$TAG.log 2>&1
tail -3 gpurun_out/ncu_$TAG.log
