#!/bin/bash
# Round measurement on the GPU box: GPU test suite, default bench, ncu launch list of the bench command, ncu --set full of the stage kernels.
cd "$(dirname "$0")/.."
TAG=${1:-r1}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -3 gpurun_out/pytest_gpu_$TAG.log
python bench.py > gpurun_out/bench_full_$TAG.log 2>&1; tail -1 gpurun_out/bench_full_$TAG.log
python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_ref_$TAG.log 2>&1; tail -1 gpurun_out/bench_ref_$TAG.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches_$TAG.log 2>&1; tail -1 gpurun_out/ncu_launches_$TAG.log | cut -c1-300
bash tools/ncu_full.sh $TAG "k_p2g_cell3|k_g2p_cell|k_stress_cell"
