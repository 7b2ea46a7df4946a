#!/bin/bash
# round 2, call 10 (1 GPU): software-pipelined record loads in P2G / re-projection: parity, then A/B at 100 M particles
cd "$(dirname "$0")/.."
TAG=${1:-r2j}
mkdir -p gpurun_out
python -m pytest tests/test_parity_gpu.py tests/test_properties_gpu.py tests/test_particle_order.py tests/test_large_block.py -m gpu -q --timeout 900 -x > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -3 gpurun_out/pytest_gpu_$TAG.log
run() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']; print(round(d['ms_per_step'],3), {k:v['ms'] for k,v in r['per_stage'].items()}, d['clocks']['sm_mhz'], 'permutes', r['physical_permutes_in_timed_region_rank0'], 'flags', d['error_flags'])
except Exception as e: print('FAILED', e)"; }
{
run KML_P2G_PIPE=1
run KML_P2G_PIPE=0
run KML_P2G_PIPE=1 KML_V2G_NB=1
run KML_P2G_PIPE=1 KML_SEGLEN_P2G=64
run KML_P2G_PIPE=1 KML_P2G_NB=2
} > gpurun_out/ab_$TAG.log 2>&1
cat gpurun_out/ab_$TAG.log
