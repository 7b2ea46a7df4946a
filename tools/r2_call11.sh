#!/bin/bash
# round 2, call 11 (1 GPU): DFMA with three register operands (microbenchmark), bulk-copied G2P tile (KML_G2P_TMA=4): parity + A/B
cd "$(dirname "$0")/.."
TAG=${1:-r2k}
mkdir -p gpurun_out
./karamelo_b200/lib/fp64_peak > gpurun_out/fp64_peak_$TAG.txt 2>&1; tail -6 gpurun_out/fp64_peak_$TAG.txt
KML_G2P_TMA=4 python -m pytest tests/test_parity_gpu.py tests/test_properties_gpu.py tests/test_particle_order.py tests/test_large_block.py -m gpu -q --timeout 900 -x > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -3 gpurun_out/pytest_gpu_$TAG.log
run() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']; print(round(d['ms_per_step'],3), {k:v['ms'] for k,v in r['per_stage'].items()}, d['clocks']['sm_mhz'], 'permutes', r['physical_permutes_in_timed_region_rank0'], 'flags', d['error_flags'])
except Exception as e: print('FAILED', e)"; }
{
run KML_G2P_TMA=4
run KML_G2P_TMA=0
run KML_G2P_TMA=4 KML_SEGLEN_G2P=24
run KML_G2P_TMA=4 KML_SEGLEN_G2P=48
run KML_G2P_TMA=4 KML_G2P_THREADS=128
} > gpurun_out/ab_$TAG.log 2>&1
cat gpurun_out/ab_$TAG.log
