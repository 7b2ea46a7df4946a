#!/bin/bash
# round 2, call 26 (1 GPU): last knobs - four node columns per lane in the re-projection, stress segment lengths around 24
cd "$(dirname "$0")/.."
TAG=${1:-r2y}
mkdir -p gpurun_out
KML_V2G_NB=4 python -m pytest tests/test_parity_gpu.py tests/test_large_block.py -m gpu -q --timeout 600 -x -k "c5 or block or taylor_cubic or large" 2>&1 | tail -2
run() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']; print(round(d['ms_per_step'],3), {k:v['ms'] for k,v in r['per_stage'].items()}, d['clocks']['sm_mhz'], 'permutes', r['physical_permutes_in_timed_region_rank0'], 'flags', d['error_flags'])
except Exception as e: print('FAILED', e)"; }
{
run KML_X=0
run KML_V2G_NB=4
run KML_SEGLEN_STRESS=28
} > gpurun_out/ab_$TAG.log 2>&1
cat gpurun_out/ab_$TAG.log
