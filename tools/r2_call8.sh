#!/bin/bash
# round 2, call 8 (1 GPU): parity of the fourth-generation P2G (cp.async raw prefetch, 4 blocks per SM) + membership test, A/B at 100 M particles
cd "$(dirname "$0")/.."
TAG=${1:-r2h}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -6 gpurun_out/pytest_gpu_$TAG.log
python -m pytest tests/test_weight_zero_skip.py -m gpu -q -s --timeout 900 2>&1 | grep -E "either|passed|failed" | cut -c1-700
run() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']; print(round(d['ms_per_step'],3), {k:v['ms'] for k,v in r['per_stage'].items()}, d['clocks']['sm_mhz'], 'permutes', r['physical_permutes_in_timed_region_rank0'], 'flags', d['error_flags'])
except Exception as e: print('FAILED', e)"; }
{
run KML_P2G_V=4
run KML_P2G_V=3
run KML_P2G_V=4 KML_P2G_MINB=3
run KML_P2G_V=4 KML_V2G_NB=1
run KML_P2G_V=4 KML_SEGLEN_P2G=64
run KML_P2G_V=4 KML_SEGLEN_P2G=24
} > gpurun_out/ab_$TAG.log 2>&1
cat gpurun_out/ab_$TAG.log
