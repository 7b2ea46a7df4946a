#!/bin/bash
# round 2, call 23 (1 GPU): P2G / re-projection segment length at 100 M particles
cd "$(dirname "$0")/.."
TAG=${1:-r2w}
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']; print(round(d['ms_per_step'],3), {k:v['ms'] for k,v in r['per_stage'].items()}, d['clocks']['sm_mhz'], 'permutes', r['physical_permutes_in_timed_region_rank0'], 'flags', d['error_flags'])
except Exception as e: print('FAILED', e)"; }
{
run KML_SEGLEN_P2G=32
run KML_SEGLEN_P2G=64
run KML_SEGLEN_P2G=96
} > gpurun_out/ab_$TAG.log 2>&1
cat gpurun_out/ab_$TAG.log
