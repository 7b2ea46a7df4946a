#!/bin/bash
# round 2, call 6 (1 GPU): full GPU suite (multi-solid cell path, cost-based permute, drop-in), full-size bench incl. e2e + cpu baseline, small configs, ncu
cd "$(dirname "$0")/.."
TAG=${1:-r2f}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -6 gpurun_out/pytest_gpu_$TAG.log
KML_DEBUG=1 python bench.py > gpurun_out/bench_full_$TAG.log 2> gpurun_out/bench_full_$TAG.err; tail -1 gpurun_out/bench_full_$TAG.log; grep "kml rank" gpurun_out/bench_full_$TAG.err | tail -8
for c in c1 c2 c3 c4; do python bench.py --config $c --steps 200 > gpurun_out/bench_$c\_$TAG.log 2>&1; tail -1 gpurun_out/bench_$c\_$TAG.log | cut -c1-600; done
run() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']; print(round(d['ms_per_step'],3), {k:v['ms'] for k,v in r['per_stage'].items()}, d['clocks']['sm_mhz'], 'permutes', r['physical_permutes_in_timed_region_rank0'])
except Exception as e: print('FAILED', e)"; }
{
run KML_X=0
run KML_PERMUTE_EVERY=4
run KML_PERMUTE_EVERY=8
run KML_PERMUTE_EVERY=12
} > gpurun_out/ab_$TAG.log 2>&1
cat gpurun_out/ab_$TAG.log
