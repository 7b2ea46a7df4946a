#!/bin/bash
# round 2, call 20 (1 GPU): final tree - smoke, the GPU suite, the default bench
cd "$(dirname "$0")/.."
TAG=${1:-r2t}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; tail -1 gpurun_out/smoke_$TAG.log | cut -c1-200
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -3 gpurun_out/pytest_gpu_$TAG.log
python bench.py > gpurun_out/bench_full_$TAG.log 2> gpurun_out/bench_full_$TAG.err; tail -1 gpurun_out/bench_full_$TAG.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; print(round(d['ms_per_step'],3), round(d['value']/1e9,3), {k:v['ms'] for k,v in r['per_stage'].items()}, d['clocks'], 'e2e', d['e2e']['value'], 'flags', d['error_flags'])"
