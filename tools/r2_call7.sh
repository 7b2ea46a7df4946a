#!/bin/bash
# round 2, call 7 (8 GPUs): slab tests at 2 / 4 / 8 ranks, strong-scaling bench at N = 8 and 4 with --check
cd "$(dirname "$0")/.."
TAG=${1:-r2g}
mkdir -p gpurun_out
python -m pytest tests/test_slab.py -m gpu -q --timeout 900 > gpurun_out/pytest_slab_$TAG.log 2>&1; tail -6 gpurun_out/pytest_slab_$TAG.log
runN() { n=$1; shift; echo "== N=$n $*"; env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $n --steps 20 --warmup 5 --no-e2e 2> gpurun_out/n${n}_stderr_$TAG.log > gpurun_out/bench_n${n}_$TAG.log; grep '^{' gpurun_out/bench_n${n}_$TAG.log | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']; print(round(d['ms_per_step'],3), round(d['value']/1e9,3), 'G p-s/s', {k:v['ms'] for k,v in r['per_stage'].items()}, 'comm', r['comm_ms'], 'sum0', r['stage_sum_ms_rank0'], 'wall0', r['wall_ms_per_step_rank0'], 'permutes', r['physical_permutes_in_timed_region_rank0'], 'parity', d.get('parity_ok'), d.get('parity',{}).get('worst_rel'), d['config'].get('particles_per_rank'), d['clocks'])
except Exception as e: print('FAILED', e)"; }
{
runN 8 KML_X=0
runN 4 KML_X=0
runN 8 KML_PERMUTE_FRAC=-1 KML_NOCHECK=1
} > gpurun_out/ab_$TAG.log 2>&1
cat gpurun_out/ab_$TAG.log
tail -3 gpurun_out/n8_stderr_$TAG.log
