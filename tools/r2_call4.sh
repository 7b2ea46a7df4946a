#!/bin/bash
# round 2, call 4 (2 GPUs): full GPU suite incl. the two-slab tests, N=2 bench at 12.5 M particles per GPU (the per-GPU load of N=8 at 100 M) with the stage breakdown and --check
cd "$(dirname "$0")/.."
TAG=${1:-r2d}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -8 gpurun_out/pytest_gpu_$TAG.log
run2() { echo "== N=2 $*"; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 --cells 62 250 202 --no-e2e 2>&1 | grep '^{' | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), {k:v['ms'] for k,v in d['roofline']['per_stage'].items()}, 'comm', d['roofline']['comm_ms'], 'sum', d['roofline']['stage_sum_ms_rank0'], 'parity', d.get('parity_ok'), d.get('parity',{}).get('worst_rel'), d['config'].get('particles_per_rank'))
except Exception as e: print('FAILED', e)"; }
run1() { echo "== N=1 $*"; env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --cells 31 250 202 --no-e2e --no-cpu-baseline 2>&1 | grep '^{' | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), {k:v['ms'] for k,v in d['roofline']['per_stage'].items()}, 'comm', d['roofline']['comm_ms'])
except Exception as e: print('FAILED', e)"; }
{
run1 KML_X=0
run2 KML_X=0
run2 KML_PERMUTE_FRAC=-1
run2 NCCL_DEBUG=WARN KML_PERMUTE_FRAC=0.02
} > gpurun_out/ab_$TAG.log 2>&1
cat gpurun_out/ab_$TAG.log
