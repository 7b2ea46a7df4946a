#!/bin/bash
# round 2, call 17 (2 GPUs): halo sums over NVLink peer memory (CUDA IPC) - the 2-slab tests, then N = 2 with both halo paths
cd "$(dirname "$0")/.."
TAG=${1:-r2q}
mkdir -p gpurun_out; rm -f gpurun_out/slab_results.log
KML_DEBUG=1 python -m pytest tests/test_slab.py -m gpu -q --timeout 600 -k "two_slabs or second_solid" > gpurun_out/pytest_slab_$TAG.log 2>&1; tail -4 gpurun_out/pytest_slab_$TAG.log; grep -E "SLAB-FAIL|rror" gpurun_out/pytest_slab_$TAG.log | head -10 | cut -c1-500
cut -c1-200 gpurun_out/slab_results.log
runN() { n=$1; shift; echo "== N=$n $*"; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $n --steps 20 --warmup 5 --no-e2e --cells 124 250 202 2> gpurun_out/n${n}_stderr_$TAG.log > gpurun_out/bench_n${n}_$TAG.log; grep "halo exchange" gpurun_out/n${n}_stderr_$TAG.log | head -2; grep '^{' gpurun_out/bench_n${n}_$TAG.log | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']; print(round(d['ms_per_step'],3), round(d['value']/1e9,3), 'G p-s/s', {k:v['ms'] for k,v in r['per_stage'].items()}, 'comm', r['comm_ms'], 'parity', d.get('parity_ok'))
    for k,v in r.get('per_rank_stage_ms',{}).items(): print('   ', k, v)
except Exception as e: print('FAILED', e)"; }
{
runN 2 KML_DEBUG=1
runN 2 KML_DEBUG=1 KML_HALO=nccl KML_NOCHECK=1
} > gpurun_out/ab_$TAG.log 2>&1
cat gpurun_out/ab_$TAG.log
