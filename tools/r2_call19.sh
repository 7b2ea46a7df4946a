#!/bin/bash
# round 2, call 19 (4 GPUs): the whole slab suite on the final tree (peer-memory halo), strong-scaling bench at N = 4 and N = 2 with the parity check
cd "$(dirname "$0")/.."
TAG=${1:-r2s}
mkdir -p gpurun_out; rm -f gpurun_out/slab_results.log
python -m pytest tests/test_slab.py -m gpu -q --timeout 900 > gpurun_out/pytest_slab_$TAG.log 2>&1; tail -3 gpurun_out/pytest_slab_$TAG.log; grep -E "SLAB-FAIL" gpurun_out/pytest_slab_$TAG.log | head -5 | cut -c1-500
runN() { n=$1; shift; echo "== N=$n $*"; env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $n --steps 20 --warmup 5 --no-e2e 2> gpurun_out/n${n}_stderr_$TAG.log > gpurun_out/bench_n${n}_$TAG.log; grep '^{' gpurun_out/bench_n${n}_$TAG.log | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']; print(round(d['ms_per_step'],3), round(d['value']/1e9,3), 'G p-s/s', {k:v['ms'] for k,v in r['per_stage'].items()}, 'comm', r['comm_ms'], 'parity', d.get('parity_ok'), d['clocks'])
    for k,v in r.get('per_rank_stage_ms',{}).items(): print('   ', k, v)
except Exception as e: print('FAILED', e)"; }
{
runN 4 KML_X=0
runN 2 KML_X=0
} > gpurun_out/ab_$TAG.log 2>&1
cat gpurun_out/ab_$TAG.log
