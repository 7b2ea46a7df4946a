#!/bin/bash
# round 2, call 5 (2 GPUs): drop-in over CUDA, diagnosis of the re-bin / host gaps at N=2
cd "$(dirname "$0")/.."
TAG=${1:-r2e}
mkdir -p gpurun_out
python -m pytest tests/test_dropin.py tests/test_device_setup.py tests/test_forces_gpu.py -m gpu -q --timeout 900 > gpurun_out/pytest_dropin_$TAG.log 2>&1; tail -4 gpurun_out/pytest_dropin_$TAG.log
run2() { echo "== N=2 $*"; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 --cells 62 250 202 --no-e2e --no-check 2> gpurun_out/n2_stderr_$TAG.log | grep '^{' | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']; print(round(d['ms_per_step'],3), {k:v['ms'] for k,v in r['per_stage'].items()}, 'comm', r['comm_ms'], 'sum0', r['stage_sum_ms_rank0'], 'wall0', r['wall_ms_per_step_rank0'], 'host', r['host_ms_in_calls_rank0'], 'permutes', r['physical_permutes_in_timed_region_rank0'])
except Exception as e: print('FAILED', e)"; grep "kml rank" gpurun_out/n2_stderr_$TAG.log | tail -12; }
run1() { echo "== N=1 $*"; env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --cells 31 250 202 --no-e2e --no-cpu-baseline 2> gpurun_out/n1_stderr_$TAG.log | grep '^{' | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']; print(round(d['ms_per_step'],3), {k:v['ms'] for k,v in r['per_stage'].items()}, 'wall0', r['wall_ms_per_step_rank0'], 'host', r['host_ms_in_calls_rank0'], 'permutes', r['physical_permutes_in_timed_region_rank0'])
except Exception as e: print('FAILED', e)"; grep "kml rank" gpurun_out/n1_stderr_$TAG.log | tail -8; }
{
run1 KML_DEBUG=1
run2 KML_DEBUG=1
run2 KML_DEBUG=1 KML_PERMUTE_FRAC=-1
run2 KML_DEBUG=1 KML_PERMUTE_MIN_STEPS=16
} > gpurun_out/ab_$TAG.log 2>&1
cat gpurun_out/ab_$TAG.log
