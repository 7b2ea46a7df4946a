#!/bin/bash
# round 2, call 3: GPU suite (device set-up, permute policies, TMA G2P parity), A/B of the TMA-fed persistent G2P kernel
cd "$(dirname "$0")/.."
TAG=${1:-r2c}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -5 gpurun_out/pytest_gpu_$TAG.log
KML_G2P_TMA=1 timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_particle_order.py tests/test_large_block.py -m gpu -q -k "c5 or e_block or p_block or large or shuffled or parity_size" --timeout 600 > gpurun_out/pytest_tma_$TAG.log 2>&1; tail -5 gpurun_out/pytest_tma_$TAG.log
run() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), 'setup', d['config']['setup_s'], {k:v['ms'] for k,v in d['roofline']['per_stage'].items()}, d['clocks']['sm_mhz'], 'flags', d['error_flags'])
except Exception as e: print('FAILED', e)"; }
{
run KML_X=0
run KML_G2P_TMA=1
run KML_G2P_TMA=1 KML_G2P_THREADS=128
run KML_G2P_TMA=3 KML_G2P_THREADS=128
run KML_G2P_TMA=1 KML_SEGLEN_G2P=16
run KML_G2P_TMA=1 KML_SEGLEN_G2P=24 KML_G2P_THREADS=128
run KML_PERMUTE_FRAC=0.03
} > gpurun_out/ab_$TAG.log 2>&1
cat gpurun_out/ab_$TAG.log
