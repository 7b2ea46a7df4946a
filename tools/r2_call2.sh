#!/bin/bash
# round 2, call 2: GPU suite, A/B of the physical permute and of the segment lengths at full size (plastic regime), ncu capture
cd "$(dirname "$0")/.."
TAG=${1:-r2b}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -5 gpurun_out/pytest_gpu_$TAG.log
run() { echo "== $*"; env "$@" python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), 'setup', d['config']['setup_s'], {k:v['ms'] for k,v in d['roofline']['per_stage'].items()}, d['clocks']['sm_mhz'])
except Exception as e: print('FAILED', e)"; }
{
run KML_X=0
run KML_PERMUTE_FRAC=-1
run KML_PERMUTE_FRAC=0.02
run KML_PERMUTE_FRAC=0.10
run KML_SEGLEN_G2P=16
run KML_SEGLEN_G2P=48
run KML_SEGLEN_G2P=64 KML_GATHER_THREADS=128
run KML_GATHER_THREADS=128
run KML_SEGLEN_P2G=16
run KML_SEGLEN_P2G=64
run KML_SEGLEN_STRESS=16
run KML_SEGLEN_STRESS=32
} > gpurun_out/ab_$TAG.log 2>&1
cat gpurun_out/ab_$TAG.log
# capture the 4 stage kernels of step 42 (40 warm-up + 2 timed): 4 matching launches per step
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_p2g_cell3|k_g2p_cell|k_stress_cell" --launch-skip 164 --launch-count 4 -o gpurun_out/prof_$TAG -f \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log | cut -c1-300
ls -la gpurun_out/prof_$TAG.ncu-rep
