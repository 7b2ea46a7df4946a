#!/bin/bash
# A/B measurement helper run on the GPU box (gpurun -- bash tools/ab.sh [env settings...]): GPU parity suite, then timing at 96^3 cells.
cd "$(dirname "$0")/.."
python -m pytest tests/test_parity_gpu.py -x -q > gpurun_out/pytest_ab.log 2>&1; tail -3 gpurun_out/pytest_ab.log
run() { echo "== $*"; env "$@" python bench.py --steps 5 --warmup 3 --cells 96 96 96 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), {k:v['ms'] for k,v in d['roofline']['per_stage'].items()})"; }
run KML_DEFAULT=1
run KML_GATHER_THREADS=128
