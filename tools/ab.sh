#!/bin/bash
# A/B measurement helper run on the GPU box (gpurun -- bash tools/ab.sh): parity subset, then kernel variants at 96^3 cells.
cd "$(dirname "$0")/.."
python -m pytest tests/test_parity_gpu.py -x -q -k "c5 or c2_taylor_cubic" > gpurun_out/pytest_ab.log 2>&1; tail -3 gpurun_out/pytest_ab.log
run() { echo "== $*"; env "$@" python bench.py --steps 5 --warmup 3 --cells 96 96 96 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), {k:v['ms'] for k,v in d['roofline']['per_stage'].items()})"; }
run KML_DEFAULT=1
run KML_P2G_PIPE=0
run KML_P2G_PIPE=0 KML_V2G_NB=2
run KML_P2G_PIPE=1 KML_V2G_NB=2
run KML_P2G_PIPE=1 KML_V2G_NB=1 KML_P2G_NB=2
run KML_P2G_PIPE=0 KML_V2G_NB=1
