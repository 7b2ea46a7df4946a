#!/bin/bash
cd "$(dirname "$0")/.."
run() { echo "== $*"; env "$@" python bench.py --steps 5 --warmup 3 --cells 96 96 96 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), {k:v['ms'] for k,v in d['roofline']['per_stage'].items()})"; }
run KML_DEFAULT=1
run KML_DEBUG_NORED=1
