#!/bin/bash
# A/B of the tuning knobs at full size (stage times in ms)
cd "$(dirname "$0")/.."
run() { echo "== $*"; env "${@:4}" python bench.py --steps 5 --warmup 3 --cells $1 $2 $3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['particles'], round(d['ms_per_step'],3), {k:v['ms'] for k,v in d['roofline']['per_stage'].items()})"; }
run 248 250 202 KML_V2G_NB=1
run 248 250 202 KML_V2G_NB=4
run 248 250 202 KML_SEGLEN=48
run 248 250 202 KML_SEGLEN=24
