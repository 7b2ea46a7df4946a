#!/bin/bash
# Experiment: SoA component-stride padding at full size (G2P reads 6 and writes 6 streams a fixed stride apart).
cd "$(dirname "$0")/.."
run() { echo "== $*"; env "${@:4}" python bench.py --steps 5 --warmup 3 --cells $1 $2 $3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); n=d['config']['particles']; print(n, round(d['ms_per_step'],3), {k:v['ms'] for k,v in d['roofline']['per_stage'].items()})"; }
run 248 250 202 KML_X=0
run 248 250 202 KML_CAP_PAD=2080
run 248 250 202 KML_CAP_PAD=16416
run 248 250 202 KML_CAP_PAD=262176
run 248 250 202 KML_X=0
