#!/bin/bash
# Experiment: per-particle stage time against problem size, and SoA component-stride padding at full size.
cd "$(dirname "$0")/.."
run() { echo "== $*"; env "${@:4}" python bench.py --steps 5 --warmup 3 --cells $1 $2 $3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); n=d['config']['particles']; print(n, round(d['ms_per_step'],3), {k:(v['ms'], round(v['ms']/n*1e6,4)) for k,v in d['roofline']['per_stage'].items()})"; }
run 96 96 96 KML_X=0
run 160 160 160 KML_X=0
run 248 250 202 KML_X=0
run 248 250 202 KML_CAP_PAD=4128
run 248 250 202 KML_CAP_PAD=1048608
