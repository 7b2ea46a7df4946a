#!/bin/bash
# defaults at 96^3 cells and at full size (stage times in ms)
cd "$(dirname "$0")/.."
run() { echo "== $*"; env "${@:4}" python bench.py --steps 5 --warmup 3 --cells $1 $2 $3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), {k:v['ms'] for k,v in d['roofline']['per_stage'].items() if k in ('p2g','g2p','v2g','stress')})"; }
run 96 96 96 KML_X=0
run 248 250 202 KML_X=0
run 248 250 202 KML_SEGLEN_G2P=24
