#!/bin/bash
# round 2, call 15 (1 GPU): closing set of the round on the final tree: smoke, full GPU suite, default bench (incl. e2e + cpu baseline), reference arm,
# ncu launch list and --set full capture at the full size
cd "$(dirname "$0")/.."
TAG=${1:-r2o}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; tail -1 gpurun_out/smoke_$TAG.log
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -3 gpurun_out/pytest_gpu_$TAG.log
python bench.py > gpurun_out/bench_full_$TAG.log 2> gpurun_out/bench_full_$TAG.err; tail -1 gpurun_out/bench_full_$TAG.log | cut -c1-1500
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_full20_$TAG.log 2>> gpurun_out/bench_full_$TAG.err; tail -1 gpurun_out/bench_full20_$TAG.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; print('20+5:', round(d['ms_per_step'],3), {k:(v['ms'],v['frac'],v.get('fp64_frac')) for k,v in r['per_stage'].items()}, d['clocks'], 'e2e', d['e2e']['value'])"
python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_ref_$TAG.log 2>&1; tail -1 gpurun_out/bench_ref_$TAG.log | cut -c1-400
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 440 -c 90 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches_$TAG.log 2>&1; tail -1 gpurun_out/ncu_launches_$TAG.log | cut -c1-200
# the 4 stage kernels of step 42 (40 warm-up + 2 timed): 4 matching launches per step
timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:k_p2g_cell3|k_g2p_cell|k_stress_cell" --launch-skip 164 --launch-count 4 -o gpurun_out/prof_$TAG -f \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_$TAG.log 2>&1
tail -1 gpurun_out/ncu_$TAG.log | cut -c1-200
ls -la gpurun_out/prof_$TAG.ncu-rep
