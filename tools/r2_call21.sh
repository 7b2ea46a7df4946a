#!/bin/bash
# round 2, call 21 (2 GPUs): migration scratch growth path (KML_MIG_CAP=16 forces it in the drifting 2-slab tests)
cd "$(dirname "$0")/.."
TAG=${1:-r2u}
mkdir -p gpurun_out; rm -f gpurun_out/slab_results.log
KML_MIG_CAP=16 python -m pytest tests/test_slab.py -m gpu -q --timeout 600 -k "two_slabs_match or (second_solid and 2-12)" > gpurun_out/pytest_slab_$TAG.log 2>&1; tail -3 gpurun_out/pytest_slab_$TAG.log; grep -E "SLAB-FAIL" gpurun_out/pytest_slab_$TAG.log | head -5 | cut -c1-600
cut -c1-220 gpurun_out/slab_results.log
