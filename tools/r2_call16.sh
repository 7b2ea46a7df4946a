#!/bin/bash
# round 2, call 16 (4 GPUs): second solid / rigid solid / delete_particles / force_nodes on decomposed grids
cd "$(dirname "$0")/.."
TAG=${1:-r2p}
mkdir -p gpurun_out; rm -f gpurun_out/slab_results.log
python -m pytest tests/test_slab.py -m gpu -q --timeout 900 -k "second_solid or delete or force" > gpurun_out/pytest_slab_$TAG.log 2>&1; tail -4 gpurun_out/pytest_slab_$TAG.log; grep -E "SLAB-FAIL|Error|error" gpurun_out/pytest_slab_$TAG.log | head -20 | cut -c1-600
cut -c1-250 gpurun_out/slab_results.log
