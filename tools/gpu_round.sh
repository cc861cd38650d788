#!/bin/bash
# One GPU round: parity tests, A/B kernel timings on Au20, the default bench line.  Outputs under gpurun_out/.
# usage: tools/gpu_round.sh <tag> [ab args...]
set -u
TAG=${1:-r}; shift || true
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -5 gpurun_out/${TAG}_pytest.log
timeout 300 python tools/ab_kernels.py "$@" 2>&1 | tail -20
( time timeout 600 python bench.py --steps 3 --warmup 3 ) > gpurun_out/${TAG}_bench.log 2>&1
tail -3 gpurun_out/${TAG}_bench.log | cut -c1-3000
