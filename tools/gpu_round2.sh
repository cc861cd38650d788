#!/bin/bash
# parity tests + A/B timing + ncu full captures of selected kernels on Au20
set -u
TAG=${1:-r}; shift || true
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -4 gpurun_out/${TAG}_pytest.log
timeout 300 python tools/ab_kernels.py "$@" 2>&1 | tail -20
bash tools/ncu_capture.sh cfg3 gpurun_out/${TAG}_ncu 3 k_fastT k_link k_fallbackT > /dev/null 2>&1
bash tools/ncu_capture.sh cfg3 gpurun_out/${TAG}_ncu 24 k_type1S > /dev/null 2>&1
ls gpurun_out/${TAG}_ncu
