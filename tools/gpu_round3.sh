#!/bin/bash
# quick: small-config parity diag + A/B timing
set -u
mkdir -p gpurun_out
rm -f gpurun_out/diag.txt
for v in "$@"; do
  echo "=== $v"
  env $v timeout 300 python tools/gpu_diag.py cfg2 au2 au4 cfg4a 2>&1 | grep -E "^==|VIOL|MISMATCH|worst|max" | cut -c1-250
done
