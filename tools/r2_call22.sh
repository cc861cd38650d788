#!/bin/bash
# round 2, GPU call 22: packed fallback evaluation (k_fbw_eval2): parity, A/B, bench
set -u
D=gpurun_out/r2c22; mkdir -p $D
( timeout 1500 python -m pytest tests -m gpu -q ) > $D/pytest_gpu.log 2>&1
tail -6 $D/pytest_gpu.log
rm -f gpurun_out/ab_kernels.jsonl
for wl in cfg3 cfg5_60; do
  timeout 300 python tools/ab_kernels.py $wl LIBECP_B200_FBPACK=0,1 2>&1 | tail -2 | cut -c1-330
done
cp gpurun_out/ab_kernels.jsonl $D/
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu > $D/bench.json 2>> $D/bench.err
python tools/bench_brief.py < $D/bench.json | head -4
