#!/bin/bash
# round 2, GPU call 11: k_shift2 (flattened pass 1), k_chi2, compact F default: parity, A/B, bench
set -u
D=gpurun_out/r2c11; mkdir -p $D
( timeout 1200 python -m pytest tests -m gpu -q ) > $D/pytest_gpu.log 2>&1
tail -12 $D/pytest_gpu.log
rm -f gpurun_out/ab_kernels.jsonl
for wl in cfg3 cfg5_60; do
  timeout 300 python tools/ab_kernels.py $wl LIBECP_B200_SHIFT=two,2 LIBECP_B200_CHI=lanes,2 2>&1 | tail -4 | cut -c1-300
done
cp gpurun_out/ab_kernels.jsonl $D/
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu > $D/bench.json 2>> $D/bench.err
python tools/bench_brief.py < $D/bench.json
tail -3 $D/bench.err
