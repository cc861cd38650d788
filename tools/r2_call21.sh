#!/bin/bash
# round 2, GPU call 21: k_Ftab2 occupancy bound A/B (serial per-kernel times on cfg3 / cfg5_60 / bench)
set -u
D=gpurun_out/r2c21; mkdir -p $D
rm -f gpurun_out/ab_kernels.jsonl
for wl in cfg3 cfg5_60; do
  timeout 300 python tools/ab_kernels.py $wl 2>&1 | tail -1 | cut -c1-330
done
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu > $D/bench.json 2>> $D/bench.err
python tools/bench_brief.py < $D/bench.json | head -3
