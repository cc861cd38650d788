#!/bin/bash
# round 2, GPU call 29: link triples-per-block knob; bench with the 6 M default
set -u
D=gpurun_out/r2c29; mkdir -p $D
rm -f gpurun_out/ab_kernels.jsonl
for wl in cfg3 cfg5_60; do
  timeout 400 python tools/ab_kernels.py $wl LIBECP_B200_LINKTPB=-,4,8,32 2>&1 | tail -4 | sed 's/.*"env": //' | cut -c1-200
done
cp gpurun_out/ab_kernels.jsonl $D/
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu > $D/bench.json 2>> $D/bench.err
python tools/bench_brief.py < $D/bench.json | head -4
( timeout 900 python -m pytest tests -m gpu -q -x ) > $D/pytest_gpu.log 2>&1
tail -2 $D/pytest_gpu.log
