#!/bin/bash
# round 2, GPU call 6: k_fallbackW parity + A/B, full GPU tier, bench line
set -u
D=gpurun_out/r2c6; mkdir -p $D
( timeout 1200 python -m pytest tests -m gpu -q ) > $D/pytest_gpu.log 2>&1
tail -12 $D/pytest_gpu.log
rm -f gpurun_out/ab_kernels.jsonl
for wl in cfg3 cfg5_60; do
  timeout 300 python tools/ab_kernels.py $wl LIBECP_B200_FB=group,warp 2>&1 | tail -2 | cut -c1-330
  timeout 300 python tools/ab_kernels.py $wl LIBECP_B200_FBMINB=4 2>&1 | tail -1 | cut -c1-330
  timeout 300 python tools/ab_kernels.py $wl LIBECP_B200_FBOCC=2 2>&1 | tail -1 | cut -c1-330
done
cp gpurun_out/ab_kernels.jsonl $D/
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu > $D/bench.json 2> $D/bench.err
python tools/bench_brief.py < $D/bench.json
tail -3 $D/bench.err
