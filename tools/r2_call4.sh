#!/bin/bash
# round 2, GPU call 4: k_link4 + k_type1D: parity (GPU tier), A/B timings, config-5 digest with the two-build envelope
set -u
D=gpurun_out/r2c4; mkdir -p $D
( timeout 1200 python -m pytest tests -m gpu -q ) > $D/pytest_gpu.log 2>&1
tail -12 $D/pytest_gpu.log
rm -f gpurun_out/ab_kernels.jsonl
for wl in cfg3 cfg5_60; do
  timeout 200 python tools/ab_kernels.py $wl LIBECP_B200_LINK=global,4 LIBECP_B200_T1=level,dense 2>&1 | tail -4 | cut -c1-330
  timeout 200 python tools/ab_kernels.py $wl LIBECP_B200_LINKTPB=2,4,8,16 2>&1 | tail -4 | cut -c1-330
done
cp gpurun_out/ab_kernels.jsonl $D/
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu > $D/bench.json 2> $D/bench.err
python tools/bench_brief.py < $D/bench.json
tail -3 $D/bench.err
