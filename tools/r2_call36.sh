#!/bin/bash
# round 2, GPU call 36: GPU tier + bench with the streamed sparse download; Bessel recurrence 2-FMA (default) vs 3-op; one rank of 8
set -u
D=gpurun_out/r2c36; mkdir -p $D
( timeout 900 python -m pytest tests -m gpu -q -x ) > $D/pytest_gpu.log 2>&1
tail -5 $D/pytest_gpu.log
rm -f gpurun_out/ab_kernels.jsonl
for wl in cfg3 cfg5_60; do
  timeout 300 python tools/ab_kernels.py $wl LIBECP_B200_X=a,b 2>&1 | tail -2 | cut -c1-420
  LIBECP_B200_SO=$PWD/libecp_b200/lib/libecp_b200_b3op.so timeout 300 python tools/ab_kernels.py $wl LIBECP_B200_X=3op,3op 2>&1 | tail -2 | cut -c1-420
done
cp gpurun_out/ab_kernels.jsonl $D/ab_bessel2fma.jsonl
timeout 500 python bench.py --steps 5 --warmup 3 --no-cpu > $D/bench.json 2>> $D/bench.err
python tools/bench_brief.py < $D/bench.json | head -4
timeout 300 python tools/shard_profile.py 8 0 3 > $D/shard8.jsonl 2>&1
cat $D/shard8.jsonl | cut -c1-900
