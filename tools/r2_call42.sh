#!/bin/bash
# round 2, GPU call 42: merged first chunk of k_type1S - Q bit-identical to the unmerged build? timing A/B; stats; GPU tier
set -u
D=gpurun_out/r2c42; mkdir -p $D
NM=$PWD/libecp_b200/lib/libecp_b200_t1nomerge.so
for so in "" $NM; do
  LIBECP_B200_SO=$so timeout 200 python tools/dump_inter.py Q 3000000 cfg3_4 cfg3_20 cfg4a cfg4b cfg5_40
done > $D/q_hash.out 2>&1
cat $D/q_hash.out
rm -f gpurun_out/ab_kernels.jsonl
for wl in cfg3 cfg5_60; do
  timeout 300 python tools/ab_kernels.py $wl LIBECP_B200_X=merge 2>&1 | tail -1 | cut -c1-330
  LIBECP_B200_SO=$NM timeout 300 python tools/ab_kernels.py $wl LIBECP_B200_X=nomerge 2>&1 | tail -1 | cut -c1-330
done
cp gpurun_out/ab_kernels.jsonl $D/ab_t1merge.jsonl
( timeout 900 python -m pytest tests -m gpu -q -x ) > $D/pytest_gpu.log 2>&1
tail -3 $D/pytest_gpu.log
timeout 500 python bench.py --steps 5 --warmup 3 --no-cpu > $D/bench.json 2>> $D/bench.err
python tools/bench_brief.py < $D/bench.json | head -4
