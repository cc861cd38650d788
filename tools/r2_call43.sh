#!/bin/bash
# round 2, GPU call 43: k_type1A wave + k_type1S survivors - bit-identity test, A/B vs legacy, GPU tier, bench
set -u
D=gpurun_out/r2c43; mkdir -p $D
( timeout 600 python -m pytest tests -m gpu -q -x -k "type1_wave or getintegrals or config3" ) > $D/pytest_t1.log 2>&1
tail -12 $D/pytest_t1.log
rm -f gpurun_out/ab_kernels.jsonl
for wl in cfg3 cfg5_60; do
  timeout 300 python tools/ab_kernels.py $wl LIBECP_B200_T1=-,legacy 2>&1 | tail -2 | cut -c1-330
done
cp gpurun_out/ab_kernels.jsonl $D/ab_t1wave.jsonl
( timeout 900 python -m pytest tests -m gpu -q -x ) > $D/pytest_gpu.log 2>&1
tail -3 $D/pytest_gpu.log
timeout 500 python bench.py --steps 5 --warmup 3 --no-cpu > $D/bench.json 2>> $D/bench.err
python tools/bench_brief.py < $D/bench.json | head -4
