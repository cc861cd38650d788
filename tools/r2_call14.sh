#!/bin/bash
# round 2, GPU call 14: k_type1Q + cheaper Bessel + fastA (warp-uniform class) parity and A/B
set -u
D=gpurun_out/r2c14; mkdir -p $D
( timeout 1500 python -m pytest tests -m gpu -q ) > $D/pytest_gpu.log 2>&1
tail -15 $D/pytest_gpu.log
rm -f gpurun_out/ab_kernels.jsonl
for wl in cfg3 cfg5_60; do
  timeout 300 python tools/ab_kernels.py $wl LIBECP_B200_T1=old,new LIBECP_B200_FAST=old,new 2>&1 | tail -4 | cut -c1-330
done
cp gpurun_out/ab_kernels.jsonl $D/
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu > $D/bench.json 2>> $D/bench.err
python tools/bench_brief.py < $D/bench.json
bash tools/ncu_capture.sh cfg3 $D/ncu_cfg3 4 'k_type1Q' 'k_fastA' > /dev/null 2>&1
python tools/ncu_keys.py "$D/ncu_cfg3/*.raw.csv" > $D/ncu_cfg3_keys.txt 2>&1
grep -E "^==|time_duration|issue_active|thread_inst|pipe_fp64|registers_per|inst_executed.sum" $D/ncu_cfg3_keys.txt
rm -f $D/ncu_cfg3/*.src.csv.gz
