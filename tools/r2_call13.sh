#!/bin/bash
# round 2, GPU call 13: f1 (derivatives) + f4 (ordering) + k_fastA/B parity, A/B of the fast path, ncu of the type-1 kernels
set -u
D=gpurun_out/r2c13; mkdir -p $D
( timeout 1500 python -m pytest tests -m gpu -q -x ) > $D/pytest_gpu.log 2>&1
tail -15 $D/pytest_gpu.log
rm -f gpurun_out/ab_kernels.jsonl
for wl in cfg3 cfg5_60; do
  timeout 300 python tools/ab_kernels.py $wl LIBECP_B200_FAST=old 2>&1 | tail -1 | cut -c1-330
  timeout 300 python tools/ab_kernels.py $wl LIBECP_B200_FAST=new LIBECP_B200_FASTUNROLL=1,2,4 2>&1 | tail -3 | cut -c1-330
done
cp gpurun_out/ab_kernels.jsonl $D/
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu > $D/bench.json 2>> $D/bench.err
python tools/bench_brief.py < $D/bench.json
bash tools/ncu_capture.sh cfg3 $D/ncu_cfg3 4 'k_type1S' 'k_type1L' 'k_t1prep' 'k_fastA' 'k_fastB' > /dev/null 2>&1
python tools/ncu_keys.py "$D/ncu_cfg3/*.raw.csv" > $D/ncu_cfg3_keys.txt 2>&1
grep -E "^==|time_duration|issue_active|thread_inst|pipe_fp64|registers_per" $D/ncu_cfg3_keys.txt
rm -f $D/ncu_cfg3/*.src.csv.gz
