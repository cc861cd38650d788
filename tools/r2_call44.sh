#!/bin/bash
# round 2, GPU call 44: type-1 kernels one by one (launch list), k_type1A at 3 resident blocks, ncu --set full of k_type1A
set -u
D=gpurun_out/r2c44; mkdir -p $D
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_type1|k_t1prep" --csv --log-file $D/launches_t1_cfg5.csv \
   python bench.py --steps 1 --warmup 1 --no-cpu --no-secondary --no-parity > /dev/null 2>&1
python tools/launch_summary.py $D/launches_t1_cfg5.csv | head -40 | tee $D/launches_t1_cfg5.summary.txt
gzip -f $D/launches_t1_cfg5.csv
rm -f gpurun_out/ab_kernels.jsonl
for wl in cfg3 cfg5_60; do
  timeout 300 python tools/ab_kernels.py $wl LIBECP_B200_X=minb2 2>&1 | tail -1 | cut -c1-330
  LIBECP_B200_SO=$PWD/libecp_b200/lib/libecp_b200_t1a3.so timeout 300 python tools/ab_kernels.py $wl LIBECP_B200_X=minb3 2>&1 | tail -1 | cut -c1-330
done
cp gpurun_out/ab_kernels.jsonl $D/ab_t1a_minb.jsonl
bash tools/ncu_capture.sh cfg3 $D/ncu 12 "k_type1A" > /dev/null 2>&1
python tools/ncu_keys.py $D/ncu/k_type1A.raw.csv 2>/dev/null | head -30
head -30 $D/ncu/k_type1A.src.txt
