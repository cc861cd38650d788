#!/bin/bash
# round 2, GPU call 52: k_type1A with 56 pairs per block and three resident blocks
set -u
D=gpurun_out/r2c52; mkdir -p $D
rm -f gpurun_out/ab_kernels.jsonl
for wl in cfg3 cfg5_120; do
  timeout 200 python tools/ab_kernels.py $wl LIBECP_B200_X=pb8 2>&1 | tail -1 | cut -c1-300
  LIBECP_B200_SO=$PWD/libecp_b200/lib/libecp_b200_t1apb7.so timeout 200 python tools/ab_kernels.py $wl LIBECP_B200_X=pb7 2>&1 | tail -1 | cut -c1-300
done
cp gpurun_out/ab_kernels.jsonl $D/ab_t1a_pb7.jsonl
