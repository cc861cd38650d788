#!/bin/bash
# round 2, GPU call 45: k_type1A pairs per block (8/8 = 64, 6/8, 5/8, 4/8 with 3 resident blocks)
set -u
D=gpurun_out/r2c45; mkdir -p $D
rm -f gpurun_out/ab_kernels.jsonl
for wl in cfg3 cfg5_120; do
  timeout 300 python tools/ab_kernels.py $wl LIBECP_B200_X=pb8 2>&1 | tail -1 | cut -c1-300
  for n in 6 5 4; do
    LIBECP_B200_SO=$PWD/libecp_b200/lib/libecp_b200_t1apb$n.so timeout 300 python tools/ab_kernels.py $wl LIBECP_B200_X=pb$n 2>&1 | tail -1 | cut -c1-300
  done
done
cp gpurun_out/ab_kernels.jsonl $D/ab_t1a_pb.jsonl
( timeout 300 python -m pytest tests -m gpu -q -x -k "type1_wave" ) 2>&1 | tail -2
