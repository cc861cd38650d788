#!/bin/bash
# round 2, GPU call 26: occupancy variants (separate builds) of the type-1 kernels and k_fbw_eval, A/B on Au20 and cfg5_60
set -u
D=gpurun_out/r2c26; mkdir -p $D
rm -f gpurun_out/ab_kernels.jsonl
for so in libecp_b200 libecp_b200_t1a libecp_b200_t1b libecp_b200_fbw8; do
  for wl in cfg3 cfg5_60; do
    echo "== $so $wl"
    LIBECP_B200_SO=$PWD/libecp_b200/lib/$so.so timeout 300 python tools/ab_kernels.py $wl 2>&1 | tail -1 | cut -c1-300
  done
done
for tb in 32 96 128; do
  echo "== t1block $tb"
  LIBECP_B200_T1BLOCK=$tb timeout 300 python tools/ab_kernels.py cfg3 2>&1 | tail -1 | cut -c1-300
done
cp gpurun_out/ab_kernels.jsonl $D/
