#!/bin/bash
# round 2, GPU call 27: launch-bound variants of k_fbw_eval / k_fbw_book / k_t1prep / k_fastT*, A/B on Au20 and cfg5_60
set -u
D=gpurun_out/r2c27; mkdir -p $D
rm -f gpurun_out/ab_kernels.jsonl
for so in libecp_b200 libecp_b200_fbw6 libecp_b200_fbw10 libecp_b200_v1 libecp_b200_v2; do
  for wl in cfg3 cfg5_60; do
    echo "== $so $wl"
    LIBECP_B200_SO=$PWD/libecp_b200/lib/$so.so timeout 300 python tools/ab_kernels.py $wl 2>&1 | tail -1 | cut -c1-300
  done
done
cp gpurun_out/ab_kernels.jsonl $D/
