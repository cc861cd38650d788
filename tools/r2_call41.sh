#!/bin/bash
# round 2, GPU call 41: lane / chunk statistics of k_type1S (diagnostic build)
set -u
D=gpurun_out/r2c41; mkdir -p $D
LIBECP_B200_SO=$PWD/libecp_b200/lib/libecp_b200_t1stats.so timeout 300 python tools/t1_stats.py cfg3 cfg5_60 cfg5_200 > $D/t1_stats.out 2>&1
cat $D/t1_stats.out
