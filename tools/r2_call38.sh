#!/bin/bash
# round 2, GPU call 38: catch the slow e2e runs
set -u
D=gpurun_out/r2c38; mkdir -p $D
timeout 300 python tools/e2e_outliers.py 30 > $D/outliers_stream.out 2>&1
head -3 $D/outliers_stream.out; grep -c "slow run" $D/outliers_stream.out
grep -A40 "slow run" $D/outliers_stream.out | grep -v "alloc+H2D" | head -60
LIBECP_B200_STREAM_D2H=0 timeout 300 python tools/e2e_outliers.py 30 > $D/outliers_nostream.out 2>&1
head -3 $D/outliers_nostream.out
