#!/bin/bash
# round 2, GPU call 39: where inside the download set-up do the stalls of the streamed download sit?
set -u
D=gpurun_out/r2c39; mkdir -p $D
timeout 300 python tools/e2e_outliers.py 16 > $D/outliers_stream.out 2>&1
head -2 $D/outliers_stream.out
grep "sparse d2h" $D/outliers_stream.out | awk '{ if ($16+0 > 5.0) print }' | cut -c1-330 | head -20
