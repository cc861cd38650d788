#!/bin/bash
# round 2, GPU call 3: ncu --set full of the d-d link launch on Au20: k_link (global), k_link3 tpb 4 and 16
set -u
D=gpurun_out/r2c3; mkdir -p $D
LIBECP_B200_LINK=global tools/ncu_capture.sh cfg3 $D/global 0 'k_link<3, 3>' > /dev/null 2>&1
LIBECP_B200_LINKTPB=4 tools/ncu_capture.sh cfg3 $D/tpb4 3 'k_link3' > /dev/null 2>&1
LIBECP_B200_LINKTPB=16 tools/ncu_capture.sh cfg3 $D/tpb16 3 'k_link3' > /dev/null 2>&1
for f in $D/*/*.raw.csv; do python tools/ncu_keys.py $f | head -40; done
for f in $D/*/*.src.txt; do echo "== $f"; head -45 $f; done
