#!/bin/bash
# round 2, GPU call 5: ncu --set full of k_link4 (d-d, Au20), k_fallbackG, k_fastT, k_fastT2, k_shiftJ, k_shiftI on Au20
set -u
D=gpurun_out/r2c5; mkdir -p $D
tools/ncu_capture.sh cfg3 $D 3 'k_link4' > /dev/null 2>&1
tools/ncu_capture.sh cfg3 $D 0 'k_fallbackG' 'k_fastT\(' 'k_fastT2' 'k_shiftJ' 'k_shiftI' 'k_chi' 'k_t1prep' > /dev/null 2>&1
for f in $D/*.raw.csv; do python tools/ncu_keys.py $f | head -32; done
for f in $D/*.src.txt; do echo "== $f"; head -30 $f | cut -c1-170; done
