#!/bin/bash
# round 2, GPU call 8: launch list (per-launch durations) of one Au20 pass and ncu --set full of the wave kernels
set -u
D=gpurun_out/r2c8; mkdir -p $D
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 160 --csv --log-file $D/launches_cfg3.csv python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu --no-secondary --no-parity > /dev/null 2>&1
python tools/launch_summary.py $D/launches_cfg3.csv | head -50
tools/ncu_capture.sh cfg3 $D 0 'k_fbw_eval' 'k_fbw_book' > /dev/null 2>&1
tools/ncu_capture.sh cfg3 $D 3 'k_link4' > /dev/null 2>&1
for f in $D/*.raw.csv; do python tools/ncu_keys.py $f | head -26; done
for f in $D/*.src.txt; do echo "== $f"; head -16 $f | cut -c1-170; done
