#!/bin/bash
# Evidence round for profiles/<round>/: GPU parity log, default bench line, ncu launch lists, ncu --set full captures.
# usage: tools/gpu_profile_round.sh <tag>       (outputs under gpurun_out/<tag>_*)
set -u
TAG=${1:-p}
O=gpurun_out/${TAG}
mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -q ) > $O/pytest_gpu.log 2>&1
tail -3 $O/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench_cfg5_n1.json 2> $O/bench_stderr.log
python tools/bench_brief.py < $O/bench_cfg5_n1.json
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_reference_arm.json 2>> $O/bench_stderr.log
python tools/bench_brief.py < $O/bench_reference_arm.json
# launch lists (serialised, cold cache: compare shares)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches_bench_default_cfg5.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-secondary > /dev/null 2>&1
python tools/launch_summary.py $O/launches_bench_default_cfg5.csv > $O/launches_bench_default_cfg5.summary.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_cfg3.csv \
    python bench.py --workload cfg3 --steps 1 --warmup 3 --no-cpu --no-secondary > /dev/null 2>&1
python tools/launch_summary.py $O/launches_cfg3.csv > $O/launches_cfg3.summary.txt
head -12 $O/launches_cfg3.summary.txt
# full captures on Au20 (one launch each; skip counts pick the warm 4th pass)
bash tools/ncu_capture.sh cfg3 $O/ncu 3 k_fallbackG k_fastT k_fastT2 k_chi k_shiftJ k_shiftI k_Ftab > /dev/null 2>&1
bash tools/ncu_capture.sh cfg3 $O/ncu 24 k_type1S > /dev/null 2>&1
bash tools/ncu_capture.sh cfg3 $O/ncu 58 k_link > /dev/null 2>&1
rm -f $O/ncu/*.src.csv.gz
ls $O/ncu | head -30
