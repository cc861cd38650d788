#!/bin/bash
# round 2, GPU call 12: baseline of the committed tree - GPU tier, bench, compute-sanitizer (memcheck + racecheck on au4 /
# cfg4b), launch list and ncu --set full of the main kernels on Au20 and on cfg5
set -u
D=gpurun_out/r2c12; mkdir -p $D
( timeout 1200 python -m pytest tests -m gpu -q ) > $D/pytest_gpu.log 2>&1
tail -5 $D/pytest_gpu.log
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu > $D/bench.json 2>> $D/bench.err
python tools/bench_brief.py < $D/bench.json
# sanitizers (small shapes: the tools slow kernels down 10-100x)
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/gpu_diag.py au4 cfg4b > $D/sanitizer_memcheck.log 2>&1
echo "memcheck rc $?"; grep -E "ERROR SUMMARY|matrix:" $D/sanitizer_memcheck.log | tail -4
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/gpu_diag.py au2 > $D/sanitizer_racecheck.log 2>&1
echo "racecheck rc $?"; grep -E "RACECHECK SUMMARY|matrix:" $D/sanitizer_racecheck.log | tail -4
# launch list (serialised, cold) of one cfg5 step
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $D/launches_cfg5.csv \
   python bench.py --steps 1 --warmup 1 --no-cpu --no-secondary > /dev/null 2>&1
python tools/launch_summary.py $D/launches_cfg5.csv > $D/launches_cfg5.summary.txt 2>&1; head -30 $D/launches_cfg5.summary.txt
gzip -f $D/launches_cfg5.csv
bash tools/ncu_capture.sh cfg3 $D/ncu_cfg3 0 'k_type1S<4' 'k_type1S<6' 'k_fastT$' 'k_fastT2' 'k_chi' 'k_shift2' 'k_link4' 'k_fbw_eval' 'k_Ftab2' 'k_type1L<4' > /dev/null 2>&1
bash tools/ncu_capture.sh cfg5 $D/ncu_cfg5 0 'k_type1S<2' 'k_type1S<3' 'k_fastT$' 'k_chi' > /dev/null 2>&1
python tools/ncu_keys.py "$D/ncu_cfg3/*.raw.csv" > $D/ncu_cfg3_keys.txt 2>&1
python tools/ncu_keys.py "$D/ncu_cfg5/*.raw.csv" > $D/ncu_cfg5_keys.txt 2>&1
rm -f $D/ncu_cfg3/*.src.csv.gz $D/ncu_cfg5/*.src.csv.gz
du -sh $D
