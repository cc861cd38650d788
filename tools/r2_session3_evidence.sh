#!/bin/bash
# round 2, session 3, evidence call: smoke, GPU tier, bench (both arms), launch list, ncu --set full of the kernels of the session
set -u
D=gpurun_out/r2s3final; mkdir -p $D
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
( timeout 1500 python -m pytest tests -m gpu -q ) > $D/pytest_gpu.log 2>&1
tail -3 $D/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > $D/bench_cfg5_n1.json 2> $D/bench.err
python tools/bench_brief.py < $D/bench_cfg5_n1.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $D/bench_reference_arm.json 2>> $D/bench.err
cut -c1-300 $D/bench_reference_arm.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $D/launches_cfg5.csv \
   python bench.py --steps 1 --warmup 1 --no-cpu --no-secondary --no-parity > /dev/null 2>&1
python tools/launch_summary.py $D/launches_cfg5.csv > $D/launches_cfg5.summary.txt 2>&1; head -14 $D/launches_cfg5.summary.txt
gzip -f $D/launches_cfg5.csv
for spec in k_type1A:4 k_type1S:4 k_rows_flag:0 k_rows_pack_sparse:0; do
  k=${spec%%:*}; skip=${spec##*:}
  bash tools/ncu_capture.sh cfg3 $D/ncu_cfg3 "$skip" "$k" > /dev/null 2>&1
done
python tools/ncu_keys.py "$D/ncu_cfg3/*.raw.csv" > $D/ncu_cfg3_keys.txt 2>&1
rm -f $D/ncu_cfg3/*.src.csv.gz
LIBECP_B200_TRACE=1 timeout 900 ncu --set full --clock-control none -k "regex:k_type1A|k_type1S|k_type1L|k_t1prep" -s 22 -c 22 -f -o /tmp/cfg5_t1 \
   python bench.py --steps 1 --warmup 1 --no-cpu --no-secondary --no-parity > /dev/null 2> $D/cfg5_family_trace.log
ncu -i /tmp/cfg5_t1.ncu-rep --page raw --csv > $D/ncu_full_cfg5_type1_family.raw.csv 2>/dev/null
grep "batch:" $D/cfg5_family_trace.log | head -3
du -sh $D
