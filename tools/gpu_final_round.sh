#!/bin/bash
# Short evidence round (fits ~3.5 GPU-minutes): GPU parity log, default bench line, launch list of one cfg5 step,
# ncu --set full of the two heaviest kernels.  usage: tools/gpu_final_round.sh <tag>   (outputs under gpurun_out/<tag>/)
set -u
TAG=${1:-f}
O=gpurun_out/${TAG}
mkdir -p $O
( time timeout 200 python -m pytest tests -m gpu -q ) > $O/pytest_gpu.log 2>&1
tail -3 $O/pytest_gpu.log
timeout 300 python bench.py --steps 5 --warmup 3 > $O/bench_cfg5_n1.json 2> $O/bench_stderr.log
python tools/bench_brief.py < $O/bench_cfg5_n1.json
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 1700 --csv --log-file $O/launches_bench_cfg5.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-secondary > /dev/null 2>&1
python tools/launch_summary.py $O/launches_bench_cfg5.csv > $O/launches_bench_cfg5.summary.txt
head -8 $O/launches_bench_cfg5.summary.txt
bash tools/ncu_capture.sh cfg3 $O/ncu 24 k_type1S > /dev/null 2>&1
bash tools/ncu_capture.sh cfg3 $O/ncu 58 k_link > /dev/null 2>&1
rm -f $O/ncu/*.src.csv.gz
python tools/ncu_keys.py "$O/ncu/*.raw.csv" | grep -E "^==|time_duration|pipe_fp64|thread_inst_executed_per|issue_active|warps_active|dram__bytes"
