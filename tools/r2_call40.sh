#!/bin/bash
# round 2, GPU call 40: streamed download with handle-owned, parked download buffers - outliers gone? + GPU tier + bench
set -u
D=gpurun_out/r2c40; mkdir -p $D
timeout 300 python tools/e2e_outliers.py 30 > $D/outliers_stream.out 2>&1
head -2 $D/outliers_stream.out
LIBECP_B200_STREAM_D2H=0 timeout 300 python tools/e2e_outliers.py 12 > $D/outliers_nostream.out 2>&1
head -2 $D/outliers_nostream.out
( timeout 900 python -m pytest tests -m gpu -q -x ) > $D/pytest_gpu.log 2>&1
tail -3 $D/pytest_gpu.log
timeout 500 python bench.py --steps 5 --warmup 3 --no-cpu > $D/bench.json 2>> $D/bench.err
python tools/bench_brief.py < $D/bench.json | head -2
