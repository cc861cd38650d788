#!/bin/bash
# round 2, GPU call 33: full GPU tier + bench after the sparse download and n = 2
set -u
D=gpurun_out/r2c33; mkdir -p $D
( timeout 900 python -m pytest tests -m gpu -q -x ) > $D/pytest_gpu.log 2>&1
tail -5 $D/pytest_gpu.log
timeout 500 python bench.py --steps 5 --warmup 3 --no-cpu > $D/bench.json 2>> $D/bench.err
python tools/bench_brief.py < $D/bench.json | head -6
