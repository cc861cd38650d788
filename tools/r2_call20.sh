#!/bin/bash
# round 2, GPU call 20: type-1 chain started before the tables, two-batch passes: full GPU tier + bench
set -u
D=gpurun_out/r2c20; mkdir -p $D
( timeout 1500 python -m pytest tests -m gpu -q ) > $D/pytest_gpu.log 2>&1
tail -4 $D/pytest_gpu.log
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu > $D/bench.json 2>> $D/bench.err
python tools/bench_brief.py < $D/bench.json
