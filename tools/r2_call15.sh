#!/bin/bash
# round 2, GPU call 15: full GPU tier on the committed tree (f1, f4, device unit tests, cheaper Bessel) + bench
set -u
D=gpurun_out/r2c15; mkdir -p $D
( timeout 1500 python -m pytest tests -m gpu -q ) > $D/pytest_gpu.log 2>&1
tail -15 $D/pytest_gpu.log
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu > $D/bench.json 2>> $D/bench.err
python tools/bench_brief.py < $D/bench.json
