#!/bin/bash
# round 2, GPU call 18: device-side triple enumeration (parity, bench)
set -u
D=gpurun_out/r2c18; mkdir -p $D
( timeout 1500 python -m pytest tests -m gpu -q -x ) > $D/pytest_gpu.log 2>&1
tail -25 $D/pytest_gpu.log
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu > $D/bench.json 2>> $D/bench.err
python tools/bench_brief.py < $D/bench.json
tail -5 $D/bench.err
