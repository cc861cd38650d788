#!/bin/bash
# round 2, GPU call 19: device enumeration with staged slots (parity, bench, per-batch trace)
set -u
D=gpurun_out/r2c19; mkdir -p $D
( timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "enumeration or getintegrals or config5_full or structure" ) > $D/pytest_gpu.log 2>&1
tail -4 $D/pytest_gpu.log
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu > $D/bench.json 2>> $D/bench.err
python tools/bench_brief.py < $D/bench.json
LIBECP_B200_TRACE=1 timeout 300 python bench.py --steps 1 --warmup 2 --no-cpu --no-secondary > /dev/null 2> $D/trace.log
grep "batch:" $D/trace.log | tail -9 | cut -c1-200
