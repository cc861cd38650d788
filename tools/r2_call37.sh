#!/bin/bash
# round 2, GPU call 37: why is e2e slow inside bench.py (178 ms) and fast in tools/e2e_trace.py (91 ms)? traces
set -u
D=gpurun_out/r2c37; mkdir -p $D
LIBECP_B200_TRACE=1 timeout 500 python bench.py --steps 3 --warmup 3 --no-cpu --no-secondary --no-parity > $D/bench.json 2> $D/bench_trace.err
python tools/bench_brief.py < $D/bench.json | head -1
grep -n "integrals_host\|sparse d2h" $D/bench_trace.err | tail -30
grep -c "batch:" $D/bench_trace.err
tail -60 $D/bench_trace.err | grep "batch:" | tail -24
