#!/bin/bash
# round 2, GPU call 31: e2e timeline per number of host panels + matrix fill
set -u
D=gpurun_out/r2c31; mkdir -p $D
timeout 600 python tools/e2e_trace.py 1 2 3 4 > $D/e2e_trace.out 2> $D/e2e_trace.err
cat $D/e2e_trace.out
grep -v "batch:" $D/e2e_trace.err | tail -60
