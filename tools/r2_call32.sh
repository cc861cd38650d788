#!/bin/bash
# round 2, GPU call 32: sparse download of the result matrix - bit-identity test, e2e per panel count, sparse vs dense
set -u
D=gpurun_out/r2c32; mkdir -p $D
( timeout 600 python -m pytest tests -m gpu -q -x -k "sparse_download or row_panels or getintegrals" ) > $D/pytest.log 2>&1
tail -15 $D/pytest.log
timeout 600 python tools/e2e_trace.py 1 2 3 > $D/e2e_trace.out 2> $D/e2e_trace.err
cat $D/e2e_trace.out
grep "d2h+add\|whole call\|----" $D/e2e_trace.err | tail -40
