#!/bin/bash
# round 2, GPU call 47: host add with prefetched destinations - tests, add time (no streaming: whole matrix in one download), e2e
set -u
D=gpurun_out/r2c47; mkdir -p $D
( timeout 600 python -m pytest tests -m gpu -q -x -k "sparse_download or streamed or row_panels or getintegrals" ) > $D/pytest.log 2>&1
tail -3 $D/pytest.log
LIBECP_B200_STREAM_D2H=0 timeout 300 python tools/e2e_trace.py 1 2>&1 | grep "panels 1:\|sparse d2h" | cut -c1-330
timeout 300 python tools/e2e_outliers.py 20 2>&1 | head -2
