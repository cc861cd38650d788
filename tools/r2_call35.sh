#!/bin/bash
# round 2, GPU call 35: streamed download with smaller batches at the end of the pass
set -u
D=gpurun_out/r2c35; mkdir -p $D
( timeout 600 python -m pytest tests -m gpu -q -x -k "streamed or sparse_download" ) > $D/pytest.log 2>&1
tail -4 $D/pytest.log
for ev in 0 1; do
  echo "== STREAM_EVEN=$ev"
  ( [ $ev = 1 ] && export LIBECP_B200_STREAM_EVEN=1; timeout 300 python tools/e2e_trace.py 1 2>&1 | grep -v "batch:\|first batch\|^fill\|dense" )
done > $D/e2e_stream.out 2>&1
cat $D/e2e_stream.out
