#!/bin/bash
# round 2, GPU call 34: streamed download - test, e2e with / without, per batch size
set -u
D=gpurun_out/r2c34; mkdir -p $D
( timeout 600 python -m pytest tests -m gpu -q -x -k "streamed or sparse_download or row_panels or getintegrals or config5_full" ) > $D/pytest.log 2>&1
tail -8 $D/pytest.log
for bt in - 2000000 1500000; do
  for st in 1 0; do
    echo "== BATCH_TRIPLES=$bt STREAM=$st"
    ( [ $bt != - ] && export LIBECP_B200_BATCH_TRIPLES=$bt; export LIBECP_B200_STREAM_D2H=$st; timeout 300 python tools/e2e_trace.py 1 2>&1 | grep -v "batch:\|first batch\|^fill\|dense" | tail -6 )
  done
done > $D/e2e_stream.out 2>&1
cat $D/e2e_stream.out
