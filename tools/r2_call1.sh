#!/bin/bash
# round 2, GPU call 1: full GPU tier on HEAD (incl. the full / sharded config-5 digest tests), the three never-run
# experimental variants, and a short bench line with the new parity key
set -u
mkdir -p gpurun_out/r2c1
( timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2c1/pytest_gpu.log 2>&1
tail -15 gpurun_out/r2c1/pytest_gpu.log
( LIBECP_B200_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "link_smem or fused_shift or ftab_compact" ) > gpurun_out/r2c1/experimental.log 2>&1
tail -12 gpurun_out/r2c1/experimental.log
timeout 200 python tools/ab_kernels.py cfg3 LIBECP_B200_LINK=-,smem LIBECP_B200_SHIFT=-,fused LIBECP_B200_FTAB=-,compact 2>&1 | tail -8 | cut -c1-600
timeout 200 python tools/ab_kernels.py cfg5_60 LIBECP_B200_LINK=-,smem LIBECP_B200_SHIFT=-,fused LIBECP_B200_FTAB=-,compact 2>&1 | tail -8 | cut -c1-600
cp gpurun_out/ab_kernels.jsonl gpurun_out/r2c1/ 2>/dev/null
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2c1/bench.json 2> gpurun_out/r2c1/bench.err
python tools/bench_brief.py < gpurun_out/r2c1/bench.json 2>/dev/null || head -c 1500 gpurun_out/r2c1/bench.json
tail -3 gpurun_out/r2c1/bench.err
