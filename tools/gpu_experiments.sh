#!/bin/bash
# First GPU call of the next round (one B200, ~2 minutes): validate and time the variants that were written after the
# round-1 GPU budget was spent and are OFF by default:
#   LIBECP_B200_LINK=smem    k_link2, Omega slices + T staged in shared memory
#   LIBECP_B200_SHIFT=fused  one shift of 4 pi chi + 16 pi^2 gamma in matrix-only runs
#   LIBECP_B200_FTAB=compact k_Ftab2, only the window of every shell slot is tabulated
# usage: tools/gpu_experiments.sh [tag]      (outputs under gpurun_out/<tag>_*)
set -u
TAG=${1:-x}
mkdir -p gpurun_out
( LIBECP_B200_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "link_smem or fused_shift or ftab_compact" ) \
    > gpurun_out/${TAG}_experimental_pytest.log 2>&1
tail -5 gpurun_out/${TAG}_experimental_pytest.log
# per-kernel serial / overlapped times with a matrix checksum per combination (appends to gpurun_out/ab_kernels.jsonl)
timeout 200 python tools/ab_kernels.py cfg3 LIBECP_B200_LINK=-,smem LIBECP_B200_SHIFT=-,fused LIBECP_B200_FTAB=-,compact 2>&1 | tail -8 | cut -c1-400
timeout 200 python tools/ab_kernels.py cfg5_60 LIBECP_B200_LINK=-,smem LIBECP_B200_SHIFT=-,fused LIBECP_B200_FTAB=-,compact 2>&1 | tail -8 | cut -c1-400
