#!/bin/bash
# round 2, GPU call 2: k_link3 + FMA contraction + fastT2 unroll: parity, A/B timings; diagnosis of the config-5 sample violations
set -u
D=gpurun_out/r2c2; mkdir -p $D
( timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_parity.py::test_config5_full_digest --deselect tests/test_gpu_parity.py::test_config5_sharded_gather_digest ) > $D/pytest_gpu.log 2>&1
tail -8 $D/pytest_gpu.log
rm -f gpurun_out/ab_kernels.jsonl
for wl in cfg3 cfg5_60; do
  timeout 200 python tools/ab_kernels.py $wl LIBECP_B200_LINK=global,3 2>&1 | tail -2 | cut -c1-420
  timeout 200 python tools/ab_kernels.py $wl LIBECP_B200_LINKTPB=4,8,32,64 2>&1 | tail -4 | cut -c1-420
  timeout 200 python tools/ab_kernels.py $wl LIBECP_B200_FASTUNROLL=2,4 2>&1 | tail -2 | cut -c1-420
  LIBECP_B200_SO=$PWD/libecp_b200/lib/libecp_b200_nofmad.so timeout 200 python tools/ab_kernels.py $wl LIBECP_B200_SOTAG=nofmad 2>&1 | tail -1 | cut -c1-420
done
cp gpurun_out/ab_kernels.jsonl $D/
timeout 600 python tools/parity_diag.py > $D/parity_diag.log 2>&1
tail -14 $D/parity_diag.log
cp gpurun_out/parity_diag.json $D/ 2>/dev/null
