#!/bin/bash
# round 2, GPU call 49: compute-sanitizer on the final tree (new kernels: k_type1A, survivor mode of k_type1S, k_rows_flag / k_rows_pack_sparse)
set -u
D=gpurun_out/r2c49; mkdir -p $D
timeout 400 compute-sanitizer --tool memcheck python tools/gpu_diag.py au4 cfg4b > $D/sanitizer_memcheck_au4_cfg4b.log 2>&1
grep -n "ERROR SUMMARY\|== au4\|== cfg4b\|max" $D/sanitizer_memcheck_au4_cfg4b.log | head -8
timeout 400 compute-sanitizer --tool racecheck python tools/gpu_diag.py au2 > $D/sanitizer_racecheck_au2.log 2>&1
grep -n "RACECHECK SUMMARY\|== au2\|hazard" $D/sanitizer_racecheck_au2.log | head -8
