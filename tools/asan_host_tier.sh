#!/bin/bash
# AddressSanitizer + UBSan build of the host C layer (tables.c, builder.c, api.c, loaders.c; the CUDA object is linked
# unchanged) and a run of the CPU test tier against it.  Log: profiles/r2/asan_host_tier.log
set -eu
cd "$(dirname "$0")/.."
OBJ=/tmp/asan_obj; mkdir -p $OBJ /tmp/asan_lib
for f in tables builder api loaders; do
  gcc -O1 -g -fPIC -fno-omit-frame-pointer -fsanitize=address,undefined -ffp-contract=off -std=gnu11 -fopenmp -Wall -Wno-comment \
      -c libecp_b200/csrc/$f.c -o $OBJ/$f.o
done
nvcc -shared -o /tmp/asan_lib/libecp_b200.so $OBJ/tables.o $OBJ/builder.o $OBJ/api.o $OBJ/loaders.o libecp_b200/lib/obj/ecp_cuda.o \
     -lm -lgomp -Xlinker -lasan -Xlinker -lubsan
export LIBECP_B200_SO=/tmp/asan_lib/libecp_b200.so
export LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)"
export ASAN_OPTIONS=detect_leaks=0:abort_on_error=0:halt_on_error=1
export UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1
python -m pytest tests/test_host.py tests/test_dist_cpu.py -q -m "not gpu" -p no:cacheprovider 2>&1 | tail -15
