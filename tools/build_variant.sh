#!/bin/bash
# build a second copy of the product library with other nvcc flags for A/B timing (select with LIBECP_B200_SO=<path>)
# usage: tools/build_variant.sh <name> <extra nvcc flags...>      -> libecp_b200/lib/libecp_b200_<name>.so
set -e
NAME=$1; shift
cd "$(dirname "$0")/.."
O=libecp_b200/lib/obj
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-fopenmp "$@" \
     -c libecp_b200/csrc/ecp_cuda.cu -o $O/ecp_cuda_$NAME.o
nvcc -shared -o libecp_b200/lib/libecp_b200_$NAME.so $O/tables.o $O/builder.o $O/api.o $O/loaders.o $O/ecp_cuda_$NAME.o -lm -lgomp
echo libecp_b200/lib/libecp_b200_$NAME.so
