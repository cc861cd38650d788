#!/bin/bash
# multi-GPU trace: bench.py on N GPUs with the per-batch trace of rank 0 (stderr); extra environment in $2
set -u
N=$1; EXTRA=${2:-}
D=gpurun_out/r2multi; mkdir -p $D
env $EXTRA LIBECP_B200_TRACE=1 LIBECP_B200_BUILD_PROFILE=1 NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 2 --warmup 3 > $D/trace_n$N.json 2> $D/trace_n$N.err
python tools/bench_brief.py < $D/trace_n$N.json | head -4
grep -E "rank 0 batch|builder\] rank 0" $D/trace_n$N.err | tail -36 | cut -c1-200
nproc; cat /proc/cpuinfo | grep -c processor
