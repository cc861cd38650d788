#!/bin/bash
# round 2, multi-GPU call: bench.py under torchrun on N GPUs of one box (value, gathered step, C-ABI NCCL all-gather, parity)
# usage: tools/r2_multi.sh N [steps]
set -u
N=$1; STEPS=${2:-5}
D=gpurun_out/r2multi; mkdir -p $D
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps $STEPS --warmup 3 > $D/bench_n$N.json 2> $D/bench_n$N.err
python tools/bench_brief.py < $D/bench_n$N.json
tail -5 $D/bench_n$N.err | cut -c1-300
