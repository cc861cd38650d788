#!/bin/bash
# e2e trace of a 2-rank run (library trace lines of both ranks)
set -u
D=gpurun_out/r2multi; mkdir -p $D
LIBECP_B200_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NR:-2} --master-addr 127.0.0.1 --master-port 29519 \
    bench.py --gpus ${NR:-2} --steps 2 --warmup 3 --no-parity > $D/bench_n${NR:-2}_trace.json 2> $D/bench_n${NR:-2}_trace.err
python tools/bench_brief.py < $D/bench_n${NR:-2}_trace.json | head -3
grep -n "integrals_host\|sparse d2h\|first batch" $D/bench_n${NR:-2}_trace.err | tail -24 | cut -c1-330
