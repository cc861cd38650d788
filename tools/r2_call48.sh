#!/bin/bash
# round 2, GPU call 48: env-only knobs with the final kernels: type-1 block size (survivor kernel), triples per batch
set -u
D=gpurun_out/r2c48; mkdir -p $D
rm -f gpurun_out/ab_kernels.jsonl
timeout 200 python tools/ab_kernels.py cfg5_120 LIBECP_B200_T1BLOCK=32,64,128 2>&1 | cut -c1-260 | tail -3
cp gpurun_out/ab_kernels.jsonl $D/ab_t1block.jsonl
for bt in - 9000000 12000000; do
  ( [ $bt != - ] && export LIBECP_B200_BATCH_TRIPLES=$bt; timeout 200 python bench.py --steps 4 --warmup 3 --no-cpu --no-secondary --no-parity 2>/dev/null | python tools/bench_brief.py | head -1 | sed "s/^/batch $bt: /" )
done | tee $D/batch_size.out
