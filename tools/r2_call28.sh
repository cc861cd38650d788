#!/bin/bash
# round 2, GPU call 28: levels of the first fast-path launch, batch size (environment knobs)
set -u
D=gpurun_out/r2c28; mkdir -p $D
rm -f gpurun_out/ab_kernels.jsonl
for wl in cfg3 cfg5_60; do
  timeout 400 python tools/ab_kernels.py $wl LIBECP_B200_FASTLIM=3,4,5,6 2>&1 | tail -4 | sed 's/.*"env": //' | cut -c1-160
done
cp gpurun_out/ab_kernels.jsonl $D/
for bt in 2000000 3000000 4500000 6000000 9000000; do
  LIBECP_B200_BATCH_TRIPLES=$bt timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-secondary --no-parity > $D/bench_bt$bt.json 2>> $D/bench.err
  python -c "
import json; d=json.load(open('$D/bench_bt$bt.json')); print('batch $bt: %.2f ms/step device %.2f e2e %.1f' % (d['ms_per_step'], d['device_ms_per_step'], d['e2e']['ms_per_step']))"
done
