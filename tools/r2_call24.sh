#!/bin/bash
# round 2, GPU call 24: e2e against the number of host row panels
set -u
D=gpurun_out/r2c24; mkdir -p $D
for P in 2 3 4 6; do
  LIBECP_B200_HOST_PANELS=$P timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu --no-secondary --no-parity > $D/bench_p$P.json 2>> $D/bench.err
  python -c "
import json; d=json.load(open('$D/bench_p$P.json')); e=d['e2e']; print('panels $P: e2e %.1f ms (init %.1f, integrate+d2h %.1f, free %.1f), value %.1f ms' % (e['ms_per_step'], e['ms_init'], e['ms_integrate_d2h'], e['ms_free'], d['ms_per_step']))"
done
