#!/bin/bash
# round 2, GPU call 23: full GPU tier (spherical output, initial batch estimate), smoke(), bench line with cfg5 traffic
set -u
D=gpurun_out/r2c23; mkdir -p $D
( timeout 1500 python -m pytest tests -m gpu -q ) > $D/pytest_gpu.log 2>&1
tail -5 $D/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu > $D/bench.json 2>> $D/bench.err
python tools/bench_brief.py < $D/bench.json | head -3
python -c "
import json; d=json.load(open('$D/bench.json')); print('traffic', d['roofline']['traffic'], d['gpu_launches'])"
