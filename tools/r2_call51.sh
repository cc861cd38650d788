#!/bin/bash
# round 2, last GPU call: HEAD sanity - smoke, GPU tier, default bench line
set -u
D=gpurun_out/r2c51; mkdir -p $D
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( timeout 900 python -m pytest tests -m gpu -q ) > $D/pytest_gpu.log 2>&1
tail -2 $D/pytest_gpu.log
timeout 600 python bench.py > $D/bench_default.json 2> $D/bench.err
python tools/bench_brief.py < $D/bench_default.json | head -2
