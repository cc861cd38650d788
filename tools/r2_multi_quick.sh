#!/bin/bash
# quick 2-rank bench (e2e result check over the ranks)
set -u
D=gpurun_out/r2multi; mkdir -p $D
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 \
    bench.py --gpus 2 --steps 2 --warmup 3 > $D/bench_n2_quick.json 2> $D/bench_n2_quick.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2multi/bench_n2_quick.json").read().strip().splitlines()[-1])
print("value ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "e2e parity", d["e2e"].get("parity"), "parity ok", d["parity"]["ok"])
PY
tail -2 $D/bench_n2_quick.err | cut -c1-200
