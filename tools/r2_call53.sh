#!/bin/bash
# round 2, GPU call 53: result check of the e2e path (host consumer on the full configuration 5) - test + bench line
set -u
D=gpurun_out/r2c53; mkdir -p $D
( timeout 300 python -m pytest tests -m gpu -q -x -k "host_consumer_digest" ) > $D/pytest.log 2>&1
tail -3 $D/pytest.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-secondary > $D/bench.json 2> $D/bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2c53/bench.json").read().strip().splitlines()[-1])
print("value ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "e2e parity", d["e2e"].get("parity"), "parity ok", d["parity"]["ok"])
PY
tail -3 $D/bench.err
