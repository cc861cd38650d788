#!/bin/bash
# round 2, GPU call 46: one rank of 8 on one GPU - row deal by shell (default) vs by (atom, l)
set -u
D=gpurun_out/r2c46; mkdir -p $D
timeout 400 python tools/shard_profile.py 8 0 1 2 3 4 5 6 7 > $D/shard8_default.jsonl 2>&1
LIBECP_B200_DEAL=atom timeout 400 python tools/shard_profile.py 8 0 1 2 3 4 5 6 7 > $D/shard8_atom.jsonl 2>&1
python - <<'PY'
import json
for f in ("default","atom"):
    rows=[json.loads(l) for l in open(f"gpurun_out/r2c46/shard8_{f}.jsonl") if l.startswith("{")]
    print(f, "device overlap", [r["overlap"]["ms_device_total"] for r in rows], "wall", [r["overlap"]["wall_ms"] for r in rows])
    print("   serial link", [r["serial"]["ms_link"] for r in rows], "type1", [r["serial"]["ms_type1"] for r in rows], "tables", [r["serial"]["ms_tables"] for r in rows])
    print("   triples", [int(r["overlap"]["triples"]) for r in rows])
PY
