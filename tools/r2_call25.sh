#!/bin/bash
# round 2, GPU call 25: new parity tests (random systems, derivative triangle) + time of a derivative run
set -u
D=gpurun_out/r2c25; mkdir -p $D
( timeout 1500 python -m pytest tests -m gpu -q ) > $D/pytest_gpu.log 2>&1
tail -8 $D/pytest_gpu.log
python - <<'PY'
import time, numpy as np
from libecp_b200 import capi, synth
# first-derivative run on a 20-atom cluster with TZ(2) + ECP(4) (the reference's validity domain: lbs + n <= L - 1)
c = synth.au20_coords(20)
s = synth.assemble("au20_tz2", c, [synth.tz_basis(2)] * 20, [synth.ecp_set(4)] * 20)
for n in (0, 1):
    with capi.Handle(s, n=n) as h:
        h.callbacks(keep_blocks=False)
        t0 = time.perf_counter(); rc, recs = h.callbacks(keep_blocks=False); dt = time.perf_counter() - t0
        st = h.stats()
    print(f"Au20 TZ(2)+ECP(4) n={n}: rc {rc}, {len(recs)} callbacks, {st['executed_triples']} triples, wall {1e3*dt:.1f} ms (python callback replay included), device {st['ms_device_total']:.2f} ms")
PY
