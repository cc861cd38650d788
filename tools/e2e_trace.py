"""e2e (libECP_init + libecp_b200_integrals_host + libECP_free) on configuration 5 per number of host panels, with the
library's trace lines, and the fill of the result matrix at the granularity of the host add (runs of 32 doubles).
Usage (GPU box): python tools/e2e_trace.py [panels ...]"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from libecp_b200 import capi, synth  # noqa: E402


def main():
    panels = [int(x) for x in sys.argv[1:]] or [1, 2, 3]
    s = synth.cfg5(500)
    dim = int(s["dim"])
    host = torch.zeros((dim, dim), dtype=torch.float64).pin_memory().numpy()
    capi.lib().libecp_b200_set_host_threads(max(1, (os.cpu_count() or 2) - 1))

    def once():
        host[:] = 0.0
        t0 = time.perf_counter()
        with capi.Handle(s) as hh:
            t1 = time.perf_counter()
            capi.lib().libecp_b200_integrals_host(C.c_void_p(hh.h), dim, host.ctypes.data_as(capi._pd))
            t2 = time.perf_counter()
        t3 = time.perf_counter()
        return 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2)

    for P, mode in [(P, m) for m in (("sparse", "dense") if os.environ.get("E2E_DENSE") else ("sparse",)) for P in panels]:
        os.environ["LIBECP_B200_HOST_PANELS"] = str(P)
        os.environ["LIBECP_B200_D2H"] = mode
        os.environ.pop("LIBECP_B200_TRACE", None)
        once()
        once()
        ts = [once() for _ in range(3)]
        print(f"{mode} panels {P}: init/run/free ms", [tuple(round(x, 1) for x in t) for t in ts], flush=True)
        os.environ["LIBECP_B200_TRACE"] = "1"
        sys.stderr.write(f"---- trace, {mode}, panels {P}\n")
        sys.stderr.flush()
        once()
    os.environ.pop("LIBECP_B200_TRACE", None)
    # fill of the upper triangle at the granularity of the host add
    nz32 = tot32 = 0
    nz = 0
    for i in range(0, dim, 7):
        row = host[i, i:]
        m = len(row) // 32 * 32
        r = row[:m].reshape(-1, 32)
        nz32 += int((r != 0).any(axis=1).sum())
        tot32 += r.shape[0]
        nz += int((row != 0).sum())
    print(f"fill: runs of 32 with a non-zero {nz32 / tot32:.3f} (sampled rows), elements non-zero {nz / (tot32 * 32):.3f}")


if __name__ == "__main__":
    main()
