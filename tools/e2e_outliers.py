"""Repeat the e2e call (libECP_init + libecp_b200_integrals_host + libECP_free) on configuration 5 with the library's
trace on and a long-lived second handle alive (as in bench.py); prints every run's time and, for runs slower than
1.15x the median, their trace lines.  Usage (GPU box): python tools/e2e_outliers.py [runs]"""
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
    runs = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    s = synth.cfg5(500)
    dim = int(s["dim"])
    host = torch.zeros((dim, dim), dtype=torch.float64).pin_memory().numpy()
    capi.lib().libecp_b200_set_host_threads(max(1, (os.cpu_count() or 2) - 1))
    keep = capi.Handle(s)
    for _ in range(3):
        keep.integrals_device()
    os.environ["LIBECP_B200_TRACE"] = "1"
    r, w = os.pipe()
    saved = os.dup(2)
    times, traces = [], []
    for k in range(runs + 2):
        host[:] = 0.0
        torch.cuda.synchronize()
        os.dup2(w, 2)
        t0 = time.perf_counter()
        with capi.Handle(s) as hh:
            t1 = time.perf_counter()
            capi.lib().libecp_b200_integrals_host(C.c_void_p(hh.h), dim, host.ctypes.data_as(capi._pd))
            t2 = time.perf_counter()
        t3 = time.perf_counter()
        os.dup2(saved, 2)
        os.set_blocking(r, False)
        buf = b""
        try:
            while True:
                chunk = os.read(r, 1 << 16)
                if not chunk:
                    break
                buf += chunk
        except BlockingIOError:
            pass
        if k >= 2:
            times.append((1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2)))
            traces.append(buf.decode(errors="replace"))
    tot = np.array([sum(t) for t in times])
    med = float(np.median(tot))
    print("runs", runs, "median", round(med, 1), "mean", round(float(tot.mean()), 1), "max", round(float(tot.max()), 1))
    print("all:", [round(float(x), 1) for x in tot])
    for k, t in enumerate(times):
        if sum(t) > 1.15 * med:
            print(f"---- slow run {k}: init/run/free {tuple(round(x, 1) for x in t)}")
            print(traces[k])
    keep.close()


if __name__ == "__main__":
    main()
