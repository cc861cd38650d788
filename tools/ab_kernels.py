"""A/B timing of kernel variants selected by environment variables (read at handle creation).

    python tools/ab_kernels.py [workload] [VAR=a,b,c ...]

For every combination of the listed values: per-kernel device times of a serial (single-stream) pass and the
overlapped device time, plus a checksum of the matrix so that variants can be compared for bit-identity.
Appends one JSON line per combination to gpurun_out/ab_kernels.jsonl.
"""
import itertools
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libecp_b200 import capi, synth  # noqa: E402

KEYS = ["ms_tables", "ms_fastT", "ms_fallback", "ms_link", "ms_type1", "ms_chi", "ms_shift", "ms_device_total", "ms_build"]


def workload(name):
    if name == "cfg3":
        return synth.cfg3(20)
    if name.startswith("au"):
        return synth.cfg3(int(name[2:]))
    if name.startswith("cfg5_"):
        return synth.cfg5(int(name.split("_")[1]))
    if name == "cfg4b":
        return synth.cfg4("b")
    raise SystemExit("unknown workload " + name)


def main():
    args = sys.argv[1:]
    wl = args[0] if args and "=" not in args[0] else "cfg3"
    sweeps = [a.split("=", 1) for a in args if "=" in a]
    names = [k for k, _ in sweeps]
    values = [v.split(",") for _, v in sweeps]
    s = workload(wl)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    out = open(os.path.join(ROOT, "gpurun_out", "ab_kernels.jsonl"), "a")
    import torch

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for combo in itertools.product(*values) if values else [()]:
        for k, v in zip(names, combo):
            if v == "-":
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        with capi.Handle(s) as h:
            rc, M = h.integrals_host()
            chk = float(np.abs(M).sum())
            for _ in range(2):
                h.integrals_device()
            res = {}
            for serial in (1, 0):
                h.set_serial_kernels(bool(serial))
                h.integrals_device()
                agg = None
                n = 5
                for _ in range(n):
                    flush.zero_()
                    torch.cuda.synchronize()
                    h.integrals_device()
                    st = h.stats()
                    agg = st if agg is None else {k: agg[k] + v for k, v in st.items()}
                res["serial" if serial else "overlap"] = {k: round(agg[k] / n, 4) for k in KEYS}
        line = {"workload": wl, "env": dict(zip(names, combo)), "rc": rc, "abs_sum": chk, **res}
        print(json.dumps(line), flush=True)
        out.write(json.dumps(line) + "\n")
        out.flush()


if __name__ == "__main__":
    main()
