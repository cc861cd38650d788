"""One rank of a sharded configuration-5 run on a single GPU: family times of a serial pass, overlapped device time and
wall time per pass for rank r of W (libecp_b200_set_shard) - what a rank of a W-GPU run does, without the other ranks.
Usage (GPU box): python tools/shard_profile.py [W] [rank ...]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from libecp_b200 import capi, synth  # noqa: E402

KEYS = ["ms_tables", "ms_fastT", "ms_fallback", "ms_link", "ms_type1", "ms_chi", "ms_shift", "ms_device_total", "ms_build"]


def main():
    W = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    ranks = [int(x) for x in sys.argv[2:]] or [0]
    s = synth.cfg5(500)
    capi.lib().libecp_b200_set_host_threads(int(os.environ.get("HOST_THREADS", "4")))
    for r in ranks:
        with capi.Handle(s) as h:
            h.set_shard(r, W)
            for _ in range(3):
                h.integrals_device()
            res = {}
            for serial in (1, 0):
                h.set_serial_kernels(bool(serial))
                h.integrals_device()
                agg, wall, n = None, 0.0, 6
                for _ in range(n):
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    h.integrals_device()
                    wall += time.perf_counter() - t0
                    st = h.stats()
                    agg = st if agg is None else {k: agg[k] + v for k, v in st.items() if isinstance(v, (int, float))}
                res["serial" if serial else "overlap"] = dict({k: round(agg[k] / n, 3) for k in KEYS}, wall_ms=round(1e3 * wall / n, 3),
                                                              batches=agg["batches"] / n, triples=agg["executed_triples"] / n)
        print(json.dumps({"world": W, "rank": r, **res}), flush=True)


if __name__ == "__main__":
    main()
