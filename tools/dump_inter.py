"""sha256 of an intermediate array of a run (libecp_b200_debug_fetch: F, T, gamma, chi, Q) - compare builds of the library
(LIBECP_B200_SO=...) for bit-identity.  Usage: python tools/dump_inter.py <what> <n> [workload ...]"""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libecp_b200 import capi, synth  # noqa: E402

what, n = sys.argv[1], int(sys.argv[2])
for wl in sys.argv[3:] or ["cfg3_4"]:
    if wl.startswith("cfg3_"):
        s = synth.cfg3(int(wl.split("_")[1]))
    elif wl.startswith("cfg5_"):
        s = synth.cfg5(int(wl.split("_")[1]))
    elif wl == "cfg4a":
        s = synth.cfg4("a")
    else:
        s = synth.cfg4("b")
    with capi.Handle(s) as h:
        h.integrals_device()
        a = h.debug_fetch(what, n)
    print(wl, what, hashlib.sha256(a.tobytes()).hexdigest()[:16], float(abs(a).sum()))
