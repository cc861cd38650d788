"""Per-block residency of the persistent kernels (LIBECP_B200_TAILS probe): python tools/tails.py [workload]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["LIBECP_B200_TAILS"] = "1"
from libecp_b200 import capi  # noqa: E402
from tools.ab_kernels import workload  # noqa: E402

s = workload(sys.argv[1] if len(sys.argv) > 1 else "cfg3")
with capi.Handle(s) as h:
    h.integrals_device()
    sys.stderr.write("==== second pass\n")
    h.integrals_device()
    print(h.stats())
