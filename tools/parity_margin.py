"""Parity margin on Au20 (cfg3): max over the matrix of |x-ref| / (1e-12 + 1e-10 |ref|) against the golden fixture of the
compiled reference (1.0 = at tolerance).  GPU only.  Usage: python tools/parity_margin.py [ENV=VALUE ...]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
for a in sys.argv[1:]:
    if "=" in a:
        k, v = a.split("=", 1)
        os.environ[k] = v
from libecp_b200 import capi, synth  # noqa: E402

z = np.load(os.path.join(ROOT, "tests", "golden", "cfg3_matrix.npz"))
dim = int(z["dim"])
ref = np.zeros((dim, dim))
ref[np.triu_indices(dim)] = z["triu"]
with capi.Handle(synth.cfg3(20)) as h:
    rc, M = h.integrals_host()
    st = h.stats()
d = np.abs(M - ref)
ratio = d / (1e-12 + 1e-10 * np.abs(ref))
i = np.unravel_index(np.argmax(ratio), ratio.shape)
print(f"parity_margin cfg3 {sys.argv[1:]}: rc={rc} max|d|={d.max():.3e} max ratio={ratio.max():.4f} at {i} ref={ref[i]:.6e} "
      f"elements>0.1: {(ratio > 0.1).sum()} fallback_items={st['fallback_items']} fast_failed={st['fast_failed']} "
      f"t1_fallback_pairs={st['type1_fallback_pairs']}")
