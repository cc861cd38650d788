"""GPU-side diagnostic: block-by-block comparison of the CUDA path with the oracle on small configs.
Usage (on a GPU box): python tools/gpu_diag.py [cfg ...]   writes gpurun_out/diag.txt
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libecp_b200 import capi, synth  # noqa: E402
from oracle.refbind import RefLib, have  # noqa: E402

CASES = {
    "cfg1": synth.cfg1, "cfg2": synth.cfg2, "au2": lambda: synth.cfg3(2), "au4": lambda: synth.cfg3(4),
    "cfg4a": lambda: synth.cfg4("a"), "cfg4b": lambda: synth.cfg4("b"), "S3": lambda: synth.probe(3, 3, 4),
    "cfg5s": lambda: synth.cfg5(24),
}


def main():
    names = sys.argv[1:] or ["cfg1", "cfg2", "au2", "cfg4a"]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    out = open(os.path.join(ROOT, "gpurun_out", "diag.txt"), "a")

    def P(*a):
        s = " ".join(str(x) for x in a)
        print(s, flush=True)
        out.write(s + "\n")
        out.flush()

    orc = RefLib("ref" if have("ref") else "port")
    P("oracle kind:", orc.kind)
    for nm in names:
        s = CASES[nm]()
        t0 = time.time()
        rc_o, ro = orc.callbacks(s)
        t_o = time.time() - t0
        t0 = time.time()
        with capi.Handle(s) as h:
            t_init = time.time() - t0
            t0 = time.time()
            rc_g, rg = h.callbacks()
            t_g = time.time() - t0
            st = h.stats()
            t0 = time.time()
            rc_m, M = h.integrals_host()
            t_m = time.time() - t0
            st2 = h.stats()
        P(f"== {nm}: oracle rc={rc_o} {len(ro)} cbs {t_o:.2f}s | gpu rc={rc_g} {len(rg)} cbs init {t_init:.3f}s cb-run {t_g:.3f}s matrix-run {t_m:.3f}s")
        P("   stats:", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in st2.items()})
        if len(ro) != len(rg):
            P("   CALLBACK COUNT MISMATCH")
            continue
        worst = {0: (0.0, None), 1: (0.0, None)}
        nviol = {0: 0, 1: 0}
        nel = {0: 0, 1: 0}
        bad_examples = []
        for k, (a, b) in enumerate(zip(ro, rg)):
            if a[:9] != b[:9]:
                P("   KEY MISMATCH at", k, a[:9], b[:9])
                break
            ty = k % 2
            d = np.abs(a[9] - b[9])
            tol = 1e-12 + 1e-10 * np.abs(a[9])
            bad = d > tol
            nb_ = int(bad.sum() + np.isnan(b[9]).sum())
            nviol[ty] += nb_
            nel[ty] += d.size
            m = float(np.nanmax(d)) if d.size else 0.0
            if np.isnan(b[9]).any():
                m = float("inf")
            if m > worst[ty][0]:
                worst[ty] = (m, a[:9])
            if nb_ and len(bad_examples) < 6:
                i = int(np.argmax(np.where(np.isnan(d), np.inf, d)))
                bad_examples.append((ty + 1, a[:9], i, float(a[9][i]), float(b[9][i])))
        for ty in (0, 1):
            P(f"   type{ty + 1}: elements {nel[ty]} violations {nviol[ty]} max|d| {worst[ty][0]:.3e} at {worst[ty][1]}")
        for e in bad_examples:
            P("   bad:", e)
        # matrix path vs accumulated oracle
        dim = s["dim"]
        Io = orc.get_integrals(s)
        d = np.abs(M - Io)
        tol = 1e-12 + 1e-10 * np.abs(Io)
        P(f"   matrix: rc={rc_m} max|d| {d.max():.3e} violations {(d > tol).sum()} of {dim * dim}; lower-tri nonzero {(np.tril(M, -1) != 0).sum()}")
    out.close()


if __name__ == "__main__":
    main()
