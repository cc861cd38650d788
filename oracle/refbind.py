"""ctypes bindings for the two CPU checkers (TEST INFRASTRUCTURE ONLY - never imported by the product).

* ``RefLib("ref")``    -> oracle/_ref/libecp_ref.so : the unmodified reference compiled by oracle/Makefile
* ``RefLib("port")``   -> oracle/liboracle_ecp.so   : our plain-C restatement (oracle_ecp.c)

Both export the reference's public C API (src/libecp.h:15-29, src/getIntegrals.h:7-13); the port
prefixes its symbols with ``oracle_``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libecp_ref.so")
PORT_SO = os.path.join(HERE, "liboracle_ecp.so")

CALLBACK = C.CFUNCTYPE(None, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                       C.POINTER(C.c_double), C.c_void_p)

_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int)


def _p(a, t):
    return a.ctypes.data_as(t)


def have(kind: str) -> bool:
    return os.path.exists(REF_SO if kind == "ref" else PORT_SO)


class RefLib:
    def __init__(self, kind: str = "ref"):
        self.kind = kind
        self.lib = C.CDLL(REF_SO if kind == "ref" else PORT_SO)
        pre = "" if kind == "ref" else "oracle_"
        self.f_get = getattr(self.lib, pre + "getIntegrals")
        self.f_get.restype = C.c_int
        self.f_init = getattr(self.lib, pre + "libECP_init")
        self.f_init.restype = C.c_void_p
        self.f_calc = getattr(self.lib, pre + "calculateECPIntegrals")
        self.f_calc.restype = C.c_int
        self.f_free = getattr(self.lib, pre + "libECP_free")
        self.f_free.restype = None

    def get_integrals(self, s, tol=1e-12, acc=1e-14, large=1024):
        dim = int(s["dim"])
        I = np.zeros((dim, dim), dtype=np.float64)
        rc = self.f_get(C.c_int(s["nat"]), _p(s["geometry"], _pd),
                        _p(s["shellsECP"], _pi), _p(s["KECP"], _pi), _p(s["lECP"], _pi),
                        _p(s["nECP"], _pd), _p(s["dECP"], _pd), _p(s["aECP"], _pd),
                        _p(s["shellsBS"], _pi), _p(s["lBS"], _pi), _p(s["KBS"], _pi),
                        _p(s["dBS"], _pd), _p(s["aBS"], _pd),
                        C.c_int(large), C.c_double(tol), C.c_double(acc), C.c_int(dim), _p(I, _pd))
        if rc != 0:
            raise RuntimeError(f"getIntegrals rc={rc}")
        return I

    def callbacks(self, s, tol=1e-12, acc=1e-14, large=1024, keep_blocks=True, n=0, ordering=None, lmax=-1):
        """Run init/calculate/free with a recording callback.

        Returns (rc, records) with records = list of (A,s1,la,shifta,B,s2,lb,shiftb,C,block ndarray|None)
        in the call order of src/libecp.c:372.  n = derivative order (only the compiled reference, kind "ref",
        implements n > 0: shifted-momentum blocks, src/libecp.c:246-250,322-369).
        """
        recs = []
        if ordering is not None:  # ordering: int32 array laid out like cartesianShellOrder(lmax) (src/libecp.c:152-166)
            ordering = np.ascontiguousarray(ordering, np.int32)

        order = n

        def cb(A, s1, la, sha, B, s2, lb, shb, Cc, I, args):
            # second derivatives, shifts (+1,0) / (0,+1): momentum reset before calcPolynomials (src/libecp.c:362-369)
            ea, eb = (0, 0) if order == 2 and (sha, shb) in ((1, 0), (0, 1)) else (sha, shb)
            n = ((la + ea + 1) * (la + ea + 2) // 2) * ((lb + eb + 1) * (lb + eb + 2) // 2)
            blk = np.ctypeslib.as_array(I, shape=(n,)).copy() if keep_blocks else None
            recs.append((A, s1, la, sha, B, s2, lb, shb, Cc, blk))

        cbf = CALLBACK(cb)
        h = self.f_init(C.c_int(s["nat"]), _p(s["geometry"], _pd),
                        _p(s["shellsECP"], _pi), _p(s["lECP"], _pi), _p(s["KECP"], _pi),
                        _p(s["nECP"], _pd), _p(s["dECP"], _pd), _p(s["aECP"], _pd),
                        _p(s["shellsBS"], _pi), _p(s["lBS"], _pi), _p(s["KBS"], _pi),
                        _p(s["dBS"], _pd), _p(s["aBS"], _pd),
                        C.c_int(n), C.c_int(lmax), None if ordering is None else _p(ordering, _pi),
                        C.c_int(large), C.c_double(tol), C.c_double(acc))
        if not h:
            raise RuntimeError("libECP_init returned NULL")
        rc = self.f_calc(C.c_void_p(h), cbf, None)
        self.f_free(C.c_void_p(h))
        return rc, recs


COUNTER_NAMES = ("triples_exec callbacks ps93_calls ps93_fail psm92_calls psm92_fail P_T_used P_T_all P_Q2_used "
                 "P_Q2_all P_Q1 tab2_touched tab2_tabulated tab1s_touched tab1s_tabulated tab1l_touched "
                 "tab1l_tabulated flops_tab flops_Ftab M_link M_chi M_poly T_quads_used T_quads_all fb_pairs "
                 "t1_pairs stale_center_hits flops_tab2 flops_tab1").split()


def port_counters(s, tol=1e-12, acc=1e-14, large=1024, stale=1):
    """Run the instrumented restatement (oracle_ecp.c) and return (rc, matrix, counters dict).

    Algorithmic flops (touched/used work only, FMA = 2):
      fastT    = 4 P_T_used
      fallback = 7 P_Q2_used + flops_tab2
      type1    = 6 P_Q1 + flops_tab1
      link     = 2 M_link ; chi = 2 M_chi ; shift = 2 M_poly ; tables = flops_Ftab
    """
    p = RefLib("port")
    L = p.lib
    L.oracle_set_stale_buffers.argtypes = [C.c_void_p, C.c_int]
    L.oracle_counters.argtypes = [C.c_void_p, _pd, C.c_int]
    dim = int(s["dim"])
    ao = np.zeros(int(s["nshells"]) + 1, np.int64)
    ao[1:] = np.cumsum([(l + 1) * (l + 2) // 2 for l in s["lBS"]])
    first = np.zeros(int(s["nat"]) + 1, np.int64)
    first[1:] = np.cumsum(s["shellsBS"])
    M = np.zeros((dim, dim))

    def cb(A, s1, la, sha, B, s2, lb, shb, Cc, I, args):
        na, nb = (la + 1) * (la + 2) // 2, (lb + 1) * (lb + 2) // 2
        blk = np.ctypeslib.as_array(I, shape=(na, nb))
        a0, b0 = ao[first[A] + s1], ao[first[B] + s2]
        sub = M[a0:a0 + na, b0:b0 + nb]
        if a0 == b0:
            sub += np.triu(blk)
        else:
            sub += blk

    cbf = CALLBACK(cb)
    h = p.f_init(C.c_int(s["nat"]), _p(s["geometry"], _pd), _p(s["shellsECP"], _pi), _p(s["lECP"], _pi),
                 _p(s["KECP"], _pi), _p(s["nECP"], _pd), _p(s["dECP"], _pd), _p(s["aECP"], _pd),
                 _p(s["shellsBS"], _pi), _p(s["lBS"], _pi), _p(s["KBS"], _pi), _p(s["dBS"], _pd), _p(s["aBS"], _pd),
                 C.c_int(0), C.c_int(-1), None, C.c_int(large), C.c_double(tol), C.c_double(acc))
    L.oracle_set_stale_buffers(C.c_void_p(h), stale)
    rc = p.f_calc(C.c_void_p(h), cbf, None)
    cnt = np.zeros(len(COUNTER_NAMES))
    L.oracle_counters(C.c_void_p(h), _p(cnt, _pd), len(COUNTER_NAMES))
    p.f_free(C.c_void_p(h))
    return rc, M, dict(zip(COUNTER_NAMES, cnt.tolist()))


def algorithmic_flops(c):
    """Per-kernel algorithmic flops from a counters dict (see port_counters)."""
    k = {
        "tables": c["flops_Ftab"], "fastT": 4 * c["P_T_used"], "fallback": 7 * c["P_Q2_used"] + c["flops_tab2"],
        "type1": 6 * c["P_Q1"] + c["flops_tab1"], "link": 2 * c["M_link"], "chi": 2 * c["M_chi"],
        "shift": 2 * c["M_poly"],
    }
    k["total"] = sum(k.values())
    return k
