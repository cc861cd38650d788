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

    def callbacks(self, s, tol=1e-12, acc=1e-14, large=1024, keep_blocks=True):
        """Run init/calculate/free with a recording callback.

        Returns (rc, records) with records = list of (A,s1,la,shifta,B,s2,lb,shiftb,C,block ndarray|None)
        in the call order of src/libecp.c:372.
        """
        recs = []

        def cb(A, s1, la, sha, B, s2, lb, shb, Cc, I, args):
            n = ((la + sha + 1) * (la + sha + 2) // 2) * ((lb + shb + 1) * (lb + shb + 2) // 2)
            blk = np.ctypeslib.as_array(I, shape=(n,)).copy() if keep_blocks else None
            recs.append((A, s1, la, sha, B, s2, lb, shb, Cc, blk))

        cbf = CALLBACK(cb)
        h = self.f_init(C.c_int(s["nat"]), _p(s["geometry"], _pd),
                        _p(s["shellsECP"], _pi), _p(s["lECP"], _pi), _p(s["KECP"], _pi),
                        _p(s["nECP"], _pd), _p(s["dECP"], _pd), _p(s["aECP"], _pd),
                        _p(s["shellsBS"], _pi), _p(s["lBS"], _pi), _p(s["KBS"], _pi),
                        _p(s["dBS"], _pd), _p(s["aBS"], _pd),
                        C.c_int(0), C.c_int(-1), None, C.c_int(large), C.c_double(tol), C.c_double(acc))
        if not h:
            raise RuntimeError("libECP_init returned NULL")
        rc = self.f_calc(C.c_void_p(h), cbf, None)
        self.f_free(C.c_void_p(h))
        return rc, recs
