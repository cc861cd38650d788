"""ctypes binding of the input loaders (include/libecp_b200_io.h; reference example/ex1.c:11-123): text files -> the
array dictionary the tests and bench.py pass to the C API (same keys as libecp_b200.synth)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi

INDEXED, SHIPPED = 0, 1
_pi, _pd = C.POINTER(C.c_int), C.POINTER(C.c_double)


def _take(L, ptr, n, dtype):
    out = np.ctypeslib.as_array(ptr, shape=(max(int(n), 1),))[: int(n)].astype(dtype).copy()
    L.libecp_io_free(ptr)
    return out


def load(xyz, ecp, bs, ecp_format=INDEXED, name=None):
    L = capi.lib()
    L.libecp_io_free.argtypes = [C.c_void_p]
    L.libecp_io_load_xyz.argtypes = [C.c_char_p, _pi, C.POINTER(_pd)]
    L.libecp_io_load_bs.argtypes = [C.c_char_p, C.c_int] + [C.POINTER(_pi)] * 3 + [C.POINTER(_pd)] * 2 + [_pi]
    L.libecp_io_load_ecp.argtypes = [C.c_char_p, C.c_int, C.c_int] + [C.POINTER(_pi)] * 3 + [C.POINTER(_pd)] * 3
    nat, geom = C.c_int(0), _pd()
    rc = L.libecp_io_load_xyz(str(xyz).encode(), C.byref(nat), C.byref(geom))
    if rc:
        raise OSError(f"load_xyz({xyz}) failed: {rc}")
    n = nat.value
    g = _take(L, C.cast(geom, _pd), 3 * n, np.float64)
    sh, l, k = _pi(), _pi(), _pi()
    a, d, p = _pd(), _pd(), _pd()
    nsh = C.c_int(0)
    rc = L.libecp_io_load_bs(str(bs).encode(), n, C.byref(sh), C.byref(l), C.byref(k), C.byref(a), C.byref(d), C.byref(nsh))
    if rc:
        raise OSError(f"load_bs({bs}) failed: {rc}")
    shells_bs = _take(L, sh, n, np.int32)
    ns = int(shells_bs.sum())
    l_bs, k_bs = _take(L, l, ns, np.int32), _take(L, k, ns, np.int32)
    npb = int(k_bs.sum())
    a_bs, d_bs = _take(L, a, npb, np.float64), _take(L, d, npb, np.float64)
    sh, l, k = _pi(), _pi(), _pi()
    a, d, p = _pd(), _pd(), _pd()
    rc = L.libecp_io_load_ecp(str(ecp).encode(), n, int(ecp_format), C.byref(sh), C.byref(l), C.byref(k), C.byref(a),
                              C.byref(d), C.byref(p))
    if rc:
        raise OSError(f"load_ecp({ecp}) failed: {rc}")
    shells_e = _take(L, sh, n, np.int32)
    ne = int(shells_e.sum())
    l_e, k_e = _take(L, l, ne, np.int32), _take(L, k, ne, np.int32)
    npe = int(k_e.sum())
    a_e, d_e, n_e = _take(L, a, npe, np.float64), _take(L, d, npe, np.float64), _take(L, p, npe, np.float64)
    L.libecp_io_ao_dim.argtypes = [C.c_int, _pi]
    dim = L.libecp_io_ao_dim(ns, l_bs.ctypes.data_as(_pi))
    return {"name": name or "loaded", "nat": n, "geometry": g, "shellsECP": shells_e, "lECP": l_e, "KECP": k_e, "nECP": n_e,
            "dECP": d_e, "aECP": a_e, "shellsBS": shells_bs, "lBS": l_bs, "KBS": k_bs, "dBS": d_bs, "aBS": a_bs,
            "dim": int(dim), "nshells": ns}
