"""ctypes binding of libecp_b200.so - the same C ABI an existing libECP caller links against
(include/libecp.h, getIntegrals.h, libecp_b200.h).  Python is only the harness language for tests and
bench.py; there is no Python compute path and no CPU fallback: if the CUDA library is missing or no
device is usable, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# LIBECP_B200_SO: another build of the same library (A/B of compiler flags, tools/); never a different code path
SO_PATH = os.environ.get("LIBECP_B200_SO") or os.path.join(HERE, "lib", "libecp_b200.so")

_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int)
CALLBACK = C.CFUNCTYPE(None, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                       _pd, C.c_void_p)

EXPORTS = [
    "libECP_init", "calculateECPIntegrals", "libECP_free", "getIntegrals", "cartesianShellOrder",
    "cartesianShellOrderIndex", "libecp_b200_set_device", "libecp_b200_set_shard", "libecp_b200_pair_owner",
    "libecp_b200_integrals_device", "libecp_b200_integrals_host", "libecp_b200_get_stats", "libecp_b200_screening",
    "libecp_b200_host_table", "libecp_b200_host_itable", "libecp_b200_triple_list", "libecp_b200_set_tables_only",
    "libecp_b200_debug_fetch", "libecp_b200_fp64_peak", "libecp_b200_last_error", "libecp_b200_set_host_threads",
    "libecp_b200_set_serial_kernels", "libecp_b200_release_cache", "libecp_b200_build_only", "libecp_b200_owned_rows", "libecp_b200_pack_rows", "libecp_b200_unpack_rows",
    "libecp_b200_matrix_ptr", "libecp_b200_comm_unique_id", "libecp_b200_comm_init", "libecp_b200_comm_attach",
    "libecp_b200_allgather", "libecp_b200_device_sync", "libecp_b200_comm_free", "libecp_b200_callback_keys",
    "libecp_b200_debug_unit", "libecp_b200_spherical_dim", "libecp_b200_spherical_device", "libecp_b200_spherical_host",
]


class Stats(C.Structure):
    _fields_ = [(n, C.c_longlong) for n in (
        "nominal_triples", "executed_triples", "shell_slots", "atom_slots", "prim_pairs", "fast_quadratures",
        "fast_failed", "fallback_items", "type1_fallback_pairs", "stale_centre_events", "kernel_launches",
        "batches", "h2d_bytes", "d2h_bytes", "tables_h2d_bytes")] + [(n, C.c_double) for n in (
            "ms_build", "ms_tables", "ms_fastT", "ms_fallback", "ms_link", "ms_type1", "ms_chi", "ms_shift",
            "ms_device_total")]

    def asdict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


_lib = None


def lib():
    """Load the product library; fail loudly if it was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError(f"{SO_PATH} is missing - run `python -m libecp_b200.build` (no CPU fallback exists)")
        L = C.CDLL(SO_PATH)
        L.libECP_init.restype = C.c_void_p
        L.calculateECPIntegrals.restype = C.c_int
        L.calculateECPIntegrals.argtypes = [C.c_void_p, CALLBACK, C.c_void_p]
        L.libECP_free.argtypes = [C.c_void_p]
        L.libECP_free.restype = None
        L.getIntegrals.restype = C.c_int
        L.libecp_b200_set_device.argtypes = [C.c_int]
        L.libecp_b200_set_shard.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.libecp_b200_pair_owner.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.libecp_b200_integrals_device.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), _pi]
        L.libecp_b200_integrals_host.argtypes = [C.c_void_p, C.c_int, _pd]
        L.libecp_b200_get_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
        L.libecp_b200_screening.argtypes = [C.c_void_p, C.c_int, _pi, _pi, _pi, _pi]
        L.libecp_b200_host_table.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(_pd)]
        L.libecp_b200_host_itable.argtypes = [C.c_void_p, C.c_char_p, _pi, C.c_int]
        L.libecp_b200_triple_list.argtypes = [C.c_void_p, _pi, C.c_longlong]
        L.libecp_b200_triple_list.restype = C.c_longlong
        L.libecp_b200_set_tables_only.argtypes = [C.c_int]
        L.libecp_b200_debug_fetch.argtypes = [C.c_void_p, C.c_char_p, _pd, C.c_longlong]
        L.libecp_b200_fp64_peak.argtypes = [C.c_int, C.c_int]
        L.libecp_b200_fp64_peak.restype = C.c_double
        L.libecp_b200_last_error.restype = C.c_char_p
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t)


def set_device(dev: int) -> None:
    lib().libecp_b200_set_device(int(dev))


def set_host_threads(n: int) -> None:
    L = lib()
    L.libecp_b200_set_host_threads.argtypes = [C.c_int]
    L.libecp_b200_set_host_threads(int(n))


def fp64_peak(dev: int = 0, iters: int = 200000) -> float:
    return float(lib().libecp_b200_fp64_peak(int(dev), int(iters)))


def comm_unique_id() -> bytes:
    """128-byte NCCL unique id (rank 0 creates it, the caller distributes it to all ranks)"""
    buf = C.create_string_buffer(128)
    if lib().libecp_b200_comm_unique_id(buf):
        raise RuntimeError("comm_unique_id: " + (lib().libecp_b200_last_error() or b"").decode())
    return buf.raw


def get_integrals(s, tol=1e-12, acc=1e-14, large=1024):
    """The one-call interface (reference src/getIntegrals.h:7-13) on host buffers; returns the nAO x nAO matrix."""
    dim = int(s["dim"])
    I = np.zeros((dim, dim), dtype=np.float64)
    rc = lib().getIntegrals(C.c_int(s["nat"]), _p(s["geometry"], _pd),
                            _p(s["shellsECP"], _pi), _p(s["KECP"], _pi), _p(s["lECP"], _pi),
                            _p(s["nECP"], _pd), _p(s["dECP"], _pd), _p(s["aECP"], _pd),
                            _p(s["shellsBS"], _pi), _p(s["lBS"], _pi), _p(s["KBS"], _pi),
                            _p(s["dBS"], _pd), _p(s["aBS"], _pd),
                            C.c_int(large), C.c_double(tol), C.c_double(acc), C.c_int(dim), _p(I, _pd))
    if rc != 0:
        raise RuntimeError("getIntegrals failed: " + (lib().libecp_b200_last_error() or b"").decode())
    return I


class Handle:
    """libECP_init / calculateECPIntegrals / libECP_free (reference src/libecp.h:15-29) + device extensions."""

    def __init__(self, s, tol=1e-12, acc=1e-14, large=1024, n=0, tables_only=False, ordering=None, lmax=-1):
        """ordering: int32 array in the layout of cartesianShellOrder(lmax) - the caller's Cartesian component order
        (reference src/libecp.c:152-166); n: derivative order (0, 1 or 2)"""
        L = lib()
        self.n = n
        self.s = s  # keeps the borrowed arrays alive (the handle borrows them, reference src/libecp.c:68-74)
        self.ordering = None if ordering is None else np.ascontiguousarray(ordering, np.int32)
        if tables_only:
            L.libecp_b200_set_tables_only(1)
        try:
            self.h = L.libECP_init(C.c_int(s["nat"]), _p(s["geometry"], _pd),
                                   _p(s["shellsECP"], _pi), _p(s["lECP"], _pi), _p(s["KECP"], _pi),
                                   _p(s["nECP"], _pd), _p(s["dECP"], _pd), _p(s["aECP"], _pd),
                                   _p(s["shellsBS"], _pi), _p(s["lBS"], _pi), _p(s["KBS"], _pi),
                                   _p(s["dBS"], _pd), _p(s["aBS"], _pd),
                                   C.c_int(n), C.c_int(lmax), None if self.ordering is None else _p(self.ordering, _pi),
                                   C.c_int(large), C.c_double(tol), C.c_double(acc))
        finally:
            if tables_only:
                L.libecp_b200_set_tables_only(0)
        if not self.h:
            raise RuntimeError("libECP_init returned NULL: " + (L.libecp_b200_last_error() or b"").decode())

    def close(self):
        if getattr(self, "h", None):
            lib().libECP_free(C.c_void_p(self.h))
            self.h = None

    def __del__(self):
        self.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_shard(self, rank, world):
        lib().libecp_b200_set_shard(C.c_void_p(self.h), int(rank), int(world))

    def callbacks(self, keep_blocks=True):
        recs = []

        order = self.n

        def cb(A, s1, la, sha, B, s2, lb, shb, Cc, I, args):
            # second derivatives, shifts (+1,0) / (0,+1): blocks at the unshifted momenta (reference src/libecp.c:362-369)
            ea, eb = (0, 0) if order == 2 and (sha, shb) in ((1, 0), (0, 1)) else (sha, shb)
            n = ((la + ea + 1) * (la + ea + 2) // 2) * ((lb + eb + 1) * (lb + eb + 2) // 2)
            blk = np.ctypeslib.as_array(I, shape=(n,)).copy() if keep_blocks else None
            recs.append((A, s1, la, sha, B, s2, lb, shb, Cc, blk))

        rc = lib().calculateECPIntegrals(C.c_void_p(self.h), CALLBACK(cb), None)
        return rc, recs

    def integrals_host(self):
        dim = int(self.s["dim"])
        I = np.zeros((dim, dim), dtype=np.float64)
        rc = lib().libecp_b200_integrals_host(C.c_void_p(self.h), dim, _p(I, _pd))
        if rc < 0:
            raise RuntimeError("device failure: " + (lib().libecp_b200_last_error() or b"").decode())
        return rc, I

    def integrals_device(self):
        """Returns (rc, device pointer as int, nAO); the matrix stays in HBM (owned by the handle)."""
        ptr = C.c_void_p()
        n = C.c_int()
        rc = lib().libecp_b200_integrals_device(C.c_void_p(self.h), C.byref(ptr), C.byref(n))
        if rc < 0:
            raise RuntimeError("device failure: " + (lib().libecp_b200_last_error() or b"").decode())
        return rc, ptr.value, n.value

    def matrix_ptr(self):
        """device pointer of the handle's resident result matrix (0 before the first pass)"""
        f = lib().libecp_b200_matrix_ptr
        f.restype = C.c_void_p
        f.argtypes = [C.c_void_p]
        return f(C.c_void_p(self.h)) or 0

    def comm_init(self, rank, world, unique_id: bytes):
        """collective: communicator inside the library from the 128-byte NCCL unique id (see comm_unique_id)"""
        f = lib().libecp_b200_comm_init
        f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p]
        if f(C.c_void_p(self.h), rank, world, unique_id):
            raise RuntimeError("comm_init: " + (lib().libecp_b200_last_error() or b"").decode())

    def allgather(self):
        """after integrals_device() on every rank: NCCL all-gather of the shards; returns bytes received"""
        f = lib().libecp_b200_allgather
        f.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
        n = C.c_longlong(0)
        if f(C.c_void_p(self.h), C.byref(n)):
            raise RuntimeError("allgather: " + (lib().libecp_b200_last_error() or b"").decode())
        return int(n.value)

    def device_sync(self):
        f = lib().libecp_b200_device_sync
        f.argtypes = [C.c_void_p]
        return f(C.c_void_p(self.h))

    def set_serial_kernels(self, on=True):
        L = lib()
        L.libecp_b200_set_serial_kernels.argtypes = [C.c_void_p, C.c_int]
        L.libecp_b200_set_serial_kernels(C.c_void_p(self.h), 1 if on else 0)

    def stats(self):
        st = Stats()
        lib().libecp_b200_get_stats(C.c_void_p(self.h), C.byref(st))
        return st.asdict()

    def host_table(self, name):
        ptr = _pd()
        n = lib().libecp_b200_host_table(C.c_void_p(self.h), name.encode(), C.byref(ptr))
        return np.ctypeslib.as_array(ptr, shape=(n,)).copy() if n else np.zeros(0)

    def host_itable(self, name):
        n = lib().libecp_b200_host_itable(C.c_void_p(self.h), name.encode(), None, 0)
        out = np.zeros(max(n, 1), np.int32)
        lib().libecp_b200_host_itable(C.c_void_p(self.h), name.encode(), _p(out, _pi), n)
        return out[:n]

    def screening(self, centre, L):
        ns = int(self.s["nshells"])
        endl = np.zeros(max(L, 1), np.int32)
        st, en, sk = (np.zeros(ns, np.int32) for _ in range(3))
        rc = lib().libecp_b200_screening(C.c_void_p(self.h), int(centre), _p(endl, _pi), _p(st, _pi), _p(en, _pi),
                                         _p(sk, _pi))
        if rc:
            raise RuntimeError("not an ECP centre")
        return endl[:L], st, en, sk

    def unit(self, what, n, inp, ipar, nout):
        """tests: run a per-point device function on the GPU (include/libecp_b200.h: libecp_b200_debug_unit)"""
        f = lib().libecp_b200_debug_unit
        f.argtypes = [C.c_void_p, C.c_char_p, C.c_int, _pd, C.c_longlong, _pi, C.c_int, _pd, C.c_longlong]
        inp = np.ascontiguousarray(inp, np.float64)
        ipar = np.ascontiguousarray(ipar, np.int32)
        out = np.zeros(nout)
        if f(C.c_void_p(self.h), what.encode(), int(n), _p(inp, _pd), inp.size, _p(ipar, _pi), ipar.size, _p(out, _pd), nout):
            raise RuntimeError("debug_unit failed: " + (lib().libecp_b200_last_error() or b"").decode())
        return out

    def spherical_host(self):
        """S = C^T M C of an n = 0 run in pure spherical-harmonic functions (upper triangle), on host memory"""
        L = lib()
        L.libecp_b200_spherical_dim.argtypes = [C.c_void_p]
        n = int(L.libecp_b200_spherical_dim(C.c_void_p(self.h)))
        S = np.zeros((n, n))
        L.libecp_b200_spherical_host.argtypes = [C.c_void_p, C.c_int, _pd]
        rc = L.libecp_b200_spherical_host(C.c_void_p(self.h), n, _p(S, _pd))
        if rc < 0:
            raise RuntimeError("spherical_host: " + (L.libecp_b200_last_error() or b"").decode())
        return rc, S

    def callback_keys(self):
        """(A,s1,la,shifta,B,s2,lb,shiftb,C) of every executed (shifted) triple in call order (host only)"""
        f = lib().libecp_b200_callback_keys
        f.restype = C.c_longlong
        f.argtypes = [C.c_void_p, _pi, C.c_longlong]
        n = int(f(C.c_void_p(self.h), None, 0))
        out = np.zeros((max(n, 1), 9), np.int32)
        f(C.c_void_p(self.h), _p(out, _pi), n)
        return out[:n]

    def owned_rows(self, rank, world):
        """AO rows (ascending) of the shells whose shell-pair rows `rank` of `world` owns"""
        f = lib().libecp_b200_owned_rows
        f.restype = C.c_longlong
        f.argtypes = [C.c_void_p, C.c_int, C.c_int, _pi, C.c_longlong]
        n = f(C.c_void_p(self.h), rank, world, None, 0)
        out = np.zeros(max(int(n), 1), np.int32)
        f(C.c_void_p(self.h), rank, world, _p(out, _pi), n)
        return out[:n]

    def packed_size(self, rows):
        """doubles that the upper-triangle parts of `rows` take"""
        n = int(self.s["dim"])
        return int((n - np.asarray(rows, np.int64)).sum())

    def pack_rows(self, rows, dev_ptr, cap):
        f = lib().libecp_b200_pack_rows
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, _pi, C.c_longlong, C.c_void_p, C.c_longlong, C.POINTER(C.c_longlong)]
        rows = np.ascontiguousarray(rows, np.int32)
        e = C.c_longlong(0)
        if f(C.c_void_p(self.h), _p(rows, _pi), len(rows), C.c_void_p(dev_ptr), cap, C.byref(e)):
            raise RuntimeError("pack_rows: " + lib().libecp_b200_last_error().decode())
        return int(e.value)

    def unpack_rows(self, rows, dev_ptr, cap):
        f = lib().libecp_b200_unpack_rows
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, _pi, C.c_longlong, C.c_void_p, C.c_longlong]
        rows = np.ascontiguousarray(rows, np.int32)
        if f(C.c_void_p(self.h), _p(rows, _pi), len(rows), C.c_void_p(dev_ptr), cap):
            raise RuntimeError("unpack_rows: " + lib().libecp_b200_last_error().decode())

    def build_only(self):
        """(ms, triples, batches) of the host batch builder alone"""
        f = lib().libecp_b200_build_only
        f.restype = C.c_double
        f.argtypes = [C.c_void_p, C.POINTER(C.c_longlong), C.POINTER(C.c_int)]
        n, nb = C.c_longlong(0), C.c_int(0)
        ms = f(C.c_void_p(self.h), C.byref(n), C.byref(nb))
        return float(ms), int(n.value), int(nb.value)

    def triple_list(self):
        n = lib().libecp_b200_triple_list(C.c_void_p(self.h), None, 0)
        out = np.zeros((max(n, 1), 7), np.int32)
        lib().libecp_b200_triple_list(C.c_void_p(self.h), _p(out, _pi), n)
        return out[:n]

    def debug_fetch(self, what, n):
        out = np.zeros(n, np.float64)
        rc = lib().libecp_b200_debug_fetch(C.c_void_p(self.h), what.encode(), _p(out, _pd), n)
        if rc:
            raise RuntimeError("debug_fetch failed")
        return out
