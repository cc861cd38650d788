"""GPU tier (-m gpu): the per-point device functions of the kernels, run ON THE GPU through the C ABI
(libecp_b200_debug_unit), against the oracle's unit accessors - SURVEY section 4's per-primitive plan:
Bessel z-sweep over the three branches (reference src/bessel.c:105,136), real spherical harmonics
(src/spherical_harmonics.c:15-114), PS93 on recorded windowed integrand tables (src/gc_integrators.c:156-217),
and the device intermediates F / T / gamma / chi / Q of a whole run (libecp_b200_debug_fetch) against each other
across kernel variants."""
import ctypes as C

import numpy as np
import pytest

from libecp_b200 import capi, synth
from test_host import Oracle, _p, _pd, _pi

pytestmark = pytest.mark.gpu


def test_device_bessel_matches_oracle_on_all_branches():
    s = synth.cfg4("b")
    o = Oracle(s)
    rng = np.random.default_rng(7)
    zs = np.concatenate([[0.0, -1.0, 1e-9, 9.9e-8, 1e-7, 1.00001e-7, 0.005, 15.995, 15.999999, 16.0, 16.00001, 50.0, 1e3, 1e5],
                         rng.uniform(0, 16, 3000), 10 ** rng.uniform(-9, 4, 1000)])
    with capi.Handle(s) as h:
        for lmax in (0, 3, 6, 10):
            got = h.unit("bessel", len(zs), zs, [lmax], len(zs) * (lmax + 1)).reshape(len(zs), lmax + 1)
            ref = np.zeros((len(zs), lmax + 1))
            for i, z in enumerate(zs):
                o.L.oracle_bessel(C.c_void_p(o.h), lmax, float(z), _p(ref[i], _pd))
            taylor = (zs >= 1e-7) & (zs < 16.0)
            # small-z branch: products and exact small-integer divisions only -> within an ulp or two; asymptotic branch
            # (z >= 16): the alternating sum R_l(-z) cancels (sum of |terms| / |result| up to ~1e3 at l = 10, z = 16)
            # and the device contracts K += f A[i] into an fma -> relative 1e-12; Taylor branch: fma in the derivative
            # recurrence (ecp_math.h): a few ulp, or < 1e-24 absolute for the tiny high orders at small z
            small, asym = zs < 1e-7, zs >= 16.0
            assert np.all(np.abs(got[small] - ref[small]) <= 2 * np.spacing(np.abs(ref[small]))), lmax
            assert np.all(np.abs(got[asym] - ref[asym]) <= 2e-12 * np.abs(ref[asym])), lmax
            assert np.all(np.abs(got[taylor] - ref[taylor]) <= 8 * np.spacing(np.abs(ref[taylor])) + 1e-24), lmax
            exact = np.all(np.abs(got[taylor] - ref[taylor]) <= np.spacing(np.abs(ref[taylor])), axis=1)
            assert exact.mean() > 0.9
    o.close()


def test_device_spherical_harmonics_match_oracle():
    s = synth.cfg4("b")
    o = Oracle(s)
    rng = np.random.default_rng(11)
    ang = [(0, 0), (np.pi, 0), (np.arccos(0.0), 0.5 * np.pi), (1.0, 1.5 * np.pi), (0.3, -1.2)]
    ang += [(rng.uniform(0, np.pi), rng.uniform(-1.5, 4.7)) for _ in range(500)]
    ang = np.array(ang, np.float64)
    with capi.Handle(s) as h:
        for lmax in (0, 1, 6, 10):
            n2 = (lmax + 1) ** 2
            got = h.unit("rsh", len(ang), ang.ravel(), [lmax], len(ang) * n2).reshape(len(ang), n2)
            ref = np.zeros((len(ang), n2))
            for i, (th, ph) in enumerate(ang):
                o.L.oracle_rsh(C.c_void_p(o.h), lmax, float(th), float(ph), _p(ref[i], _pd))
            # device sin / cos / sqrt / acos differ from glibc by <= 1-2 ulp; the recurrences amplify that mildly
            assert np.all(np.abs(got - ref) <= 1e-13 * (1 + np.abs(ref))), lmax
    o.close()


def test_device_ps93_on_windowed_tables_matches_oracle():
    """value, return code and number of evaluated points of the slot-ordered PS93 (the fast-path quadrature) on the GPU
    equal integrateGC_PS93 on the same integrand tables, full and windowed"""
    s = synth.cfg2()
    o = Oracle(s)
    rng = np.random.default_rng(3)
    with capi.Handle(s) as h:
        x = h.host_table("small_x")
        oidx = h.host_itable("small_oidx").astype(np.int32)
        m = oidx >= 0

        def perm(v):
            out = np.zeros(384)
            out[m] = v[oidx[m]]
            return out

        rows, wins, ref = [], [], []
        for trial in range(300):
            a, c = rng.uniform(0.05, 30), rng.uniform(0, 6)
            st = int(rng.integers(0, 200))
            en = int(rng.integers(st, 383))
            if trial % 3 == 0:
                st, en = 0, 382
            Fa = np.exp(-a * (x - c) ** 2)
            Fb = 1.0 / (1 + x)
            U = x ** rng.integers(0, 4)
            Fa[:st] = 0
            Fa[en:] = 0
            r1, n1 = np.zeros(1), np.zeros(1, np.int32)
            rc1 = o.L.oracle_ps93_table(C.c_void_p(o.h), _p(Fa * Fb * U, _pd), st, en, _p(r1, _pd), _p(n1, _pi))
            rows += [perm(Fa), perm(Fb), perm(U)]
            wins += [st, en]
            ref.append((r1[0], rc1, n1[0]))
        got = h.unit("ps93", 300, np.concatenate(rows), wins, 900).reshape(300, 3)
    ref = np.array(ref)
    assert np.array_equal(got[:, 1], ref[:, 1]) and np.array_equal(got[:, 2], ref[:, 2])
    ok = ref[:, 1] == 0
    assert 0 < (~ok).sum() < 300
    # same products in the same order; the compiler may contract a*b+c into an fma on the device (ecp_math.h header)
    assert np.all(np.abs(got[ok, 0] - ref[ok, 0]) <= 1e-15 * np.abs(ref[ok, 0]) + 1e-300)
    o.close()


def test_device_potential_matches_host_tables():
    """evalECP on the device (large-grid kernels) vs the host-built small-grid potential table of the same handle
    (typeUL = U_L(r_n), reference src/libecp.c:269-270)"""
    s = synth.cfg2()
    with capi.Handle(s) as h:
        r = h.host_table("small_rs")
        UL = h.host_table("typeUL")[:384]
        L = int(h.host_itable("dims")[0])
        got = h.unit("pot", 384, r, [0, L], 384)
        oidx = h.host_itable("small_oidx")
        live = (oidx >= 0) & (UL != 0.0)  # the table is cut where the potential falls below the accuracy
        assert live.sum() > 100
        assert np.all(np.abs(got[live] - UL[live]) <= 4e-16 * np.abs(UL[live]) + 1e-300)


def test_device_intermediates_are_finite_and_consistent():
    """F, T, gamma, chi, Q of the last batch (libecp_b200_debug_fetch): finite, non-trivial, F zero outside the shell
    windows (k_Ftab2 tabulates the window only), and identical between two runs of the same handle"""
    s = synth.cfg3(4)
    with capi.Handle(s) as h:
        rc, M = h.integrals_host()
        first = {k: h.debug_fetch(k, 300000) for k in ("F", "T", "gamma", "chi", "Q")}
        rc2, M2 = h.integrals_host()
        second = {k: h.debug_fetch(k, 300000) for k in ("F", "T", "gamma", "chi", "Q")}
    assert rc == rc2 == 0
    for k, v in first.items():
        assert np.all(np.isfinite(v)) and np.count_nonzero(v) > 0, k
        assert np.array_equal(v, second[k]), k
    assert (first["F"] == 0).mean() > 0.3
