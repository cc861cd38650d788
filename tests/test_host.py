"""CPU tier: host logic of the product (no compute calls - there is no CPU compute path):
C-ABI exports, bit-exact host tables, screening decisions, executed-triple list, shard partition,
and the per-thread device math compiled for the host (tests/hostcheck) against the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN
from libecp_b200 import build, capi, synth
from oracle.refbind import RefLib, _p, _pd, _pi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SHAPES = {"cfg1": synth.cfg1, "au4": lambda: synth.cfg3(4), "cfg4b": lambda: synth.cfg4("b"),
          "cfg5s": lambda: synth.cfg5(30)}


class Oracle:
    """thin wrapper over the restatement's unit-level accessors"""

    def __init__(self, s):
        self.p = RefLib("port")
        L = self.L = self.p.lib
        L.oracle_table.restype = _pd
        L.oracle_table.argtypes = [C.c_void_p, C.c_char_p, _pi]
        L.oracle_bessel.argtypes = [C.c_void_p, C.c_int, C.c_double, _pd]
        L.oracle_rsh.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, _pd]
        L.oracle_screening.argtypes = [C.c_void_p, C.c_int, _pi, _pi, _pi, _pi]
        L.oracle_ps93_table.argtypes = [C.c_void_p, _pd, C.c_int, C.c_int, _pd, _pi]
        self.s = s
        self.h = self.p.f_init(
            C.c_int(s["nat"]), _p(s["geometry"], _pd), _p(s["shellsECP"], _pi), _p(s["lECP"], _pi), _p(s["KECP"], _pi),
            _p(s["nECP"], _pd), _p(s["dECP"], _pd), _p(s["aECP"], _pd), _p(s["shellsBS"], _pi), _p(s["lBS"], _pi),
            _p(s["KBS"], _pi), _p(s["dBS"], _pd), _p(s["aBS"], _pd), C.c_int(0), C.c_int(-1), None, C.c_int(1024),
            C.c_double(1e-12), C.c_double(1e-14))
        assert self.h

    def table(self, name):
        n = C.c_int()
        ptr = self.L.oracle_table(C.c_void_p(self.h), name.encode(), C.byref(n))
        return np.ctypeslib.as_array(ptr, shape=(n.value,)).copy()

    def close(self):
        self.p.f_free(C.c_void_p(self.h))


def test_cabi_exports_every_declared_symbol():
    """the shared library loads without a GPU and exports every function include/*.h declares"""
    L = capi.lib()
    declared = set()
    for hdr in ("libecp.h", "getIntegrals.h", "dimensions.h", "libecp_b200.h", "libecp_b200_io.h"):
        txt = open(os.path.join(ROOT, "include", hdr)).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        for m in re.finditer(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", txt):
            name = m.group(1)
            if name.startswith(("libECP_", "libecp_b200_", "libecp_io_", "cartesianShellOrder")) or name in (
                    "calculateECPIntegrals", "getIntegrals"):
                declared.add(name)
    assert {"libECP_init", "calculateECPIntegrals", "libECP_free", "getIntegrals"} <= declared
    for name in declared:
        assert hasattr(L, name), name
    assert set(capi.EXPORTS) <= declared


def test_no_cpu_fallback_without_device():
    """a tables-only handle refuses every compute entry point"""
    import torch

    s = synth.cfg2()
    with capi.Handle(s, tables_only=True) as h:
        rc, recs = h.callbacks()
        assert rc < 0 and recs == []
        with pytest.raises(RuntimeError):
            h.integrals_host()
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            capi.Handle(s)  # libECP_init -> NULL: no device, loud failure


def test_unsupported_arguments_rejected():
    s = synth.cfg2()
    with pytest.raises(RuntimeError):
        capi.Handle(s, n=3, tables_only=True)  # the reference's shift table ends at second derivatives (src/libecp.c:246-250)
    bad = synth.assemble("bad", [(0, 0, 0)], [synth.tz_basis(5)], [synth.ecp_set(3)])  # maxLBS > L+1
    with pytest.raises(RuntimeError):
        capi.Handle(bad, tables_only=True)


@pytest.mark.parametrize("name", list(SHAPES))
def test_host_tables_bitwise(name):
    s = SHAPES[name]()
    o = Oracle(s)
    # twice: the first handle of a process computes the molecule-independent tables, later ones copy them from the
    # process-wide cache (tables.c) - both must be the reference's doubles
    with capi.Handle(s, tables_only=True) as h0:
        for t in ["fac", "dfac", "poly2sph", "omega", "small_x", "small_w", "large_x", "large_w", "bessel", "besselC"]:
            assert np.array_equal(o.table(t), h0.host_table(t)), t
    with capi.Handle(s, tables_only=True) as h:
        for t in ["fac", "dfac", "poly2sph", "omega", "small_x", "small_w", "large_x", "large_w", "bessel", "besselC"]:
            assert np.array_equal(o.table(t), h.host_table(t)), t
        # slot layouts are permutations of the original grids
        oidx = h.host_itable("small_oidx")
        assert sorted(oidx[oidx >= 0]) == list(range(383)) and (oidx < 0).sum() == 1
        assert np.array_equal(h.host_table("small_rs")[oidx >= 0], h.host_table("small_x")[oidx[oidx >= 0]])
        lo = h.host_itable("large_oidx")
        assert sorted(lo[lo >= 0]) == list(range(1023))
        assert np.array_equal(h.host_table("large_ws")[lo >= 0], h.host_table("large_w")[lo[lo >= 0]])
        # transposed Bessel table
        dims = h.host_itable("dims")
        lmax, stride = int(dims[5]), int(dims[6])
        bT = h.host_table("besselT").reshape(1601, stride)
        assert np.array_equal(bT[:, :lmax + 1].T, h.host_table("bessel").reshape(lmax + 1, 1601))
    o.close()


@pytest.mark.parametrize("name", list(SHAPES))
def test_screening_and_triple_list_identical(name):
    s = SHAPES[name]()
    o = Oracle(s)
    ns = int(s["nshells"])
    with capi.Handle(s, tables_only=True) as h:
        for c in range(s["nat"]):
            if s["shellsECP"][c] == 0:
                continue
            endl = np.zeros(8, np.int32)
            st, en, sk = (np.zeros(ns, np.int32) for _ in range(3))
            o.L.oracle_screening(C.c_void_p(o.h), c, _p(endl, _pi), _p(st, _pi), _p(en, _pi), _p(sk, _pi))
            e2, st2, en2, sk2 = h.screening(c, 8)
            L = int(max(s["lECP"]))  # same L everywhere in these shapes except cfg5s: compare the common prefix
            assert np.array_equal(sk, sk2)
            live = sk == 0
            assert np.array_equal(st[live], st2[live]) and np.array_equal(en[live], en2[live])
            if name != "cfg5s":
                assert np.array_equal(endl[:L], e2[:L])
        rc, recs = o.p.callbacks(s, keep_blocks=False)
        ref = np.array([(r[0], r[1], r[2], r[4], r[5], r[6], r[8]) for r in recs[::2]], np.int32).reshape(-1, 7)
        assert np.array_equal(h.triple_list(), ref)
    o.close()


def test_shards_partition_the_triples():
    s = synth.cfg5(30)
    with capi.Handle(s, tables_only=True) as h:
        full = {tuple(r) for r in h.triple_list()}
        seen = set()
        for world in (2, 8):
            seen.clear()
            sizes = []
            for rank in range(world):
                h.set_shard(rank, world)
                part = {tuple(r) for r in h.triple_list()}
                assert not (part & seen)
                seen |= part
                sizes.append(len(part))
            assert seen == full
            assert max(sizes) < 1.35 * len(full) / world
        h.set_shard(0, 1)


def test_builder_pass_dry_run_small_and_growing_batches():
    """the batch sequence of a real pass (small first batch, half-size second, two alternating buffers that share the
    screening scratch, leftover centres carried over) yields every executed triple exactly once for any batch size
    and thread count - regression for a slot-array overrun that only growing batch sizes exposed"""
    s = synth.cfg5(40)
    old = os.environ.get("LIBECP_B200_BATCH_TRIPLES")
    old_enum = os.environ.get("LIBECP_B200_ENUM")
    try:
        # what a matrix run leaves to the host (screening + slot layout for the device enumeration): the shell pairs
        # handed to the device are the same for any batch size / thread count, and the shards partition them
        cands = set()
        os.environ.pop("LIBECP_B200_ENUM", None)
        for bt in ("300", "20000", "3000000"):
            os.environ["LIBECP_B200_BATCH_TRIPLES"] = bt
            for thr in (5, 2):
                capi.set_host_threads(thr)
                with capi.Handle(s, tables_only=True) as h:
                    ms, n, nb = h.build_only()
                    cands.add(n)
                    tot = 0
                    for r in range(3):
                        h.set_shard(r, 3)
                        tot += h.build_only()[1]
                    assert tot == n and nb >= 1
        assert len(cands) == 1
        os.environ["LIBECP_B200_ENUM"] = "host"
        counts = set()
        for bt in ("300", "20000", "3000000"):
            os.environ["LIBECP_B200_BATCH_TRIPLES"] = bt
            for thr in (5, 2):
                capi.set_host_threads(thr)
                with capi.Handle(s, tables_only=True) as h:
                    ms, n, nb = h.build_only()
                    counts.add(n)
                    assert nb >= 1 and (bt != "300" or nb > 20)
                    tot = 0
                    for r in range(3):
                        h.set_shard(r, 3)
                        tot += h.build_only()[1]
                    assert tot == n
                    h.set_shard(0, 1)
                    assert len(h.triple_list()) == n
        assert len(counts) == 1
    finally:
        capi.set_host_threads(0)
        if old is None:
            os.environ.pop("LIBECP_B200_BATCH_TRIPLES", None)
        else:
            os.environ["LIBECP_B200_BATCH_TRIPLES"] = old
        if old_enum is None:
            os.environ.pop("LIBECP_B200_ENUM", None)
        else:
            os.environ["LIBECP_B200_ENUM"] = old_enum


def test_gather_row_layout_partitions_the_matrix():
    """rows of the device-resident gather (libecp_b200_owned_rows): every AO row belongs to exactly one rank, the
    packed upper-triangle shards add up to n(n+1)/2 and follow the shell-pair ownership of the builder"""
    from libecp_b200 import gather

    s = synth.cfg5(30)
    n = int(s["dim"])
    with capi.Handle(s, tables_only=True) as h:
        for world in (1, 2, 8):
            rows, sizes = gather.shard_layout(h, world)
            allr = np.concatenate(rows)
            assert np.array_equal(np.sort(allr), np.arange(n))
            assert all(np.all(np.diff(r) > 0) for r in rows if len(r) > 1)
            assert sum(sizes) == n * (n + 1) // 2
        # a rank's triples only ever write rows it owns: row shell of every triple = a shell whose AO rows are listed
        world = 4
        rows, _ = gather.shard_layout(h, world)
        first = np.concatenate([[0], np.cumsum(s["shellsBS"])])
        for rank in range(world):
            h.set_shard(rank, world)
            tl = h.triple_list()
            shells = np.unique(first[tl[:, 0]] + tl[:, 1])
            owners = {capi.lib().libecp_b200_pair_owner(C.c_void_p(h.h), int(x), int(x), world) for x in shells}
            assert owners <= {rank}
        h.set_shard(0, 1)


@pytest.fixture(scope="module")
def hc():
    L = C.CDLL(build.build_hostcheck())
    L.hc_bessel.argtypes = [_pd, C.c_int, _pd, C.c_int, C.c_double, _pd]
    L.hc_rsh.argtypes = [C.c_int, C.c_double, C.c_double, _pd, _pd, _pd]
    L.hc_ps93_fastT.argtypes = [_pd, _pd, _pd, _pd, _pi, _pi, C.c_int, C.c_int, C.c_double, _pd, _pi]
    L.hc_sphcoord.argtypes = [_pd, _pd]
    return L


def test_device_math_bessel_rsh_bitwise(hc):
    """ecp_math.h (the code the kernels run) vs oracle on a z sweep across the three Bessel branches and on
    random / special angles"""
    s = synth.cfg4("b")
    o = Oracle(s)
    rng = np.random.default_rng(7)
    with capi.Handle(s, tables_only=True) as h:
        bT, bC = h.host_table("besselT"), h.host_table("besselC")
        stride = int(h.host_itable("dims")[6])
        zs = np.concatenate([[0.0, -1.0, 1e-9, 9.9e-8, 1e-7, 1.00001e-7, 0.005, 15.995, 15.999999, 16.0, 16.00001,
                              50.0, 1e3, 1e5], rng.uniform(0, 16, 600), 10 ** rng.uniform(-9, 4, 600)])
        nexact = ntaylor = 0
        for z in zs:
            for lmax in (0, 3, 6, 10):
                a, b = np.zeros(lmax + 1), np.zeros(lmax + 1)
                o.L.oracle_bessel(C.c_void_p(o.h), lmax, float(z), _p(a, _pd))
                hc.hc_bessel(_p(bT, _pd), stride, _p(bC, _pd), lmax, float(z), _p(b, _pd))
                if 1e-7 <= z < 16.0:
                    # Taylor branch: the device code uses fma in the derivative recurrence (ecp_math.h): the reference's
                    # double almost everywhere; a few ulp - or, for the tiny high-order values at small z where the
                    # series cancels, < 1e-24 absolute on the scale K_0 ~ 1 - elsewhere
                    assert np.all(np.abs(a - b) <= 8 * np.spacing(np.abs(a)) + 1e-24), (z, lmax)
                    nexact += int(np.all(np.abs(a - b) <= np.spacing(np.abs(a))))
                    ntaylor += 1
                else:
                    assert np.array_equal(a, b), (z, lmax)
        assert nexact >= 0.9 * ntaylor, (nexact, ntaylor)
        fac, dfac = h.host_table("fac"), h.host_table("dfac")
        angles = [(0, 0), (np.pi, 0), (np.arccos(0.0), 0.5 * np.pi), (1.0, 1.5 * np.pi), (0.3, -1.2)]
        angles += [(rng.uniform(0, np.pi), rng.uniform(-1.5, 4.7)) for _ in range(100)]
        for th, ph in angles:
            for lmax in (0, 1, 6, 10):
                a, b = np.zeros((lmax + 1) ** 2), np.zeros((lmax + 1) ** 2)
                o.L.oracle_rsh(C.c_void_p(o.h), lmax, th, ph, _p(a, _pd))
                hc.hc_rsh(lmax, th, ph, _p(fac, _pd), _p(dfac, _pd), _p(b, _pd))
                assert np.array_equal(a, b)
    o.close()


def test_device_math_ps93_windowed_bitwise(hc):
    """the slot-ordered PS93 of the fast path reproduces integrateGC_PS93 (value, stop level, failures)
    on full and windowed integrand tables"""
    s = synth.cfg2()
    o = Oracle(s)
    rng = np.random.default_rng(3)
    with capi.Handle(s, tables_only=True) as h:
        x, ws = h.host_table("small_x"), h.host_table("small_ws")
        oidx = h.host_itable("small_oidx").astype(np.int32)
        meta = h.host_itable("small_meta").astype(np.int32)

        def perm(v):
            out = np.zeros(384)
            m = oidx >= 0
            out[m] = v[oidx[m]]
            return out

        nfail = 0
        for trial in range(200):
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
            f = Fa * Fb * U
            r1, n1 = np.zeros(1), np.zeros(1, np.int32)
            rc1 = o.L.oracle_ps93_table(C.c_void_p(o.h), _p(f, _pd), st, en, _p(r1, _pd), _p(n1, _pi))
            r2, n2 = np.zeros(1), np.zeros(1, np.int32)
            rc2 = hc.hc_ps93_fastT(_p(perm(Fa), _pd), _p(perm(Fb), _pd), _p(perm(U), _pd), _p(ws, _pd), _p(oidx, _pi),
                                   _p(meta, _pi), st, en, 1e-12, _p(r2, _pd), _p(n2, _pi))
            nfail += rc1
            assert rc1 == rc2 and n1[0] == n2[0]
            if rc1 == 0:
                assert r1[0] == r2[0]
        assert 0 < nfail < 200
    o.close()


def _small_layout(h):
    oidx = h.host_itable("small_oidx").astype(np.int32)
    meta = h.host_itable("small_meta").astype(np.int32)
    return oidx, meta


def test_type1_small_grid_levelwise_points_are_exactly_the_window(hc):
    """k_type1S visits, per level, the points t1_level lists instead of all slots: they must be exactly the slots of the
    level whose original index lies in the tabulated window [gs, ge) (reference src/type1.c:121), and cnt the number of
    points PS93 counts (left idx >= gs, right idx <= ge; src/gc_integrators.c:190-197)"""
    hc.hc_t1_level_slots.argtypes = [_pi, _pi, C.c_int, C.c_int, C.c_int, _pi, C.POINTER(C.c_int)]
    with capi.Handle(synth.cfg2(), tables_only=True) as h:
        oidx, meta = _small_layout(h)
    lev_slot = meta[39:53]
    rng = np.random.default_rng(5)
    windows = [(0, 383), (0, 1), (382, 383), (100, 101), (0, 200), (191, 192)]
    windows += [tuple(sorted(rng.integers(0, 384, 2))) for _ in range(300)]
    out = np.zeros(400, np.int32)
    for gs, ge in windows:
        if not gs < ge:
            continue
        for v in range(4, 13):
            cnt = C.c_int(0)
            n = hc.hc_t1_level_slots(_p(oidx, _pi), _p(meta, _pi), v, int(gs), int(ge), _p(out, _pi), C.byref(cnt))
            s0, s1 = int(lev_slot[v]), int(lev_slot[v + 1])
            sl = np.arange(s0, s1)
            oi = oidx[sl]
            want = sl[(oi >= gs) & (oi < ge)]
            assert n >= 0 and sorted(out[:n].tolist()) == want.tolist(), (gs, ge, v)
            left, right = sl[0::2], sl[1::2]
            assert cnt.value == int((oidx[left] >= gs).sum() + (oidx[right] <= ge).sum()), (gs, ge, v)


def test_large_grid_levelwise_candidates_cover_every_live_point(hc):
    """k_type1L / the gate bracket: the candidates of a level (lg_live_range + lg_level) contain every slot whose
    exponent passes the gate e(r) >= ln(acc) (reference src/type2.c:479-490, src/type1.c:163), and far fewer slots than
    the level has"""
    hc.hc_lg_level_slots.argtypes = [_pd, C.c_int, C.c_int] + [C.c_double] * 5 + [C.c_int, _pi]
    hc.hc_fm06_map.argtypes = [C.c_double, C.c_double, _pd, _pd]
    with capi.Handle(synth.cfg2(), tables_only=True) as h:
        xo = h.host_table("large_x")
        xs = h.host_table("large_xs")
    order, slots = len(xo), len(xs)
    ln_acc = np.log(1e-14) - 2.0
    rng = np.random.default_rng(11)
    out = np.zeros(slots, np.int32)
    tot = cand = live_n = 0
    for it in range(400):
        zA = 60.0 * 2.6 ** (-float(rng.integers(0, 9)))
        zB = 24.0 * 2.6 ** (-float(rng.integers(0, 9)))
        dAC, dBC = float(rng.choice([0.0, 1e-3, 4.5, 5.45, 9.44, 10.9, 31.0])), float(rng.choice([0.0, 5.45, 9.44, 22.0]))
        zp = zA + zB
        i1, i2 = np.zeros(1), np.zeros(1)
        hc.hc_fm06_map(zp, (zA * dAC + zB * dBC) / zp, _p(i1, _pd), _p(i2, _pd))
        r = i1[0] * xs + i2[0]
        e = -zA * (dAC - r) ** 2 - zB * (dBC - r) ** 2
        live = e >= ln_acc
        a, b, c0 = -zp, 2.0 * (zA * dAC + zB * dBC), -(zA * dAC * dAC + zB * dBC * dBC)
        for lev in range(3, 10):
            n = hc.hc_lg_level_slots(_p(xo, _pd), order, slots, a, b, c0 - ln_acc, float(i1[0]), float(i2[0]), lev, _p(out, _pi))
            got = set(out[:n].tolist())
            lo, hi = 1 << lev, 2 << lev
            assert all(lo <= s < hi for s in got) and len(got) == n
            need = set((np.nonzero(live[lo:hi])[0] + lo).tolist())
            assert need <= got, (zA, zB, dAC, dBC, lev)
            tot += hi - lo
            cand += n
            live_n += len(need)
    assert cand < 0.6 * tot and cand <= live_n + 4 * 7 * 400  # at most a point or two of slack per run and level


_ECP_SHIPPED_FORMAT = """6 2 4
  0 1
          37.4565000               6.8446000  2
  1 1
          -2.6739000               7.9317000  2
  2 1
          -0.5945000               6.0209000  2
  3 1
           0.0000000               1.0000000  2
"""


def _write_bs(path, s):
    """a basis-set file in the layout of the reference's example input, written from the arrays"""
    with open(path, "w") as f:
        sh = p = 0
        for a in range(int(s["nat"])):
            f.write(f"6 {int(s['shellsBS'][a])}\n")
            for _ in range(int(s["shellsBS"][a])):
                f.write(f" {int(s['lBS'][sh])} {int(s['KBS'][sh])}\n")
                for k in range(int(s["KBS"][sh])):
                    f.write(f"  {k + 1}  {float(s['aBS'][p])!r}   {float(s['dBS'][p])!r}\n")
                    p += 1
                sh += 1


def test_loaders_reproduce_config1_and_the_corrected_reading(tmp_path):
    """include/libecp_b200_io.h: the text loaders give configuration 1 ("as shipped": the reference's loader applied to
    the shipped ECP format) array for array; the corrected reading keeps signs and columns; broken
    files are reported, not read past"""
    from libecp_b200 import io as ecpio

    ref = synth.cfg1()
    (tmp_path / "c.xyz").write_text("1\ncarbon atom\nC 0.0 0.0 0.0\n")
    (tmp_path / "c.ecp").write_text(_ECP_SHIPPED_FORMAT * 2)  # the shipped file holds two carbon blocks
    _write_bs(tmp_path / "c.bs", ref)
    got = ecpio.load(tmp_path / "c.xyz", tmp_path / "c.ecp", tmp_path / "c.bs")
    for k in ("nat", "dim", "nshells"):
        assert got[k] == ref[k], k
    for k in ("geometry", "shellsECP", "lECP", "KECP", "nECP", "dECP", "aECP", "shellsBS", "lBS", "KBS", "dBS", "aBS"):
        assert np.array_equal(np.asarray(got[k]), np.asarray(ref[k])), k
    fixed = ecpio.load(tmp_path / "c.xyz", tmp_path / "c.ecp", tmp_path / "c.bs", ecp_format=ecpio.SHIPPED)
    assert np.array_equal(fixed["dECP"], [37.4565, -2.6739, -0.5945, 0.0])
    assert np.array_equal(fixed["aECP"], [6.8446, 7.9317, 6.0209, 1.0])
    assert np.array_equal(fixed["nECP"], [2.0, 2.0, 2.0, 2.0]) and np.array_equal(fixed["lECP"], [0, 1, 2, 3])
    # two atoms: second block of the same files
    (tmp_path / "c2.xyz").write_text("2\n\nC 0 0 0\nC 0.0 0.0 2.5e0\n")
    two = dict(ref)
    two.update(nat=2, shellsBS=np.concatenate([ref["shellsBS"]] * 2), lBS=np.concatenate([ref["lBS"]] * 2),
               KBS=np.concatenate([ref["KBS"]] * 2), aBS=np.concatenate([ref["aBS"]] * 2), dBS=np.concatenate([ref["dBS"]] * 2))
    _write_bs(tmp_path / "c2.bs", two)
    g2 = ecpio.load(tmp_path / "c2.xyz", tmp_path / "c.ecp", tmp_path / "c2.bs")
    assert g2["nat"] == 2 and g2["dim"] == 10 and np.array_equal(g2["geometry"], [0, 0, 0, 0, 0, 2.5])
    assert np.array_equal(g2["aECP"], np.concatenate([ref["aECP"]] * 2)) and np.array_equal(g2["shellsECP"], [4, 4])
    # errors: missing file, truncated file, three atoms but two blocks
    with pytest.raises(OSError):
        ecpio.load(tmp_path / "nope.xyz", tmp_path / "c.ecp", tmp_path / "c.bs")
    (tmp_path / "short.ecp").write_text(_ECP_SHIPPED_FORMAT[:60])
    with pytest.raises(OSError):
        ecpio.load(tmp_path / "c.xyz", tmp_path / "short.ecp", tmp_path / "c.bs")
    (tmp_path / "c3.xyz").write_text("3\n\nC 0 0 0\nC 0 0 2\nC 0 0 4\n")
    with pytest.raises(OSError):
        ecpio.load(tmp_path / "c3.xyz", tmp_path / "c.ecp", tmp_path / "c2.bs")


def test_loaders_on_the_shipped_example_files():
    """the reference's own example inputs (read only where the reference tree is present: this container)"""
    from libecp_b200 import io as ecpio

    d = "/root/reference/example"
    if not os.path.exists(os.path.join(d, "test_c.ecp")):
        pytest.skip("reference tree not present")
    import tempfile

    with tempfile.TemporaryDirectory() as tmp:
        xyz = os.path.join(tmp, "c.xyz")
        open(xyz, "w").write("1\nC\nC 0 0 0\n")
        got = ecpio.load(xyz, os.path.join(d, "test_c.ecp"), os.path.join(d, "test_c.bs"))
    ref = synth.cfg1()
    for k in ("shellsECP", "lECP", "KECP", "nECP", "dECP", "aECP", "shellsBS", "lBS", "KBS", "dBS", "aBS"):
        assert np.array_equal(np.asarray(got[k]), np.asarray(ref[k])), k


def test_example_program_compiles_against_the_public_headers(tmp_path):
    """examples/ex1.c (the reference's example program on this library) builds with nothing but include/ and the .so"""
    import subprocess

    exe = tmp_path / "ex1"
    cmd = ["gcc", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "ex1.c"),
           "-L", os.path.dirname(capi.SO_PATH), "-lecp_b200", "-Wl,-rpath," + os.path.dirname(capi.SO_PATH), "-o", str(exe)]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert p.returncode == 0, p.stdout
    p = subprocess.run([str(exe)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert p.returncode == 2 and "usage" in p.stdout


def test_screening_windows_identical_over_many_distances():
    """the lookup-bracketed window searches of the builder give the reference's integers (src/type2.c:148-180) for
    shells at distances from 1e-3 to 60 bohr of centres with different potential cut-offs (two ECP kinds), including
    shells that are skipped and windows clipped by the potential"""
    rng = np.random.default_rng(3)
    dists = np.concatenate([[1e-3, 0.05, 0.0625, 0.125, 1.0, 36.9, 37.3, 40.0, 60.0], rng.uniform(0.0, 45.0, 60)])
    coords = [(0.0, 0.0, 0.0)]
    for k, d in enumerate(dists):
        v = rng.normal(size=3)
        v *= d / np.linalg.norm(v)
        coords.append(tuple(v))
    n = len(coords)
    ecps = [synth.ecp_set(4)] + [synth.ecp_set(2, 1.5) if k % 7 == 0 else None for k in range(1, n)]
    s = synth.assemble("window_stress", coords, [synth.tz_basis(3)] * n, ecps)
    o = Oracle(s)
    ns = int(s["nshells"])
    checked = skipped = 0
    with capi.Handle(s, tables_only=True) as h:
        for c in range(n):
            if s["shellsECP"][c] == 0:
                continue
            endl = np.zeros(8, np.int32)
            st, en, sk = (np.zeros(ns, np.int32) for _ in range(3))
            o.L.oracle_screening(C.c_void_p(o.h), c, _p(endl, _pi), _p(st, _pi), _p(en, _pi), _p(sk, _pi))
            _, st2, en2, sk2 = h.screening(c, 8)
            assert np.array_equal(sk, sk2), c
            live = sk == 0
            assert np.array_equal(st[live], st2[live]) and np.array_equal(en[live], en2[live]), c
            checked += int(live.sum())
            skipped += int((~live).sum())
    o.close()
    assert checked > 500 and skipped > 500


def test_sharding_queries_on_a_handle_without_ecp_centres():
    """owned_rows / pair_owner need only the basis bookkeeping: legal on a handle with no ECP centre (every rank of a
    gather lists every rank's rows)"""
    s = synth.mask_centres(synth.cfg3(4), [])
    h = capi.Handle(s, tables_only=True)
    rows = [h.owned_rows(r, 2) for r in range(2)]
    assert sorted(np.concatenate(rows).tolist()) == list(range(int(s["dim"])))
    assert capi.lib().libecp_b200_pair_owner(h.h, 0, 5, 2) in (0, 1)
    h.close()


def test_init_rejections_carry_their_own_message():
    s = synth.cfg2()
    with pytest.raises(RuntimeError, match="derivative order n=3"):
        capi.Handle(s, n=3, tables_only=True)
    # g shells under an L = 2 potential: outside the compiled-in shape domain, rejected with the shape in the message
    bad = synth.assemble("bad", [(0.0, 0.0, 0.0)], [synth.tz_basis(4)], [synth.ecp_set(2)])
    with pytest.raises(RuntimeError, match="unsupported shape: max l of the basis 4, max L of the ECPs 2"):
        capi.Handle(bad, tables_only=True)


def test_digest_checker_accepts_the_reference_and_rejects_a_perturbation(tmp_path):
    """libecp_b200/parity.py (the result check of bench.py and of the full config-5 GPU test) on a small system:
    the reference's own matrix passes, a 1e-6 relative change of one element or a value below the diagonal does not"""
    from libecp_b200 import parity
    from oracle.refbind import RefLib, have

    s = synth.cfg3(3)
    M = RefLib("ref" if have("ref") else "port").get_integrals(s)
    f = str(tmp_path / "d.npz")
    parity.make_digest(M, s, f, nsample=4000)
    assert parity.check_digest(M, s, f)["ok"]
    i, j = np.unravel_index(np.argmax(np.abs(M)), M.shape)
    bad = M.copy()
    bad[i, j] *= 1.0 + 1e-6
    assert not parity.check_digest(bad, s, f)["ok"]
    bad = M.copy()
    bad[5, 2] = 1e-9
    assert not parity.check_digest(bad, s, f)["ok"]
    bad = M.copy()
    off = parity.atom_ao_offsets(s)
    bad[off[0]:off[1], off[2]:off[3]] = 0.0  # a whole (atom, atom) block missing: screening / indexing error
    r = parity.check_digest(bad, s, f)
    assert not r["ok"] and not r["block_support_equal"]


def test_committed_config5_digest_is_self_consistent():
    z = np.load(os.path.join(GOLDEN, "cfg5_full_digest.npz"))
    assert int(z["dim"]) == 19000 and int(z["nat"]) == 500 and int(z["nominal"]) == 9767187500
    assert abs(z["rowsum"].sum() - float(z["sum"])) <= 1e-12 * float(z["sumabs"])
    assert abs(z["ablk_sum"].sum() - float(z["sum"])) <= 1e-12 * float(z["sumabs"])
    assert abs(z["ablk_abs"].sum() - float(z["sumabs"])) <= 1e-12 * float(z["sumabs"])
    assert np.all(np.tril(z["ablk_abs"], -1) == 0.0)
    # the four single-centre digests of round 1 are partial sums of it: every element they sample that is also sampled
    # here cannot exceed it in a way additivity forbids - checked loosely through the support of the rows
    for c in (0, 1, 288, 289):
        zc = np.load(os.path.join(GOLDEN, f"cfg5_c{c}_digest.npz"))
        assert np.all((zc["rowabs"] > 0) <= (z["rowabs"] > 0))


def test_custom_shell_ordering_tables():
    """scope row f4 (reference src/libecp.c:152-166): with a caller-supplied component order the order-dependent
    host tables are the default ones with their monomial index permuted; a too short ordering is rejected"""
    s = synth.cfg2()
    with capi.Handle(s, tables_only=True) as h0:
        tm = int(h0.host_itable("dims")[4])
        p0, ijk0 = h0.host_table("poly2sph"), h0.host_itable("ijk").reshape(-1, 3)
    lmax = tm + 1
    for kind in ("reversed", "zfirst"):
        order = synth.shell_order(lmax, kind)
        with capi.Handle(s, tables_only=True, ordering=order, lmax=lmax) as h:
            p1, ijk1 = h.host_table("poly2sph"), h.host_itable("ijk").reshape(-1, 3)
            idx = h.host_itable("ijkIndex").reshape(lmax + 1, lmax + 1, lmax + 1)
        assert np.array_equal(ijk1, order.reshape(-1, 3)[:len(ijk1)])
        cols = len(p0) // len(ijk0)
        where = {tuple(e): i for i, e in enumerate(ijk0)}
        perm = np.array([where[tuple(e)] for e in ijk1])
        # same numbers up to the order of the sums over a shell's components (src/transformations.c:146-207)
        assert np.allclose(p1.reshape(-1, cols), p0.reshape(-1, cols)[perm], rtol=1e-13, atol=1e-15)
        for i, e in enumerate(ijk1):
            assert idx[tuple(e)] == i
    with pytest.raises(RuntimeError):
        capi.Handle(s, tables_only=True, ordering=synth.shell_order(tm, "reversed"), lmax=tm)
    with pytest.raises(RuntimeError):  # not a permutation of the monomials of a shell
        bad = synth.shell_order(lmax, "libint").copy()
        bad[3:6] = bad[6:9]
        capi.Handle(s, tables_only=True, ordering=bad, lmax=lmax)


DERIV = {"deriv1_tz2_L4": (2, 4), "deriv1_tz3_L5": (3, 5)}
DERIV2 = {"deriv2_tz1_L4": lambda: synth.deriv_pair(1, 4), "deriv2_tz2_L5": lambda: synth.deriv_pair(2, 5),
          "deriv2_triangle": synth.deriv2_triangle}


@pytest.mark.parametrize("name", list(DERIV2))
def test_second_derivative_callback_keys_match_reference(name):
    """scope row f1, n = 2, host part: ten shifts per executed shell pair (src/libecp.c:246-250), no translational-invariance
    skip (:325-330 only for n = 1), call order and keys of the compiled reference; tables sized from maxLBS + 2"""
    d = np.load(os.path.join(GOLDEN, f"{name}_blocks.npz"))
    s = DERIV2[name]()
    with capi.Handle(s, n=2, tables_only=True) as h:
        keys = h.callback_keys()
        dims = h.host_itable("dims")
    assert np.array_equal(keys, d["keys"][0::2]) and np.array_equal(keys, d["keys"][1::2])
    assert dims[1] == max(s["lBS"]) + 2
    # block sizes of the fixture: shifted momenta, except the (+1,0) / (0,+1) blocks (momentum unchanged)
    ijk = lambda l: (l + 1) * (l + 2) // 2
    for k in (0, len(keys) // 2, len(keys) - 1):
        A, s1, la, sa, B, s2, lb, sb, C = d["keys"][2 * k]
        ea, eb = (0, 0) if (sa, sb) in ((1, 0), (0, 1)) else (sa, sb)
        assert d["off"][2 * k + 1] - d["off"][2 * k] == ijk(la + ea) * ijk(lb + eb)


def test_derivative_callback_keys_triangle():
    """three atoms, mixed shapes, one atom without ECP: the key sequence of the n = 1 run is the reference's"""
    d = np.load(os.path.join(GOLDEN, "deriv1_triangle_blocks.npz"))
    with capi.Handle(synth.deriv_triangle(), n=1, tables_only=True) as h:
        keys = h.callback_keys()
    assert np.array_equal(keys, d["keys"][0::2]) and np.array_equal(keys, d["keys"][1::2])


@pytest.mark.parametrize("name", list(DERIV))
def test_derivative_callback_keys_match_reference(name):
    """scope row f1, host part: the (shifted) triples of a first-derivative run and their call order
    (A,s1,la,shifta,B,s2,lb,shiftb,C) are the reference's (src/libecp.c:246-250,297-330); the tables have the sizes the
    reference derives from maxLBS + n (src/libecp.c:143-147)"""
    d = np.load(os.path.join(GOLDEN, f"{name}_blocks.npz"))
    lbs, L = DERIV[name]
    with capi.Handle(synth.deriv_pair(lbs, L), n=1, tables_only=True) as h:
        keys = h.callback_keys()
        dims = h.host_itable("dims")
    assert np.array_equal(keys, d["keys"][0::2]) and np.array_equal(keys, d["keys"][1::2])
    assert dims[1] == lbs + 1 and dims[2] == lbs + 1 and dims[3] == L - 1 + lbs + 1


def test_rows_streamed_out_are_final():
    """streamed download of the host consumer (api.c: stream_after_batch): the AO rows of an atom leave the device once the
    pass is beyond lastCentre[atom] (builder.c: ecp_atom_last_centre).  They must be final: no later centre keeps a shell
    of the atom after screening (src/type2.c:148-180), and an atom with lastCentre = -1 is never touched at all"""
    far = synth.assemble("far", [(0.0, 0.0, 0.0), (0.0, 0.0, 40.0), (3.0, 0.5, 0.2), (60.0, 60.0, 0.0)], [synth.tz_basis(2)] * 4,
                         [synth.ecp_set(3), synth.ecp_set(3), None, None])
    for s in (synth.cfg5(40), synth.cfg3(4), far):
        with capi.Handle(s, tables_only=True) as h:
            last = h.host_itable("lastCentre")
            types = h.host_itable("atomType")
            first = np.concatenate([[0], np.cumsum(s["shellsBS"])])
            nat = int(s["nat"])
            assert len(last) == nat
            reached = np.full(nat, -1)
            for c in range(nat):
                if types[c] < 0:
                    continue
                _, st, en, sk = h.screening(c, 8)  # room for end_l of any L <= 6
                for x in range(nat):
                    if np.any(sk[first[x]:first[x + 1]] == 0):
                        reached[x] = c
            assert np.all(last >= reached), (s.get("name"), last, reached)  # never final too early
            assert np.all((last >= 0) | (reached < 0))
        if s is far:
            assert last[3] == -1 and reached[3] == -1
