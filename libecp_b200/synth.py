"""Closed-form synthetic basis / ECP / geometry generators for the five BASELINE.json configs.

No RNG: every language regenerates these inputs bit-identically.  The arrays are the
flat argument arrays of ``libECP_init`` / ``getIntegrals`` (reference: src/libecp.c:53-61,
src/getIntegrals.h:7-13).  Units: bohr.

A *system* is a dict of numpy arrays:
  geometry[3*nat] f64, shellsECP[nat] i32, lECP[], KECP[] i32, nECP[], dECP[], aECP[] f64,
  shellsBS[nat] i32, lBS[], KBS[] i32, dBS[], aBS[] f64, plus ``dim`` (number of Cartesian AOs),
  ``nshells`` and ``name``.
"""
from __future__ import annotations

import math

import numpy as np

LARGE_GRID_ORDER = 1024
TOLERANCE = 1.0e-12
ACCURACY = 1.0e-14

_K_PATTERN = {0: [2, 1, 1, 1, 1, 1], 1: [4, 1, 1], 2: [3, 1, 1], 3: [1], 4: [1, 1], 5: [1]}
_ZETA0 = [60.0, 24.0, 9.0, 1.6, 1.2, 1.0]


def ijk_dim(l: int) -> int:
    return (l + 1) * (l + 2) // 2


def tz_basis(lmaxbs: int):
    """``TZ(lmaxbs)``: def2-TZVP-Au-shaped segmented pattern (SURVEY.md §8d "Basis generator")."""
    ls, ks, ds, as_ = [], [], [], []
    for l in range(lmaxbs + 1):
        k = 0  # running primitive index inside the l-block
        for K in _K_PATTERN[l]:
            ls.append(l)
            ks.append(K)
            for _ in range(K):
                zeta = _ZETA0[l] * math.pow(2.6, -k)
                if K == 1:
                    d = 1.0
                else:
                    d = 0.25 + 0.5 * math.fabs(math.cos(1.3 * k + 0.7 * l))
                    if l == 0 and k % 2 == 1:
                        d = -d
                as_.append(zeta)
                ds.append(d)
                k += 1
    return ls, ks, ds, as_


def ecp_set(L: int, ascale: float = 1.0):
    """``ECP(L)``: two Gaussians per channel l=0..L, n=2; channel l=L is the local part."""
    ls, ks, ns, ds, as_ = [], [], [], [], []
    for l in range(L + 1):
        ls.append(l)
        ks.append(2)
        for k in range(2):
            if l < L:
                a = 13.0 / (1 + l) / (1 + 1.5 * k)
                d = 420.0 / (1 + 2 * l) * (0.1 if k else 1.0)
            else:
                a = 4.0 / (1 + 1.5 * k)
                d = -30.0 * (0.1 if k else 1.0)
            ns.append(2.0)
            ds.append(d)
            as_.append(a * ascale)
    return ls, ks, ns, ds, as_


def assemble(name, coords, bases, ecps):
    """coords: list of (x,y,z); bases[i] = tz_basis(...) tuple; ecps[i] = ecp_set(...) tuple or None."""
    nat = len(coords)
    g = np.array(coords, dtype=np.float64).reshape(-1)
    sE, lE, kE, nE, dE, aE = [], [], [], [], [], []
    sB, lB, kB, dB, aB = [], [], [], [], []
    for i in range(nat):
        if ecps[i] is None:
            sE.append(0)
        else:
            ls, ks, ns, ds, as_ = ecps[i]
            sE.append(len(ls))
            lE += ls; kE += ks; nE += ns; dE += ds; aE += as_
        ls, ks, ds, as_ = bases[i]
        sB.append(len(ls))
        lB += ls; kB += ks; dB += ds; aB += as_
    dim = sum(ijk_dim(l) for l in lB)
    i32 = lambda v: np.array(v, dtype=np.int32)
    f64 = lambda v: np.array(v, dtype=np.float64)
    return dict(name=name, nat=nat, geometry=g, shellsECP=i32(sE), lECP=i32(lE), KECP=i32(kE),
                nECP=f64(nE), dECP=f64(dE), aECP=f64(aE), shellsBS=i32(sB), lBS=i32(lB), KBS=i32(kB),
                dBS=f64(dB), aBS=f64(aB), dim=dim, nshells=len(lB))


def cfg1():
    """example/ex1.c on test_c.bs + test_c.ecp as shipped, single carbon at the origin.

    These are exactly the arrays ex1's loaders produce (including the ECP parser quirk,
    SURVEY.md App. C-1 / App. D; reference example/ex1.c:75-123).
    """
    return dict(
        name="cfg1_ex1_carbon", nat=1, geometry=np.zeros(3),
        shellsECP=np.array([4], np.int32), lECP=np.array([0, 1, 2, 3], np.int32),
        KECP=np.array([1, 1, 1, 1], np.int32),
        aECP=np.array([0.4565, 0.6739, 0.5945, 0.0]), dECP=np.array([6.8446, 7.9317, 6.0209, 1.0]),
        nECP=np.array([2.0, 2.0, 2.0, 2.0]),
        shellsBS=np.array([3], np.int32), lBS=np.array([0, 0, 1], np.int32), KBS=np.array([3, 3, 3], np.int32),
        aBS=np.array([153.17226, 23.07303, 4.92329, 6.616612, 0.525856, 0.169958, 4.91292, 0.997616, 0.232685]),
        dBS=np.array([0.07074, 0.39538, 0.663311, -0.08138, 0.574853, 0.502413, 0.109931, 0.462713, 0.627514]),
        dim=5, nshells=3)


def cfg2(L: int = 4):
    """single heavy atom, TZ(3) + ECP(L) (L=4: semi-local s,p,d,f + local g)."""
    return assemble(f"cfg2_heavy_atom_L{L}", [(0.0, 0.0, 0.0)], [tz_basis(3)], [ecp_set(L)])


def au20_coords(natoms: int = 20):
    d = 5.45
    e1 = (1.0, 0.0, 0.0)
    e2 = (0.5, math.sqrt(3.0) / 2.0, 0.0)
    e3 = (0.5, math.sqrt(3.0) / 6.0, math.sqrt(2.0 / 3.0))
    out = []
    for i in range(4):
        for j in range(4):
            for k in range(4):
                if i + j + k <= 3:
                    out.append(tuple(d * (i * e1[c] + j * e2[c] + k * e3[c]) for c in range(3)))
    return out[:natoms]


def cfg3(natoms: int = 20):
    """Au20 tetrahedron (or its first ``natoms`` atoms), TZ(3) + ECP(4) on every centre."""
    c = au20_coords(natoms)
    return assemble(f"cfg3_au{natoms}", c, [tz_basis(3)] * len(c), [ecp_set(4)] * len(c))


def deriv_pair(lbs: int = 2, L: int = 4):
    """two atoms off-axis, TZ(lbs) + ECP(L): shape of the derivative fixtures (tests/golden/deriv1_*; needs lbs + n <= L - 1)"""
    return assemble(f"deriv_tz{lbs}_L{L}", [(0.0, 0.0, 0.0), (0.3, -0.4, 4.1)], [tz_basis(lbs)] * 2, [ecp_set(L)] * 2)


def shell_order(lmax: int, kind: str = "libint"):
    """Cartesian component order of the shells l = 0..lmax in the layout of cartesianShellOrder(lmax)
    (reference src/dimensions.c:17-37): int32 [3 * C_DIM(lmax)], exponent triples (nx, ny, nz).
    kind "libint": the library's default; "reversed": every shell back to front (z^l first);
    "zfirst": components sorted by (nz, ny) descending - neither order is a relabelling of the axes of the default."""
    out = []
    for l in range(lmax + 1):
        comp = [(l - i, i - j, j) for i in range(l + 1) for j in range(i + 1)]
        if kind == "reversed":
            comp = comp[::-1]
        elif kind == "zfirst":
            comp = sorted(comp, key=lambda e: (-e[2], -e[1]))
        elif kind != "libint":
            raise ValueError(kind)
        out += [x for e in comp for x in e]
    return np.array(out, np.int32)


def deriv_triangle():
    """three atoms in general position with three different shapes - TZ(2) + ECP(4), TZ(1) + ECP(5, scaled exponents) and
    TZ(2) without an ECP: first-derivative triples with A != B != C, mixed L and an atom that is no ECP centre
    (needs lbs + n <= L - 1 on every centre)"""
    return assemble("deriv_triangle", [(0.0, 0.0, 0.0), (2.9, 0.4, -0.6), (-0.7, 3.1, 1.2)],
                    [tz_basis(2), tz_basis(1), tz_basis(2)], [ecp_set(4), ecp_set(5, 0.8), None])


def deriv2_triangle():
    """second-derivative companion of deriv_triangle: TZ(1) + ECP(4), TZ(1) + ECP(5, scaled exponents), TZ(0) without an ECP
    (lbs + 2 <= L - 1 on every centre)"""
    return assemble("deriv2_triangle", [(0.0, 0.0, 0.0), (2.9, 0.4, -0.6), (-0.7, 3.1, 1.2)],
                    [tz_basis(1), tz_basis(1), tz_basis(0)], [ecp_set(4), ecp_set(5, 0.8), None])


def random_system(seed: int):
    """small random molecule for randomized parity runs: 2-4 atoms at random positions (>= 1.5 bohr apart), every atom a
    TZ(0..3) basis, ECP(L) with L >= lbs + 1 on a random non-empty subset of the atoms, random exponent scale"""
    rng = np.random.default_rng(seed)
    nat = int(rng.integers(2, 5))
    pts = []
    while len(pts) < nat:
        p = rng.uniform(-3.5, 3.5, 3)
        if all(np.linalg.norm(p - q) >= 1.5 for q in pts):
            pts.append(p)
    lbs = [int(rng.integers(0, 4)) for _ in range(nat)]
    lmax = max(lbs)
    has = rng.random(nat) < 0.6
    has[int(rng.integers(0, nat))] = True
    ecps = [ecp_set(int(rng.integers(lmax + 1, 6)), float(rng.uniform(0.6, 1.6))) if has[i] else None for i in range(nat)]
    return assemble(f"random_{seed}", [tuple(p) for p in pts], [tz_basis(l) for l in lbs], ecps)


def cfg4(variant: str = "a"):
    """high-angular-momentum stress: (a) TZ(4)+ECP(5), (b) TZ(5)+ECP(6); 2 atoms on the z axis."""
    lbs, L = (4, 5) if variant == "a" else (5, 6)
    c = [(0.0, 0.0, 0.0), (0.0, 0.0, 4.5)]
    return assemble(f"cfg4{variant}", c, [tz_basis(lbs)] * 2, [ecp_set(L)] * 2)


def pbs_sites(natoms: int = 500):
    sites = []
    for x in range(8):
        for y in range(8):
            for z in range(8):
                idx = 64 * x + 8 * y + z
                d2 = (x - 3.5) ** 2 + (y - 3.5) ** 2 + (z - 3.5) ** 2
                sites.append((d2, idx, x, y, z))
    sites.sort(key=lambda t: (t[0], t[1]))
    keep = sorted(sites[:natoms], key=lambda t: t[1])
    return keep


def cfg5(natoms: int = 500, active=None):
    """500-heavy-atom PbS-like rock-salt nanocrystal.

    ``active``: optional iterable of atom indices that keep their ECP (others get shellsECP=0 and
    the flat ECP arrays are compacted, the masking the reference API supports, src/libecp.c:95-127,257).
    """
    keep = pbs_sites(natoms)
    a0 = 5.609
    coords, bases, ecps = [], [], []
    pb_b, s_b = tz_basis(3), tz_basis(2)
    pb_e, s_e = ecp_set(4), ecp_set(2, ascale=1.5)
    act = None if active is None else set(active)
    for n, (_, idx, x, y, z) in enumerate(keep):
        coords.append((a0 * x, a0 * y, a0 * z))
        pb = (x + y + z) % 2 == 0
        bases.append(pb_b if pb else s_b)
        e = pb_e if pb else s_e
        ecps.append(e if (act is None or n in act) else None)
    tag = "" if active is None else "_c" + "_".join(str(a) for a in sorted(act))
    return assemble(f"cfg5_pbs{natoms}{tag}", coords, bases, ecps)


def probe(nat: int, lmaxbs: int, LE: int, spacing: float = 5.4):
    """Session probe shape ``S(nat,lmaxbs,LE)`` of SURVEY.md App. D (secondary checks)."""
    side = int(math.ceil(nat ** (1.0 / 3.0) - 1e-9))
    coords = []
    for i in range(nat):
        coords.append((spacing * (i % side) + 0.013 * i, spacing * ((i // side) % side) - 0.007 * i,
                       spacing * (i // (side * side)) + 0.003 * i))
    ls, ks, ds, as_ = [], [], [], []
    for l in range(lmaxbs + 1):
        ls += [l, l, l]
        ks += [3, 1, 1]
        for k in range(3):
            as_.append(30.0 / (1 + l) / 3.2 ** k)
            ds.append(0.2 + 0.3 * k)
        as_.append(0.9 / (1 + 0.5 * l)); ds.append(1.0)
        as_.append(0.12 / (1 + 0.3 * l)); ds.append(1.0)
    b = (ls, ks, ds, as_)
    return assemble(f"probe_S{nat}_{lmaxbs}_{LE}", coords, [b] * nat, [ecp_set(LE)] * nat)


def mask_centres(sys_, active):
    """Keep the ECP only on atoms in ``active`` (compacting the flat ECP arrays)."""
    act = set(active)
    sE = sys_["shellsECP"].copy()
    lE, kE, nE, dE, aE = [], [], [], [], []
    si = pi = 0
    for i in range(sys_["nat"]):
        ns = int(sys_["shellsECP"][i])
        for s in range(ns):
            K = int(sys_["KECP"][si])
            if i in act:
                lE.append(int(sys_["lECP"][si])); kE.append(K)
                nE += list(sys_["nECP"][pi:pi + K]); dE += list(sys_["dECP"][pi:pi + K]); aE += list(sys_["aECP"][pi:pi + K])
            si += 1; pi += K
        if i not in act:
            sE[i] = 0
    out = dict(sys_)
    out.update(shellsECP=sE, lECP=np.array(lE, np.int32), KECP=np.array(kE, np.int32),
               nECP=np.array(nE, np.float64), dECP=np.array(dE, np.float64), aECP=np.array(aE, np.float64))
    out["name"] = sys_["name"] + "_mask" + "_".join(str(a) for a in sorted(act))
    return out


def nominal_triples(sys_) -> int:
    ncent = int((sys_["shellsECP"] > 0).sum())
    ns = int(sys_["nshells"])
    return ncent * ns * (ns + 1) // 2


CONFIGS = {
    "cfg1": cfg1, "cfg2": cfg2, "cfg3": cfg3, "cfg4a": lambda: cfg4("a"), "cfg4b": lambda: cfg4("b"), "cfg5": cfg5,
}
