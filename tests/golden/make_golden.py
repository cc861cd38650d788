"""Generate the golden fixtures from the UNMODIFIED reference (oracle/_ref/libecp_ref.so, built by
oracle/Makefile from /root/reference/src).  Run in the build container only (the GPU box has no
/root/reference); the resulting .npz files are committed.

    python tests/golden/make_golden.py [--big | --deriv | --deriv2 | --order]

Fixtures (float64, exact bytes of the reference's output):
  <cfg>_matrix.npz   : getIntegrals matrix (upper triangle incl. diagonal as a flat vector `triu`,
                       dimension `dim`), Σ and Σ|.| checksums, callback count
  cfg2_blocks.npz    : every callback block of config 2 in call order (keys, block offsets, values)
  au2_blocks.npz     : same for the first two atoms of the Au20 tetrahedron (two-centre cases, fallback)
  order_*_blocks.npz : callback blocks with a caller-supplied Cartesian component order (synth.shell_order)
  deriv1_*_blocks.npz: callback blocks of derivative order n = 1 (shifted-momentum blocks, src/libecp.c:246-250,322-369)
                       (scope row f1)
  deriv2_*_blocks.npz: derivative order n = 2: every callback block.  The (+1,0) / (0,+1) blocks (`mixed`) are computed at
                       l + 1 and shifted with the dimensions of l (src/libecp.c:362-369): the momentum-l block with
                       coefficients d zeta, IJK(la) x IJK(lb) elements.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from libecp_b200 import synth  # noqa: E402
from oracle.refbind import RefLib  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def save_matrix(ref, name, s):
    t = time.time()
    I = ref.get_integrals(s)
    dt = time.time() - t
    iu = np.triu_indices(s["dim"])
    assert np.all(np.tril(I, -1) == 0.0)
    np.savez_compressed(os.path.join(HERE, f"{name}_matrix.npz"), triu=I[iu], dim=s["dim"],
                        sum=I.sum(), sumabs=np.abs(I).sum(), ref_seconds=dt, nominal=synth.nominal_triples(s))
    print(f"{name}: dim {s['dim']} sum {I.sum():.15e} sumabs {np.abs(I).sum():.15e}  {dt:.1f}s")


def save_digest(ref, name, s):
    """19000 x 19000 partial matrices are too big to commit: keep row/column sums, checksums and a fixed
    sample of 40000 non-zero elements (flat index + value)."""
    t = time.time()
    M = ref.get_integrals(s)
    dt = time.time() - t
    nz = np.flatnonzero(M.ravel())
    rng = np.random.default_rng(12345)
    pick = np.sort(rng.choice(nz, size=min(40000, len(nz)), replace=False))
    np.savez_compressed(os.path.join(HERE, f"{name}_digest.npz"), dim=s["dim"], rowsum=M.sum(1), colsum=M.sum(0),
                        rowabs=np.abs(M).sum(1), sum=M.sum(), sumabs=np.abs(M).sum(), nnz=len(nz), sample_idx=pick,
                        sample_val=M.ravel()[pick], ref_seconds=dt, nominal=synth.nominal_triples(s))
    print(f"{name}: dim {s['dim']} sum {M.sum():.15e} sumabs {np.abs(M).sum():.15e} nnz {len(nz)}  {dt:.1f}s")


def save_blocks(ref, name, s, n=0, ordering=None, lmax=-1):
    rc, recs = ref.callbacks(s, n=n, ordering=ordering, lmax=lmax)
    assert rc == 0
    keys = np.array([r[:9] for r in recs], np.int32)
    off = np.cumsum([0] + [len(r[9]) for r in recs])
    vals = np.concatenate([r[9] for r in recs])
    extra = {}
    if n == 2:
        # (+1,0) / (0,+1) blocks: the reference evaluates chi / gamma at l + 1 and shifts with the dimensions of l
        # (src/libecp.c:362-369) - the lower-degree part of the larger array, i.e. the momentum-l block with coefficients
        # d zeta; the block handed to the callback has IJK(la) x IJK(lb) elements (refbind sizes it so; reading it with
        # the shifted size runs past the block - the "NaN" an earlier version of this script reported).  `mixed` marks them.
        mixed = np.zeros(len(vals), bool)
        for k, r in enumerate(recs):
            if (r[3], r[7]) in ((1, 0), (0, 1)):
                mixed[off[k]:off[k + 1]] = True
        extra["mixed"] = mixed
        print(f"{name}: {int(mixed.sum())} values in (+1,0)/(0,+1) blocks, {int(np.isnan(vals).sum())} NaN, "
              f"{int(np.isnan(vals[~mixed]).sum())} NaN elsewhere")
    np.savez_compressed(os.path.join(HERE, f"{name}_blocks.npz"), keys=keys, off=off, vals=vals, **extra)
    print(f"{name}: {len(recs)} callbacks, {len(vals)} values")


def main():
    ref = RefLib("ref")
    if "--deriv" in sys.argv:  # next scope row (f1): first derivatives, two shapes; nothing else is regenerated
        save_blocks(ref, "deriv1_tz2_L4", synth.deriv_pair(2, 4), n=1)
        save_blocks(ref, "deriv1_tz3_L5", synth.deriv_pair(3, 5), n=1)
        save_blocks(ref, "deriv1_triangle", synth.deriv_triangle(), n=1)
        return
    if "--deriv2" in sys.argv:  # second derivatives
        save_blocks(ref, "deriv2_tz1_L4", synth.deriv_pair(1, 4), n=2)
        save_blocks(ref, "deriv2_tz2_L5", synth.deriv_pair(2, 5), n=2)
        save_blocks(ref, "deriv2_triangle", synth.deriv2_triangle(), n=2)
        return
    if "--order" in sys.argv:  # scope row f4: caller-supplied Cartesian component order (src/libecp.c:152-166)
        save_blocks(ref, "order_rev_cfg2", synth.cfg2(), ordering=synth.shell_order(10, "reversed"), lmax=10)
        save_blocks(ref, "order_zfirst_au2", synth.cfg3(2), ordering=synth.shell_order(11, "zfirst"), lmax=11)
        return
    save_matrix(ref, "cfg1", synth.cfg1())
    save_matrix(ref, "cfg2", synth.cfg2())
    save_matrix(ref, "cfg2_L5", synth.cfg2(5))
    save_matrix(ref, "au2", synth.cfg3(2))
    save_matrix(ref, "au4", synth.cfg3(4))
    save_matrix(ref, "cfg4a", synth.cfg4("a"))
    save_matrix(ref, "cfg4b", synth.cfg4("b"))
    save_blocks(ref, "cfg1", synth.cfg1())
    save_blocks(ref, "cfg2", synth.cfg2())
    save_blocks(ref, "au2", synth.cfg3(2))
    # config 5: per-centre partial matrices of a 500-atom crystal, four representative centres
    for c in (0, 1, 288, 289):
        save_digest(ref, f"cfg5_c{c}", synth.cfg5(500, active=[c]))
    if "--big" in sys.argv:
        save_matrix(ref, "cfg3", synth.cfg3(20))


if __name__ == "__main__":
    main()
