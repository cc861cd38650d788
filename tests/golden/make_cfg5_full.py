"""Full 500-centre run of config 5 on the UNMODIFIED reference (oracle/_ref/libecp_ref.so) -> cfg5_full_digest.npz.

The reference is single-threaded; results are additive over ECP centres (masking centres through
``shellsECP`` is the API's own mechanism, src/libecp.c:95-127,257), so P worker processes pull centres from a queue,
each accumulates ``getIntegrals`` (+= into the caller's matrix, src/getIntegrals.c:40) into its own 19000 x 19000
matrix, and the parent adds the P partial matrices.  Summing partials in another order than the reference's centre-major
one perturbs last ulps only (tolerance of the parity tests: 1e-12 + 1e-10 |ref|).

The 2.9 GB matrix cannot be committed; the digest keeps
  rowsum, colsum, rowabs   : per AO row / column sums of the upper-triangular result
  ablk_sum, ablk_abs       : sum and sum|.| of every (atom A, atom B) block (500 x 500)
  sample_idx, sample_val   : 200 000 fixed non-zero elements (flat index, value)
  sum, sumabs, nnz

    python tests/golden/make_cfg5_full.py [nproc]          (about 6 minutes on 8 cores, 25 GB of RAM)
    python tests/golden/make_cfg5_full.py --variant fma     (same run with oracle/_ref/libecp_ref_fma.so, `make -C oracle
                                                             ref_fma`: the reference against itself, see oracle/Makefile)

Run in the build container only (needs oracle/_ref built from /root/reference by oracle/Makefile).
"""
import ctypes as C
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from libecp_b200 import parity, synth  # noqa: E402
from oracle.refbind import RefLib, _p, _pd, _pi  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
TMP = os.environ.get("CFG5_TMP", "/tmp")


VARIANT = None  # None: the -O2 build of oracle/Makefile (the oracle); "fma": the same sources, -mfma -ffp-contract=fast


def worker(p, queue, done):
    ref = RefLib("ref")
    if VARIANT == "fma":  # same binding, other build of the same unmodified sources
        ref.lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libecp_ref_fma.so"))
        ref.f_get = ref.lib.getIntegrals
        ref.f_get.restype = C.c_int
    full = synth.cfg5(500)
    dim = int(full["dim"])
    M = np.zeros((dim, dim))
    n = 0
    while True:
        c = queue.get()
        if c is None:
            break
        s = synth.cfg5(500, active=[c])
        rc = ref.f_get(C.c_int(s["nat"]), _p(s["geometry"], _pd), _p(s["shellsECP"], _pi), _p(s["KECP"], _pi),
                       _p(s["lECP"], _pi), _p(s["nECP"], _pd), _p(s["dECP"], _pd), _p(s["aECP"], _pd),
                       _p(s["shellsBS"], _pi), _p(s["lBS"], _pi), _p(s["KBS"], _pi), _p(s["dBS"], _pd), _p(s["aBS"], _pd),
                       C.c_int(synth.LARGE_GRID_ORDER), C.c_double(synth.TOLERANCE), C.c_double(synth.ACCURACY),
                       C.c_int(dim), _p(M, _pd))
        assert rc == 0
        n += 1
    np.save(os.path.join(TMP, f"cfg5_part_{p}.npy"), M)
    done.put((p, n))


def digest(M, s, seconds, nproc):
    assert np.all(np.isfinite(M))
    out = parity.make_digest(M, s, os.path.join(HERE, "cfg5_full_digest.npz"), ref_seconds=seconds, ref_procs=nproc,
                             nominal=synth.nominal_triples(s))
    print(f"cfg5 full: dim {out['dim']} sum {out['sum']:.15e} sumabs {out['sumabs']:.15e} nnz {out['nnz']} "
          f"lower-triangle zero: {bool(np.all(np.tril(M, -1) == 0.0))}  {seconds:.0f}s on {nproc} procs")


def variant_digest(M, seconds, nproc):
    """the reference against ITSELF: values of the other build at the sample positions of the committed digest"""
    z = np.load(os.path.join(HERE, "cfg5_full_digest.npz"))
    val = M.ravel()[z["sample_idx"]]
    ref = z["sample_val"]
    err = np.abs(val - ref)
    bad = err > 1e-12 + 1e-10 * np.abs(ref)
    np.savez_compressed(os.path.join(HERE, f"cfg5_full_variant_{VARIANT}.npz"), sample_val=val, rowsum=M.sum(1),
                        sum=M.sum(), sumabs=np.abs(M).sum(), ref_seconds=seconds, ref_procs=nproc,
                        build="gcc -O2 -fPIC -mfma -ffp-contract=fast (oracle/Makefile target ref_fma)")
    print(f"variant {VARIANT}: {int(bad.sum())} of {len(ref)} sampled elements outside 1e-12 + 1e-10|ref| of the -O2 build, "
          f"max |d| {err.max():.3e}, max rel {np.max(err / np.maximum(np.abs(ref), 1e-300)):.3e}")


def main():
    global VARIANT
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        VARIANT = sys.argv[i + 1]
        del sys.argv[i:i + 2]
    nproc = int(sys.argv[1]) if len(sys.argv) > 1 else (os.cpu_count() or 8)
    s = synth.cfg5(500)
    centres = [i for i in range(int(s["nat"])) if s["shellsECP"][i] > 0]
    # interior centres are the expensive ones: hand them out first
    xyz = s["geometry"].reshape(-1, 3)
    mid = xyz.mean(0)
    centres.sort(key=lambda c: float(((xyz[c] - mid) ** 2).sum()))
    queue, done = mp.Queue(), mp.Queue()
    for c in centres:
        queue.put(c)
    for _ in range(nproc):
        queue.put(None)
    t0 = time.time()
    procs = [mp.Process(target=worker, args=(p, queue, done)) for p in range(nproc)]
    for p in procs:
        p.start()
    got = [done.get() for _ in procs]
    for p in procs:
        p.join()
    seconds = time.time() - t0
    assert sum(n for _, n in got) == len(centres)
    M = None
    for p in range(nproc):
        f = os.path.join(TMP, f"cfg5_part_{p}.npy")
        part = np.load(f)
        M = part if M is None else M.__iadd__(part)
        os.remove(f)
    if VARIANT:
        variant_digest(M, seconds, nproc)
    else:
        digest(M, s, seconds, nproc)


if __name__ == "__main__":
    main()
