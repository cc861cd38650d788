"""GPU tier (-m gpu): parity of the CUDA path, called through the C ABI (include/libecp.h,
getIntegrals.h, libecp_b200.h), with the oracle and the golden fixtures of the compiled reference.

Tolerance (north_star): |x - ref| <= 1e-12 + 1e-10 |ref| element-wise, identical screening decisions
and integral indexing (the callback key sequence must be identical)."""
import ctypes
import os

import numpy as np
import pytest
from conftest import GOLDEN, assert_parity, load_blocks, load_matrix

from libecp_b200 import capi, synth
from oracle.refbind import RefLib, have

pytestmark = pytest.mark.gpu

SMALL = {"cfg1": synth.cfg1, "cfg2": synth.cfg2, "cfg2_L5": lambda: synth.cfg2(5), "au2": lambda: synth.cfg3(2),
         "au4": lambda: synth.cfg3(4), "cfg4a": lambda: synth.cfg4("a"), "cfg4b": lambda: synth.cfg4("b")}


@pytest.fixture(scope="module")
def oracle():
    return RefLib("ref" if have("ref") else "port")


@pytest.mark.parametrize("name", list(SMALL))
def test_getintegrals_matches_golden(name):
    """one-call interface on host buffers vs the reference's matrix (all five configs' small forms)"""
    got = capi.get_integrals(SMALL[name]())
    ref = load_matrix(name)
    assert_parity(got, ref, name)
    assert np.all(np.tril(got, -1) == 0.0)  # never writes the strict lower triangle


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "au2"])
def test_callbacks_match_golden_blocks(name):
    """calculateECPIntegrals: same callback sequence (A,s1,la,shifta,B,s2,lb,shiftb,C) and blocks"""
    keys, off, vals = load_blocks(name)
    with capi.Handle(SMALL[name]()) as h:
        rc, recs = h.callbacks()
    assert rc == 0 and len(recs) == len(keys)
    for k, r in enumerate(recs):
        assert tuple(keys[k]) == r[:9]
    assert_parity(np.concatenate([r[9] for r in recs]), vals, name)


@pytest.mark.parametrize("name", ["cfg4a", "cfg4b", "au4"])
def test_callbacks_match_oracle(oracle, name):
    s = SMALL[name]()
    rc_o, ro = oracle.callbacks(s)
    with capi.Handle(s) as h:
        rc, rg = h.callbacks()
        st = h.stats()
    assert rc == rc_o == 0 and len(ro) == len(rg)
    assert all(a[:9] == b[:9] for a, b in zip(ro, rg))
    assert_parity(np.concatenate([r[9] for r in rg]), np.concatenate([r[9] for r in ro]), name)
    assert st["executed_triples"] * 2 == len(rg)
    assert st["stale_centre_events"] == 0


def test_config3_au20_full_matrix():
    """Au20, 860 AOs, 678 600 nominal / 198 152 executed triples, vs the reference's matrix"""
    s = synth.cfg3(20)
    with capi.Handle(s) as h:
        rc, M = h.integrals_host()
        st = h.stats()
    assert rc == 0
    assert st["nominal_triples"] == 678600 and st["executed_triples"] == 198152
    assert_parity(M, load_matrix("cfg3"), "cfg3")


@pytest.mark.parametrize("centre", [0, 1, 288, 289])
def test_config5_single_centre_digest(centre):
    """500-atom PbS crystal with one active centre: row/column sums and 40 000 sampled elements of the
    19 000 x 19 000 partial matrix vs the reference"""
    z = np.load(os.path.join(GOLDEN, f"cfg5_c{centre}_digest.npz"))
    s = synth.cfg5(500, active=[centre])
    with capi.Handle(s) as h:
        rc, M = h.integrals_host()
    assert rc == 0
    # identical screening => identical row support (exact zeros inside a block may differ in the last ulp)
    assert np.array_equal(np.abs(M).sum(1) > 0, z["rowabs"] > 0)
    assert abs(int((M != 0).sum()) - int(z["nnz"])) <= 0.02 * int(z["nnz"])
    assert_parity(M.ravel()[z["sample_idx"]], z["sample_val"], f"cfg5 centre {centre} sample")
    scale = z["rowabs"]
    assert np.all(np.abs(M.sum(1) - z["rowsum"]) <= 1e-11 + 1e-10 * scale)
    assert abs(M.sum() - float(z["sum"])) <= 1e-9 * float(z["sumabs"])


def test_properties_at_scale():
    """size-independent properties on a 60-atom slice of config 5 (two ECP types, ~1.3 M executed triples):
    additivity over centres, shard union == unsharded, zero lower triangle, batch-size independence"""
    s = synth.cfg5(60)
    with capi.Handle(s) as h:
        rc, full = h.integrals_host()
        n_exec = h.stats()["executed_triples"]
        assert rc == 0 and n_exec > 100000
        assert np.all(np.tril(full, -1) == 0.0)
        acc = np.zeros_like(full)
        tot = 0
        for rank in range(4):
            h.set_shard(rank, 4)
            rc, part = h.integrals_host()
            assert rc == 0
            assert not np.any((part != 0) & (acc != 0))  # disjoint output blocks
            acc += part
            tot += h.stats()["executed_triples"]
        assert tot == n_exec
        assert_parity(acc, full, "shard union")
    os.environ["LIBECP_B200_BATCH_TRIPLES"] = "50000"
    try:
        with capi.Handle(s) as h:
            rc, small_batches = h.integrals_host()
            assert h.stats()["batches"] > 5
    finally:
        del os.environ["LIBECP_B200_BATCH_TRIPLES"]
    assert_parity(small_batches, full, "batching")
    half = [i for i in range(60) if i % 2 == 0]
    other = [i for i in range(60) if i % 2 == 1]
    a = capi.get_integrals(synth.mask_centres(s, half))
    b = capi.get_integrals(synth.mask_centres(s, other))
    assert_parity(a + b, full, "additivity over centres")


def test_structural_invariants():
    with capi.Handle(synth.cfg3(3)) as h:
        rc, recs = h.callbacks()
    assert rc == 0
    for (A, s1, la, _, B, s2, lb, _, C, blk) in recs:
        if A == B == C and (la + lb) % 2 == 1:
            assert np.all(np.abs(blk) <= 1e-13)
        if A == B and s1 == s2:
            n = (la + 1) * (la + 2) // 2
            m = blk.reshape(n, n)
            assert np.allclose(m, m.T, rtol=1e-10, atol=1e-12)


def test_edge_cases():
    # atom without ECP in the middle, atom without basis functions, ragged shells
    s = synth.cfg3(3)
    s2 = synth.mask_centres(s, [0, 2])
    ref = RefLib("ref" if have("ref") else "port").get_integrals(s2)
    assert_parity(capi.get_integrals(s2), ref, "masked centre")
    # no ECP at all: nothing to do, matrix untouched
    s0 = synth.mask_centres(s, [])
    assert np.all(capi.get_integrals(s0) == 0.0)
    # getIntegrals accumulates (+=) into a pre-filled upper triangle
    with capi.Handle(synth.cfg2()) as h:
        rc, M = h.integrals_host()
        rc, M2 = h.integrals_host()  # handle is reusable
    assert np.array_equal(M, M2) or np.allclose(M, M2, rtol=1e-13, atol=1e-14)


def test_device_resident_matrix_matches_host_copy():
    """the device-resident result (libecp_b200_integrals_device) is read through its device pointer"""
    from libecp_b200 import parity

    s = synth.cfg3(2)
    with capi.Handle(s) as h:
        rc, ptr, n = h.integrals_device()
        assert rc == 0 and n == s["dim"] and ptr and ptr == h.matrix_ptr()
        D = parity.device_view(ptr, n).cpu().numpy().copy()
        rc, M = h.integrals_host()
    assert_parity(D, load_matrix("au2"), "device-resident matrix")
    assert np.all(np.tril(D, -1) == 0.0)
    assert np.allclose(D, M, rtol=1e-13, atol=1e-15)  # host copy of a second pass (atomicAdd order is not fixed)


def test_config5_full_digest():
    """the benchmarked computation itself: one full pass over the 500-centre crystal (9.77e9 nominal / 1.85e7 executed
    triples, several 3 M-triple batches), the device-resident 19 000 x 19 000 matrix against the digest of a full run
    of the unmodified reference (tests/golden/make_cfg5_full.py)"""
    from libecp_b200 import parity

    s = synth.cfg5(500)
    with capi.Handle(s) as h:
        rc, ptr, n = h.integrals_device()
        st = h.stats()
        assert rc == 0 and n == 19000
        res = parity.check_digest(parity.device_view(ptr, n), s)
    assert st["nominal_triples"] == 9767187500 and st["batches"] >= 3
    assert res["ok"], res


def test_config5_host_consumer_digest():
    """the e2e path of the benchmark: libecp_b200_integrals_host on the full 500-centre crystal - sparse download of the
    non-zero runs, rows streamed out while the pass goes on - fills the caller's host matrix; that matrix against the
    digest of the full run of the unmodified reference; a third of the dense bytes cross PCIe"""
    import torch

    from libecp_b200 import parity

    s = synth.cfg5(500)
    dim = int(s["dim"])
    host = np.zeros((dim, dim))
    with capi.Handle(s) as h:
        rc = capi.lib().libecp_b200_integrals_host(ctypes.c_void_p(h.h), dim, host.ctypes.data_as(capi._pd))
        st = h.stats()
    assert rc == 0 and st["batches"] >= 4
    assert 0 < st["d2h_bytes"] < 0.4 * (dim * (dim + 1) // 2) * 8
    res = parity.check_digest(torch.from_numpy(host).cuda(), s)
    assert res["ok"], res


@pytest.mark.parametrize("world", [2, 8])
def test_config5_sharded_gather_digest(world):
    """multi-GPU decomposition of the benchmarked config on one GPU: `world` handles play the ranks (row ownership,
    libecp_b200_set_shard), every shard is packed and scattered into rank 0's device matrix (the device side of the
    NCCL all-gather, libecp_b200_pack_rows / _unpack_rows), and the gathered matrix is checked against the reference
    digest; executed triples of the shards add up to the unsharded count"""
    import torch

    from libecp_b200 import gather, parity

    s = synth.cfg5(500)
    n = int(s["dim"])
    h0 = capi.Handle(s)
    try:
        h0.set_shard(0, world)
        rows, sizes = gather.shard_layout(h0, world)
        assert sum(sizes) == n * (n + 1) // 2
        assert h0.integrals_device()[0] == 0
        executed = h0.stats()["executed_triples"]
        buf = torch.empty(max(sizes), dtype=torch.float64, device="cuda")
        for r in range(1, world):
            with capi.Handle(s) as hr:
                hr.set_shard(r, world)
                assert hr.integrals_device()[0] == 0
                executed += hr.stats()["executed_triples"]
                assert hr.pack_rows(rows[r], buf.data_ptr(), buf.numel()) == sizes[r]
            h0.unpack_rows(rows[r], buf.data_ptr(), buf.numel())
        res = parity.check_digest(parity.device_view(h0.matrix_ptr(), n), s)
    finally:
        h0.close()
    assert executed == 18509218, executed  # the unsharded pass (bench line, BENCH_r01.json)
    assert res["ok"], res


def _with_env(env, fn):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        return fn()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def test_launch_structure_does_not_change_results():
    """the host pipeline (builder thread), the two-launch fast path (incl. the overflow of its survivor list) and
    the block sizes of the persistent kernels are scheduling only: bit-identical matrices"""
    s = synth.cfg3(4)
    base = capi.get_integrals(s)
    assert_parity(base, load_matrix("au4"), "au4")
    for env in ({"LIBECP_B200_NO_PIPELINE": "1"}, {"LIBECP_B200_FASTLIM": "13"}, {"LIBECP_B200_FASTLIM": "7"},
                {"LIBECP_B200_SURVCAP": "100"}, {"LIBECP_B200_T1BLOCK": "32", "LIBECP_B200_FBBLOCK": "128"},
                {"LIBECP_B200_T1BLOCK": "128", "LIBECP_B200_FBBLOCK": "32", "LIBECP_B200_BATCH_TRIPLES": "500"}):
        got = _with_env(env, lambda: capi.get_integrals(s))
        # atomicAdd order into the matrix is not fixed: compare to the last ulps, not bitwise
        assert np.allclose(got, base, rtol=1e-13, atol=1e-15), env


def test_host_consumer_in_row_panels():
    """libecp_b200_integrals_host cuts its pass into row panels whose download and host += overlap the next panel's
    compute (LIBECP_B200_HOST_PANELS; default 3 for large matrices, 1 here): same matrix, also for a sharded handle,
    and += semantics into a pre-filled caller matrix"""
    for s, name in ((synth.cfg3(4), "au4"), (synth.cfg5(40), None)):
        base = _with_env({"LIBECP_B200_HOST_PANELS": "1"}, lambda: capi.get_integrals(s))
        if name:
            assert_parity(base, load_matrix(name), name)
        for P in ("2", "3", "5"):
            got = _with_env({"LIBECP_B200_HOST_PANELS": P}, lambda: capi.get_integrals(s))
            assert np.allclose(got, base, rtol=1e-13, atol=1e-15), (name, P)
            assert np.all(np.tril(got, -1) == 0.0)

        def sharded():
            acc = np.zeros_like(base)
            with capi.Handle(s) as h:
                for rank in range(2):
                    h.set_shard(rank, 2)
                    rc, part = h.integrals_host()
                    assert rc == 0 and not np.any((part != 0) & (acc != 0))
                    acc += part
                st = h.stats()
            return acc, st
        got, st = _with_env({"LIBECP_B200_HOST_PANELS": "3"}, sharded)
        assert np.allclose(got, base, rtol=1e-13, atol=1e-15), name
        assert st["batches"] >= 3


def test_sparse_download_equals_dense_download():
    """the host consumer moves only the non-zero 16-element runs of the result rows plus one bit per run
    (ecp_cuda.cu: matrix_add_to_host_sparse; LIBECP_B200_D2H=dense keeps the dense upper-triangle panels): the caller's
    matrix is the same (bit for bit on a single centre) - small matrices (rows shorter than a run), a sparse one (distant atoms), row panels,
    a sharded handle, and += into a pre-filled matrix; fewer bytes cross PCIe"""
    # two distant ECP atoms, a neighbour without ECP and an atom no centre reaches (its rows are never downloaded)
    far = synth.assemble("far", [(0.0, 0.0, 0.0), (0.0, 0.0, 40.0), (3.0, 0.5, 0.2), (60.0, 60.0, 0.0)], [synth.tz_basis(2)] * 4,
                         [synth.ecp_set(3), synth.ecp_set(3), None, None])
    for s in (synth.cfg1(), synth.cfg3(4), far, synth.cfg5(40)):
        def run(shard=None):
            dim = int(s["dim"])
            acc = np.full((dim, dim), 0.25)
            with capi.Handle(s) as h:
                if shard:
                    h.set_shard(*shard)
                rc = capi.lib().libecp_b200_integrals_host(ctypes.c_void_p(h.h), dim, acc.ctypes.data_as(capi._pd))
                st = h.stats()
            assert rc == 0
            return acc, st["d2h_bytes"]
        for P in ("1", "3"):
            for shard in (None, (1, 2)):
                dense, bd = _with_env({"LIBECP_B200_D2H": "dense", "LIBECP_B200_HOST_PANELS": P}, lambda: run(shard))
                sparse, bs = _with_env({"LIBECP_B200_HOST_PANELS": P}, lambda: run(shard))
                # two runs differ in the last bit where several centres add into an element (atomicAdd order); the
                # set of touched elements is the same and nothing else may differ
                assert np.array_equal(dense != 0.25, sparse != 0.25), (s.get("name"), P, shard)
                assert np.allclose(dense, sparse, rtol=1e-13, atol=1e-15), (s.get("name"), P, shard)
                assert np.any(dense != 0.25)
        if s is far:
            assert bs < 0.8 * bd


def test_streamed_download_equals_download_after_the_pass():
    """the host consumer hands the rows of atoms no later centre can reach to a download thread while the pass goes on
    (api.c: stream_after_batch; LIBECP_B200_STREAM_D2H=0 downloads everything after the pass): same matrix, also sharded,
    with many small batches so that several downloads really overlap the pass"""
    for s in (synth.cfg5(60), synth.cfg3(4)):
        def run(shard=None):
            dim = int(s["dim"])
            acc = np.full((dim, dim), -0.5)
            with capi.Handle(s) as h:
                if shard:
                    h.set_shard(*shard)
                rc = capi.lib().libecp_b200_integrals_host(ctypes.c_void_p(h.h), dim, acc.ctypes.data_as(capi._pd))
                st = h.stats()
            assert rc == 0
            return acc, st
        env = {"LIBECP_B200_BATCH_TRIPLES": "20000" if s["nat"] > 4 else "300", "LIBECP_B200_STREAM_MIN_BYTES": "1"}
        for shard in (None, (0, 2)):
            ref, st0 = _with_env(dict(env, LIBECP_B200_STREAM_D2H="0"), lambda: run(shard))
            got, st1 = _with_env(env, lambda: run(shard))
            assert st1["batches"] >= 3
            assert np.array_equal(ref != -0.5, got != -0.5), (s.get("name"), shard)
            assert np.allclose(ref, got, rtol=1e-13, atol=1e-15), (s.get("name"), shard)
            assert np.all(got[np.tril_indices(got.shape[0], -1)] == -0.5)


def test_handles_in_sequence_reuse_parked_buffers():
    """libECP_free parks the device scratch for the next handle: shapes of different size in sequence, then an
    explicit release, still give the reference's matrices"""
    for name in ("cfg2", "au4", "cfg4b", "cfg1", "au2", "cfg4a"):
        assert_parity(capi.get_integrals(SMALL[name]()), load_matrix(name), name)
    capi.lib().libecp_b200_release_cache()
    assert_parity(capi.get_integrals(SMALL["au2"]()), load_matrix("au2"), "au2 after release")


def test_gather_pack_unpack_rebuilds_the_full_matrix():
    """device-resident gather (libecp_b200_pack_rows / _unpack_rows): two handles play rank 0 and rank 1 of a 2-way
    shard on one GPU; rank 1's packed rows scattered into rank 0's device matrix give the unsharded result, and the
    next pass of that handle starts from a clean matrix again"""
    import torch

    from libecp_b200 import gather

    s = synth.cfg3(4)
    ref = load_matrix("au4")
    n = int(s["dim"])
    with capi.Handle(s) as h0, capi.Handle(s) as h1:
        h0.set_shard(0, 2)
        h1.set_shard(1, 2)
        rows, sizes = gather.shard_layout(h0, 2)
        assert sum(sizes) == n * (n + 1) // 2
        assert h0.integrals_device()[0] == 0 and h1.integrals_device()[0] == 0
        buf = torch.zeros(max(sizes), dtype=torch.float64, device="cuda")
        assert h1.pack_rows(rows[1], buf.data_ptr(), buf.numel()) == sizes[1]
        torch.cuda.synchronize()
        h0.unpack_rows(rows[1], buf.data_ptr(), buf.numel())
        full = np.zeros((n, n))
        assert h0.pack_rows(np.arange(n, dtype=np.int32), 0, 0) == n * (n + 1) // 2  # sizing call only
        allbuf = torch.zeros(n * (n + 1) // 2, dtype=torch.float64, device="cuda")
        h0.pack_rows(np.arange(n, dtype=np.int32), allbuf.data_ptr(), allbuf.numel())
        full[np.triu_indices(n)] = allbuf.cpu().numpy()
        assert_parity(full, ref, "gathered au4")
        # too small a buffer is refused, not overrun
        with pytest.raises(RuntimeError):
            h0.pack_rows(rows[0], buf.data_ptr(), 1)
        # the scatter made other ranks' rows non-zero: the next pass must not accumulate onto them
        rc, M = h0.integrals_host()
        rc1, M1 = h1.integrals_host()
        assert_parity(M + M1, ref, "shard union after a gather")


def test_fallback_kernels_agree():
    """the level-wave fallback (default, ecp_waves.cuh) and the persistent 8-lane kernel k_fallbackG (LIBECP_B200_FB=group)
    evaluate the same integrand values and differ only in the association of the sums inside a level: T equal to
    ~1e-15 relative, matrices within the parity tolerance of the reference, same items, same return code"""
    for fn, name in ((lambda: synth.cfg3(4), "au4"), (lambda: synth.cfg4("a"), "cfg4a"), (lambda: synth.cfg4("b"), "cfg4b"),
                     (lambda: synth.cfg5(24), None)):
        def run():
            with capi.Handle(fn()) as h:
                rc, M = h.integrals_host()
                st = h.stats()
                return rc, M, h.debug_fetch("T", 300000), st["fallback_items"]
        rc0, base, t0, n0 = _with_env({"LIBECP_B200_FB": "group"}, run)
        rc1, got, t1, n1 = run()
        assert rc0 == rc1 == 0 and n0 == n1 and n0 > 0
        if name:
            assert_parity(got, load_matrix(name), name)
        assert np.allclose(got, base, rtol=1e-12, atol=1e-14), name
        assert np.allclose(t1, t0, rtol=1e-11, atol=1e-14), (name, np.abs(t1 - t0).max())


def test_link_variants_are_bit_identical():
    """k_link4 (default for the large classes: Omega slices staged per run of triples that share both atoms, strides
    and loops fixed at compile time) performs the per-element operations of k_link (LIBECP_B200_LINK=global) in the same
    order: gamma must be bit-identical, for any triples-per-block setting; L = 2, 4 (config 5), 4 (Au), 5 (g stress)"""
    for s, name in ((synth.cfg3(4), "au4"), (synth.cfg4("a"), "cfg4a"), (synth.cfg5(24), None)):
        def run():
            with capi.Handle(s) as h:
                rc, M = h.integrals_host()
                return M, h.debug_fetch("gamma", 400000)
        base, g0 = _with_env({"LIBECP_B200_LINK": "global"}, run)
        if name:
            assert_parity(base, load_matrix(name), name)
        for env in ({}, {"LIBECP_B200_LINKTPB": "1"}, {"LIBECP_B200_LINKTPB": "5"}, {"LIBECP_B200_LINKTPB": "64"}):
            got, g1 = _with_env(env, run)
            assert np.array_equal(g0, g1), (name, env)
            assert np.allclose(got, base, rtol=1e-13, atol=1e-15), (name, env)


def test_type1_wave_is_bit_identical_to_the_one_kernel_path():
    """k_type1A (first 16 slots of every primitive pair as a block-wide wave: thread per live point, thread per
    (pair, quadrature) for the bookkeeping of levels 0..3) + k_type1S on the open pairs from level 4 performs the
    operations of k_type1S alone (LIBECP_B200_T1=legacy) in the same order: Q must be bit-identical; L = 2, 4 (config 5),
    4 (Au), g / h stress shapes (LAB up to 10), several batches"""
    for s, name in ((synth.cfg3(4), "au4"), (synth.cfg4("a"), "cfg4a"), (synth.cfg4("b"), "cfg4b"), (synth.cfg5(24), None),
                    (synth.cfg2(5), "cfg2_L5")):
        def run():
            with capi.Handle(s) as h:
                rc, M = h.integrals_host()
                st = h.stats()
                return M, h.debug_fetch("Q", 2000000), st
        for env in ({}, {"LIBECP_B200_BATCH_TRIPLES": "700"}):
            base, q0, st0 = _with_env(dict(env, LIBECP_B200_T1="legacy"), run)
            got, q1, st1 = _with_env(env, run)
            assert np.abs(q0).sum() > 0 and np.array_equal(q0, q1), (name, env)
            assert np.allclose(got, base, rtol=1e-13, atol=1e-15), (name, env)
            assert st0["t1_large_pairs"] == st1["t1_large_pairs"] if "t1_large_pairs" in st0 else True
        if name:
            assert_parity(got, load_matrix(name), name)


def test_shift_kernels_are_bit_identical():
    """k_shift2 (default: both binomial-shift passes in one kernel, J in shared memory, factors per (triple, term)) runs
    the terms of k_shiftJ / k_shiftI (LIBECP_B200_SHIFT=two) in the same order with the same fma's: callback blocks
    bit-identical, matrices equal to the last ulps (atomicAdd order)"""
    for s, name in ((synth.cfg3(3), None), (synth.cfg4("a"), "cfg4a"), (synth.cfg4("b"), "cfg4b"), (synth.cfg5(16), None)):
        def run():
            with capi.Handle(s) as h:
                rc, recs = h.callbacks()
                rc2, M = h.integrals_host()
                return rc, rc2, np.concatenate([r[9] for r in recs]), M
        rc0, rc0b, blk0, M0 = _with_env({"LIBECP_B200_SHIFT": "two"}, run)
        rc1, rc1b, blk1, M1 = run()
        assert rc0 == rc1 == rc0b == rc1b == 0
        assert np.array_equal(blk0, blk1), name
        assert np.allclose(M0, M1, rtol=1e-13, atol=1e-15), name
        if name:
            assert_parity(M1, load_matrix(name), name)


def test_ftab_variants_are_bit_identical():
    """k_Ftab2 (default) tabulates only the window of every shell slot into a cleared table, k_Ftab (LIBECP_B200_FTAB=full)
    the whole grid with zeros outside the window: same arithmetic per point, so the F table and the matrices are
    bit-identical"""
    for s, name in ((synth.cfg3(4), "au4"), (synth.cfg4("b"), "cfg4b")):
        def run():
            with capi.Handle(s) as h:
                rc, M = h.integrals_host()
                return M, h.debug_fetch("F", 400000)
        base, f0 = _with_env({"LIBECP_B200_FTAB": "full"}, run)
        got, f1 = run()
        assert np.array_equal(f0, f1), name
        assert np.allclose(got, base, rtol=1e-13, atol=1e-15), name
        assert_parity(got, load_matrix(name), name + " compact F")


ORDER_CASES = {"order_rev_cfg2": (synth.cfg2, "reversed", 10), "order_zfirst_au2": (lambda: synth.cfg3(2), "zfirst", 11)}


@pytest.mark.parametrize("name", list(ORDER_CASES))
def test_custom_shell_ordering_matches_golden(name):
    """scope row f4: caller-supplied Cartesian component order (libECP_init(..., lmax, shellOrdering, ...), reference
    src/libecp.c:152-166) vs blocks of the compiled reference run with the same order; and the blocks are the
    default-order blocks with rows / columns permuted"""
    mk, kind, lmax = ORDER_CASES[name]
    keys, off, vals = load_blocks(name)
    s = mk()
    order = synth.shell_order(lmax, kind)
    with capi.Handle(s, ordering=order, lmax=lmax) as h:
        rc, recs = h.callbacks()
    assert rc == 0 and len(recs) == len(keys)
    for k, r in enumerate(recs):
        assert tuple(keys[k]) == r[:9]
    assert_parity(np.concatenate([r[9] for r in recs]), vals, name)
    with capi.Handle(s) as h:
        rc, std = h.callbacks()
    libint = synth.shell_order(lmax, "libint").reshape(-1, 3)
    mine = order.reshape(-1, 3)

    def perm(l):  # position in the default order of every component of the custom order
        lo, n = l * (l + 1) * (l + 2) // 6, (l + 1) * (l + 2) // 2
        where = {tuple(e): i for i, e in enumerate(libint[lo:lo + n])}
        return np.array([where[tuple(e)] for e in mine[lo:lo + n]])

    for a, b in zip(recs, std):
        la, lb = a[2], a[6]
        want = b[9].reshape(len(perm(la)), len(perm(lb)))[np.ix_(perm(la), perm(lb))].ravel()
        assert np.all(np.abs(a[9] - want) <= 1e-12 + 1e-10 * np.abs(want))


def test_custom_shell_ordering_too_short_is_rejected():
    """reference src/libecp.c:159-162: lmax < maxLambda + maxAlpha + 1 -> NULL"""
    with pytest.raises(RuntimeError):
        capi.Handle(synth.cfg2(), ordering=synth.shell_order(9, "reversed"), lmax=9)


@pytest.mark.parametrize("seed", range(8))
def test_random_small_systems_match_oracle(oracle, seed):
    """randomized parity: 2-4 atoms at random positions, TZ(0..3) bases, ECP(L <= 5) on a random subset of the atoms with
    random exponent scales (synth.random_system) - matrix and callback sequence against the oracle run live"""
    s = synth.random_system(seed)
    ref = oracle.get_integrals(s)
    got = capi.get_integrals(s)
    assert_parity(got, ref, s["name"])
    assert np.all(np.tril(got, -1) == 0.0)
    rc_o, ro = oracle.callbacks(s, keep_blocks=False)
    with capi.Handle(s) as h:
        rc, rg = h.callbacks(keep_blocks=False)
    assert rc == rc_o == 0 and [r[:9] for r in rg] == [r[:9] for r in ro]


def test_first_derivative_blocks_triangle():
    """row f1 on three atoms in general position with mixed shapes and one atom without ECP (synth.deriv_triangle):
    triples with A != B != C, both ECP types, every skip rule of src/libecp.c:303-330"""
    keys, off, vals = load_blocks("deriv1_triangle")
    with capi.Handle(synth.deriv_triangle(), n=1) as h:
        rc, recs = h.callbacks()
    assert rc == 0 and len(recs) == len(keys)
    assert all(tuple(keys[k]) == r[:9] for k, r in enumerate(recs))
    assert_parity(np.concatenate([r[9] for r in recs]), vals, "deriv1_triangle")


@pytest.mark.parametrize("name,lbs,L", [("deriv1_tz2_L4", 2, 4), ("deriv1_tz3_L5", 3, 5)])
def test_first_derivative_blocks_match_golden(name, lbs, L):
    """scope row f1: derivative order n = 1 - the shifted-momentum blocks handed to the callback
    (reference src/libecp.c:246-250,322-373; zeta factors src/type1.c:239-246, src/type2.c:263-269,459-462) vs the
    blocks of the compiled reference, same key sequence, same tolerance"""
    keys, off, vals = load_blocks(name)
    with capi.Handle(synth.deriv_pair(lbs, L), n=1) as h:
        rc, recs = h.callbacks()
        with pytest.raises(RuntimeError):
            h.integrals_host()  # the matrix consumer is the n = 0 one-call interface
    assert rc == 0 and len(recs) == len(keys)
    for k, r in enumerate(recs):
        assert tuple(keys[k]) == r[:9]
        assert len(r[9]) == off[k + 1] - off[k]
    assert_parity(np.concatenate([r[9] for r in recs]), vals, name)


def test_first_derivative_is_the_gradient_of_the_integrals():
    """size-independent property of row f1: d/dA <a|U_C|b> assembled from the shifted blocks,
    2 zeta <a+1|U|b> - a <a-1|U|b> per Cartesian direction, equals the central finite difference of the n = 0 blocks
    when atom A moves (one s-p pair of a two-atom system, h = 1e-4 bohr)"""
    base = synth.deriv_pair(2, 4)

    def blocks(s, n):
        with capi.Handle(s, n=n) as h:
            rc, recs = h.callbacks()
        assert rc == 0
        return recs

    recs = blocks(base, 1)
    # shell 5 of atom 0 (the most diffuse s) with shell 6 of atom 1 (first p shell), centre C = 1: d/dA_x of the s function
    # = 2 zeta * (p_x function with the same exponents): the +1 block holds coefficients d zeta, so grad = 2 * block
    tgt = [r for r in recs if r[:9] == (0, 5, 0, 1, 1, 6, 1, 0, 1)]
    assert len(tgt) == 2
    grad = 2.0 * (tgt[0][9] + tgt[1][9]).reshape(3, 3)  # rows: direction x, y, z of the shifted p function
    h = 1e-4
    fd = np.zeros((3, 3))
    for d in range(3):
        vals = []
        for sgn in (+1, -1):
            s = dict(base)
            g = base["geometry"].copy()
            g[d] += sgn * h
            s["geometry"] = g
            r0 = [r for r in blocks(s, 0) if r[:9] == (0, 5, 0, 0, 1, 6, 1, 0, 1)]
            vals.append((r0[0][9] + r0[1][9]).reshape(1, 3)[0])
        fd[d] = (vals[0] - vals[1]) / (2 * h)
    assert np.allclose(grad, fd, rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("name,mk", [("deriv2_tz1_L4", lambda: synth.deriv_pair(1, 4)), ("deriv2_tz2_L5", lambda: synth.deriv_pair(2, 5)),
                                     ("deriv2_triangle", synth.deriv2_triangle)])
def test_second_derivative_blocks_match_golden(name, mk):
    """scope row f1, n = 2: the ten shifted blocks per executed shell pair (reference src/libecp.c:246-250) - same key
    sequence, block sizes and values as the compiled reference at the usual tolerance, including the (+1,0) / (0,+1)
    blocks (momentum unchanged, coefficients d zeta: the reference evaluates them at l + 1 and shifts the lower-degree
    part, src/libecp.c:362-369; here they are ordinary blocks of a copy of the shell)."""
    keys, off, vals = load_blocks(name)
    with capi.Handle(mk(), n=2) as h:
        rc, recs = h.callbacks()
    assert rc == 0 and len(recs) == len(keys)
    for k, r in enumerate(recs):
        assert tuple(keys[k]) == r[:9]
        assert len(r[9]) == off[k + 1] - off[k]
    assert np.isfinite(vals).all()
    assert_parity(np.concatenate([r[9] for r in recs]), vals, name)


def _comp_index(l):
    """component (nx, ny, nz) -> position in the shell, default (libint) order"""
    o = synth.shell_order(l, "libint").reshape(-1, 3)
    first = sum((k + 1) * (k + 2) // 2 for k in range(l))
    return {tuple(int(x) for x in o[first + c]): c for c in range((l + 1) * (l + 2) // 2)}


def test_second_derivative_is_the_gradient_of_the_first():
    """size-independent property of row f1, n = 2, and the definition of its (+1,0) / (0,+1) blocks: the Hessian blocks
    d2/dA_i dA_j, d2/dA_i dB_j of <a|U_C|b> assembled from the n = 2 callbacks,
        D_i D_j g(a) = 4 z^2 g(a+1i+1j) - 2 z (a_i + d_ij) g(a+1j-1i) - 2 z a_j g(a-1j+1i) + a_j (a_i - d_ij) g(a-1j-1i)
    (the two middle terms are the momentum-l blocks with coefficients d zeta), equal central finite differences of the
    gradients assembled from n = 1 callbacks when atom A moves (three atoms, A != B != C, a p shell against a p shell)."""
    base = synth.deriv2_triangle()
    A, B, Cc = 0, 2, 1
    ls = list(base["lBS"])
    first = np.concatenate([[0], np.cumsum(base["shellsBS"])])
    s1 = [j for j in range(int(base["shellsBS"][A])) if ls[first[A] + j] == 1][-1]  # most diffuse p shell of atom A
    s2 = [j for j in range(int(base["shellsBS"][B])) if ls[first[B] + j] == 0][-1]  # most diffuse s shell of atom B
    la, lb = 1, 0

    def blocks(s, n):
        with capi.Handle(s, n=n) as h:
            rc, recs = h.callbacks()
        assert rc == 0
        out = {}
        for r in recs:
            if (r[0], r[1], r[4], r[5], r[8]) == (A, s1, B, s2, Cc):
                ea, eb = (0, 0) if n == 2 and (r[3], r[7]) in ((1, 0), (0, 1)) else (r[3], r[7])
                shp = ((la + ea + 1) * (la + ea + 2) // 2, (lb + eb + 1) * (lb + eb + 2) // 2)
                out[(r[3], r[7])] = out.get((r[3], r[7]), 0.0) + r[9].reshape(shp)  # type 1 + type 2
        return out

    ia = {l: _comp_index(l) for l in range(0, 4)}
    e = np.eye(3, dtype=int)

    def get(blk, key, l1, c1, l2, c2):
        """element of a block, 0 when a lowered exponent would be negative (its coefficient vanishes)"""
        if min(c1) < 0 or min(c2) < 0 or key not in blk:
            return 0.0
        return blk[key][ia[l1][tuple(c1)], ia[l2][tuple(c2)]]

    comps_a = [np.array(c) for c in ia[la]]
    comps_b = [np.array(c) for c in ia[lb]]

    def gradients(s):
        """G_A[j], G_B[j] [ca, cb] from the n = 1 blocks: 2 z g(+1j) - a_j g(-1j)"""
        blk = blocks(s, 1)
        GA = np.zeros((3, len(comps_a), len(comps_b)))
        GB = np.zeros_like(GA)
        for j in range(3):
            for x, ca in enumerate(comps_a):
                for y, cb in enumerate(comps_b):
                    GA[j, x, y] = 2 * get(blk, (1, 0), la + 1, ca + e[j], lb, cb) - ca[j] * get(blk, (-1, 0), la - 1, ca - e[j], lb, cb)
                    GB[j, x, y] = 2 * get(blk, (0, 1), la, ca, lb + 1, cb + e[j]) - cb[j] * get(blk, (0, -1), la, ca, lb - 1, cb - e[j])
        return GA, GB

    blk2 = blocks(base, 2)
    assert set(blk2) >= {(2, 0), (1, 0), (1, 1), (-1, 1), (0, 2), (0, 1)}
    HAA = np.zeros((3, 3, len(comps_a), len(comps_b)))
    HAB = np.zeros_like(HAA)
    for i in range(3):
        for j in range(3):
            dij = int(i == j)
            for x, ca in enumerate(comps_a):
                for y, cb in enumerate(comps_b):
                    HAA[i, j, x, y] = (4 * get(blk2, (2, 0), la + 2, ca + e[i] + e[j], lb, cb)
                                       - 2 * (ca[i] + dij) * get(blk2, (1, 0), la, ca + e[j] - e[i], lb, cb)
                                       - 2 * ca[j] * get(blk2, (1, 0), la, ca - e[j] + e[i], lb, cb)
                                       + ca[j] * (ca[i] - dij) * get(blk2, (-2, 0), la - 2, ca - e[j] - e[i], lb, cb))
                    HAB[i, j, x, y] = (4 * get(blk2, (1, 1), la + 1, ca + e[i], lb + 1, cb + e[j])
                                       - 2 * cb[j] * get(blk2, (1, -1), la + 1, ca + e[i], lb - 1, cb - e[j])
                                       - 2 * ca[i] * get(blk2, (-1, 1), la - 1, ca - e[i], lb + 1, cb + e[j])
                                       + ca[i] * cb[j] * get(blk2, (-1, -1), la - 1, ca - e[i], lb - 1, cb - e[j]))
    # central differences at h and 2h, Richardson-extrapolated (error O(h^4)).  The two sides agree to 1.1e-6 absolute on a
    # scale of 0.11 - and to the same 1.1e-6 when both are assembled from the compiled reference's own blocks on the CPU
    # (the integrals themselves are only that consistent under a displacement), independent of h: tolerance 2e-5 of the
    # scale; a wrong coefficient or a wrong block shows up at order one.
    h = 4e-3

    def central(i, step):
        g = []
        for sgn in (+1, -1):
            s = dict(base)
            geo = base["geometry"].copy()
            geo[3 * A + i] += sgn * step
            s["geometry"] = geo
            g.append(gradients(s))
        return (g[0][0] - g[1][0]) / (2 * step), (g[0][1] - g[1][1]) / (2 * step)

    for i in range(3):
        a1, b1 = central(i, h)
        a2, b2 = central(i, 2 * h)
        fdA, fdB = (4 * a1 - a2) / 3, (4 * b1 - b2) / 3
        scale = max(np.abs(fdA).max(), np.abs(fdB).max(), 1e-6)
        print("n=2 vs FD, direction", i, "max |d| AA", np.abs(HAA[i] - fdA).max(), "AB", np.abs(HAB[i] - fdB).max(), "scale", scale)
        assert np.allclose(HAA[i], fdA, rtol=0, atol=2e-5 * scale), (i, np.abs(HAA[i] - fdA).max(), scale)
        assert np.allclose(HAB[i], fdB, rtol=0, atol=2e-5 * scale), (i, np.abs(HAB[i] - fdB).max(), scale)


@pytest.mark.parametrize("name", ["cfg2", "au4", "cfg4b"])
def test_spherical_output_is_the_transformed_cartesian_matrix(name):
    """scope row f4, "output to spherical AOs": libecp_b200_spherical_host returns S = C^T M C per shell pair with the
    reference's Cartesian -> real-spherical-harmonic table (TM_cart2sph, src/transformations.c:28-87) - checked against
    the same transformation of the reference's Cartesian matrix in numpy; the trace-like invariant: an s shell is
    unchanged up to the constant cart2sph[0][0][0]"""
    s = SMALL[name]()
    ref = load_matrix(name)
    full = ref + np.triu(ref, 1).T  # the Cartesian matrix is symmetric; the reference stores its upper triangle
    with capi.Handle(s) as h:
        c2s = h.host_table("cart2sph")
        rc, S = h.spherical_host()
    assert rc == 0
    ls = np.asarray(s["lBS"])
    nc = [(l + 1) * (l + 2) // 2 for l in ls]
    aoff = np.concatenate([[0], np.cumsum(nc)])
    soff = np.concatenate([[0], np.cumsum(2 * ls + 1)])
    toff = np.concatenate([[0], np.cumsum([(2 * l + 1) * (l + 1) * (l + 2) // 2 for l in range(ls.max() + 1)])])
    C = np.zeros((aoff[-1], soff[-1]))
    for k, l in enumerate(ls):
        blk = c2s[toff[l]:toff[l + 1]].reshape(2 * l + 1, nc[k])  # [m][c]
        C[aoff[k]:aoff[k + 1], soff[k]:soff[k + 1]] = blk.T
    want = np.triu(C.T @ full @ C)
    assert S.shape == want.shape and np.all(np.tril(S, -1) == 0.0)
    assert np.all(np.abs(S - want) <= 1e-12 + 1e-10 * np.abs(want)), np.abs(S - want).max()
    assert np.count_nonzero(S) > 0
