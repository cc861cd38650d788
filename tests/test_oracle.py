"""CPU tier: the oracle restatement (oracle/oracle_ecp.c) against the golden fixtures generated from the
unmodified reference, against the reference itself when oracle/_ref is present, and against the
known-answer vectors of SURVEY.md App. D."""
import os

import numpy as np
import pytest
from conftest import GOLDEN, load_blocks, load_matrix

from libecp_b200 import synth
from oracle.refbind import RefLib, have, port_counters

CASES = {"cfg1": synth.cfg1, "cfg2": synth.cfg2, "au2": lambda: synth.cfg3(2), "cfg4a": lambda: synth.cfg4("a")}


@pytest.fixture(scope="module")
def port():
    return RefLib("port")


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "au2", "cfg4a"])
def test_port_matches_golden_matrix_bitwise(port, name):
    got = port.get_integrals(CASES[name]())
    ref = load_matrix(name)
    assert np.array_equal(got, ref)  # byte/bit-exact: same arithmetic, same libm
    assert np.all(np.tril(got, -1) == 0.0)


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "au2"])
def test_port_matches_golden_blocks_bitwise(port, name):
    keys, off, vals = load_blocks(name)
    rc, recs = port.callbacks(CASES[name]())
    assert rc == 0 and len(recs) == len(keys)
    for k, r in enumerate(recs):
        assert tuple(keys[k]) == r[:9]
        assert np.array_equal(r[9], vals[off[k]:off[k + 1]])


def test_known_answers_config1(port):
    """config-1 matrix to 17 digits and its closed-form (mpmath) values."""
    I = port.get_integrals(synth.cfg1())
    assert I[0, 0] == 7.62425017208175460e-01
    assert I[0, 1] == 2.40834719211528192e+00
    assert I[1, 1] == 5.20958325982253072e+01
    assert I[2, 2] == I[3, 3] == I[4, 4] == 1.78038434232556781e+01
    exact = {(0, 0): 0.762425017208176807, (0, 1): 2.40834719211504847, (1, 1): 52.095832598225433,
             (2, 2): 17.8038434232574127}
    for (i, j), v in exact.items():
        assert abs(I[i, j] - v) <= 2e-13 * abs(v) + 2e-12
    off = I.copy()
    for (i, j) in [(0, 0), (0, 1), (1, 1), (2, 2), (3, 3), (4, 4)]:
        off[i, j] = 0.0
    assert np.all(off == 0.0)


def test_config2_L5_equals_L4(port):
    """The extra l=4 projector vanishes on a one-centre s-f basis."""
    assert np.array_equal(port.get_integrals(synth.cfg2(5)), port.get_integrals(synth.cfg2(4)))


def test_additive_over_centres(port):
    """Masking centres and accumulating the runs in centre order reproduces the full matrix bit for bit."""
    s = synth.probe(3, 2, 3)
    full = port.get_integrals(s)
    acc = np.zeros_like(full)
    for c in range(3):
        acc += port.get_integrals(synth.mask_centres(s, [c]))
    assert np.allclose(acc, full, rtol=0, atol=1e-12 * np.abs(full).max())


def test_structural_invariants(port):
    s = synth.cfg3(2)
    rc, recs = port.callbacks(s)
    for (A, s1, la, _, B, s2, lb, _, C, blk) in recs:
        if A == B == C and (la + lb) % 2 == 1:
            assert np.all(blk == 0.0)
        if A == B and s1 == s2:
            n = (la + 1) * (la + 2) // 2
            m = blk.reshape(n, n)
            assert np.allclose(m, m.T, rtol=1e-12, atol=1e-14)


def test_stale_buffer_quirk_inactive():
    """src/type2.c:443-448 reuses the fallback buffers across primitive pairs; the restatement can run with
    or without that behaviour - on the parity shapes the centre point is never beyond the cut."""
    s = synth.cfg3(2)
    rc1, M1, c1 = port_counters(s, stale=1)
    rc0, M0, c0 = port_counters(s, stale=0)
    assert rc1 == rc0 == 0
    assert c1["stale_center_hits"] == 0
    assert np.array_equal(M1, M0)
    assert c1["psm92_fail"] == 0 and c1["triples_exec"] == 600


@pytest.mark.skipif(not have("ref"), reason="oracle/_ref (compiled reference) not present on this box")
@pytest.mark.parametrize("name", ["cfg1", "cfg2", "au2", "cfg4a"])
def test_port_equals_compiled_reference(port, name):
    ref = RefLib("ref")
    s = CASES[name]()
    rc1, a = ref.callbacks(s)
    rc2, b = port.callbacks(s)
    assert rc1 == rc2 == 0 and len(a) == len(b)
    for x, y in zip(a, b):
        assert x[:9] == y[:9]
        assert np.array_equal(x[9], y[9])


def test_first_derivative_fixtures_are_the_compiled_references_output():
    """tests/golden/deriv1_*: callback blocks of derivative order n = 1 (the next scope row; the product
    rejects n > 0 today).  Where the compiled reference is available they must be its output bit for bit; everywhere the
    fixtures must be well formed: finite, 4 shift patterns, shifted block sizes."""
    from libecp_b200 import synth

    for name, lbs, L in (("deriv1_tz2_L4", 2, 4), ("deriv1_tz3_L5", 3, 5)):
        d = np.load(os.path.join(GOLDEN, name + "_blocks.npz"))
        keys, off, vals = d["keys"], d["off"], d["vals"]
        assert np.isfinite(vals).all() and len(off) == len(keys) + 1 and off[-1] == len(vals)
        shifts = {(int(k[3]), int(k[7])) for k in keys}
        assert shifts == {(1, 0), (-1, 0), (0, 1), (0, -1)}
        for k, a, b in zip(keys, off[:-1], off[1:]):
            la, lb = k[2] + k[3], k[6] + k[7]
            assert b - a == ((la + 1) * (la + 2) // 2) * ((lb + 1) * (lb + 2) // 2)
        if have("ref"):
            rc, recs = RefLib("ref").callbacks(synth.deriv_pair(lbs, L), n=1)
            assert rc == 0 and len(recs) == len(keys)
            assert np.array_equal(np.array([r[:9] for r in recs], np.int32), keys)
            assert np.array_equal(np.concatenate([r[9] for r in recs]), vals)
