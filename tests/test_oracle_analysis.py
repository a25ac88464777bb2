"""Oracle restatements around the hot path, pinned on the CPU against independent
implementations: OSPA (postanalysis/Plot.cs:531-581) against scipy's assignment solver and hand-worked cases,
measurement generation (SIMV:243-295) against the numpy host model used by the headless runs."""
import numpy as np
import pytest

from monorfs_b200 import synth
from oracle import orc


def test_ospa_hand_cases():
    # identical sets; one missing landmark; everything beyond the cutoff
    a = np.array([[0.0, 0, 0], [1, 0, 0]])
    assert orc.ospa(a, a, 1.0, 2.0) == (0.0, 0.0)
    v, card = orc.ospa(a[:1], a, 1.0, 2.0)
    assert v == pytest.approx(np.sqrt(0.5)) and card == pytest.approx(np.sqrt(0.5))
    v, card = orc.ospa(a, a + 10.0, 1.0, 2.0)
    assert v == pytest.approx(1.0) and card == 0.0
    assert orc.ospa(np.zeros((0, 3)), a, 0.7, 2.0) == (0.7, 0.7)        # Plot.cs:539-542
    assert orc.ospa(np.zeros((0, 3)), np.zeros((0, 3)), 0.7, 2.0) == (0.0, 0.0)
    # a shift of 0.3 on every landmark, order 1
    v, _ = orc.ospa(a + [0.3, 0, 0], a, 1.0, 1.0)
    assert v == pytest.approx(0.3)


@pytest.mark.parametrize("na,nb,c,p", [(5, 8, 1.0, 2.0), (40, 40, 1.0, 2.0), (60, 90, 0.5, 1.0), (120, 121, 2.0, 2.0)])
def test_ospa_against_scipy(na, nb, c, p):
    from scipy.optimize import linear_sum_assignment
    rng = np.random.default_rng(na * 1000 + nb)
    a, b = rng.uniform(-2, 2, (na, 3)), rng.uniform(-2, 2, (nb, 3))
    cost = np.full((nb, nb), c ** p)
    cost[:na] = np.minimum(c, np.linalg.norm(a[:, None] - b[None], axis=2)) ** p
    r, k = linear_sum_assignment(cost)
    ref = (cost[r, k].sum() / nb) ** (1 / p)
    v, card = orc.ospa(a, b, c, p)
    assert abs(v - ref) <= 1e-6          # entries within 1e-5 of C^P are dropped (Plot.cs:559)
    assert card == pytest.approx(c * ((nb - na) / nb) ** (1 / p))
    assert orc.ospa(b, a, c, p) == (v, card)                              # the swap of Plot.cs:533-537


def test_generate_measurements_against_host_model():
    sc = synth.make_scene(2, 300, 16, seed=5)
    ocfg = orc.make_config(sc.params)
    rng = np.random.default_rng(1)
    pose = sc.poses[0]
    n, nc = 600, 9
    lm = sc.map_m[rng.integers(0, len(sc.map_m), n)] + rng.normal(size=(n, 3)) * 0.8
    un, ga, cu = rng.random(n), rng.normal(size=(n, 3)), rng.random((nc, 3))
    R = np.asarray(sc.params["R"], float).reshape(3, 3)
    chol = np.linalg.cholesky(R)
    z, assoc = orc.generate_measurements(ocfg, pose, lm, un, ga, chol, cu)
    meas, ramp = sc.params["measurer"], sc.params["visibility_ramp"]
    zp = synth.measure_perfect(pose, lm, meas[0])
    pdet = sc.params["pd"] * synth.fuzzy_visible(zp, meas, ramp)
    hit = (pdet > 0) & (un < pdet)
    assert 0 < hit.sum() < n
    np.testing.assert_array_equal(assoc[:-nc], np.nonzero(hit)[0])
    np.testing.assert_allclose(z[:-nc], zp[hit] + ga[hit] @ chol.T, rtol=1e-12, atol=1e-12)
    assert (assoc[-nc:] == np.iinfo(np.int32).min).all()
    rmin, length = float(np.float32(meas[1])), float(np.float32(meas[2]) - np.float32(meas[1]))   # AForge.Range: floats
    want = np.stack([cu[:, 0] * meas[5] + meas[3], cu[:, 1] * meas[6] + meas[4], cu[:, 2] * length + rmin], axis=1)
    np.testing.assert_allclose(z[-nc:], want, rtol=1e-14)
