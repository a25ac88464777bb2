"""Kernels next to the hot path through the C ABI against the oracle:
measurement generation of the simulated vehicle (SIMV:243-295) and the OSPA map metric (postanalysis/Plot.cs:531-581)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
RTOL = 1e-9


@pytest.fixture(scope="module")
def ctx():
    from monorfs_b200 import capi, synth
    from oracle import orc
    sc = synth.make_scene(2, 400, 16, seed=11)
    h = capi.Handle(sc.params, max_particles=2, max_components=512, max_measurements=64, max_pairs=2048)
    h.reset(2, sc.poses[0], sc.map_w, sc.map_m, sc.map_P)
    yield dict(h=h, sc=sc, orc=orc, ocfg=orc.make_config(sc.params))
    h.close()


@pytest.mark.parametrize("n,nc,seed", [(0, 0, 1), (0, 5, 2), (1, 0, 3), (400, 12, 4), (3000, 40, 5)])
def test_generate_measurements(ctx, n, nc, seed):
    """Same detections in the same order, same data association, bit-identical values (no reduction involved)."""
    h, sc, orc, ocfg = ctx["h"], ctx["sc"], ctx["orc"], ctx["ocfg"]
    rng = np.random.default_rng(seed)
    pose = sc.poses[0]
    # landmarks around the scene's own (in view, on the ramp and out of view)
    base = sc.map_m[rng.integers(0, len(sc.map_m), n)] if n else np.zeros((0, 3))
    lm = base + rng.normal(size=(n, 3)) * 0.8
    un, ga, cu = rng.random(n), rng.normal(size=(n, 3)), rng.random((nc, 3))
    chol = np.linalg.cholesky(np.asarray(sc.params["R"], float).reshape(3, 3))
    zo, ao = orc.generate_measurements(ocfg, pose, lm, un, ga, chol, cu)
    zg, ag = h.generate_measurements(pose, lm, un, ga, cu, chol=chol)
    assert len(zg) == len(zo)
    np.testing.assert_array_equal(ag, ao)
    np.testing.assert_array_equal(zg, zo)
    if n >= 400:
        assert 0 < (ao >= 0).sum() < n          # the case exercises both branches
    assert (ao[len(ao) - nc:] == np.iinfo(np.int32).min).all()
    # the library's own root of R (Util.RandomGaussianVector) agrees with numpy's to rounding
    zl, al = h.generate_measurements(pose, lm, un, ga, cu)
    np.testing.assert_array_equal(al, ao)
    np.testing.assert_allclose(zl, zo, rtol=1e-12, atol=1e-12)


def test_generate_measurements_feeds_the_filter(ctx):
    """The generated set is a valid input of the frame path (shape, range clip)."""
    h, sc = ctx["h"], ctx["sc"]
    rng = np.random.default_rng(9)
    n = len(sc.map_m)
    z, assoc = h.generate_measurements(sc.poses[0], sc.map_m, rng.random(n), rng.normal(size=(n, 3)), rng.random((4, 3)))
    meas = sc.params["measurer"]
    assert z.shape[1] == 3 and len(z) == len(assoc)
    cl = z[assoc < 0]
    assert ((cl[:, 2] >= np.float32(meas[1])) & (cl[:, 2] <= np.float32(meas[2]))).all()


@pytest.mark.parametrize("na,nb,seed", [(0, 0, 1), (0, 7, 2), (1, 1, 3), (5, 8, 4), (8, 5, 5), (60, 60, 6),
                                        (300, 340, 7), (500, 540, 8)])
@pytest.mark.parametrize("c,p", [(1.0, 2.0), (0.5, 1.0)])
def test_ospa(ctx, na, nb, seed, c, p):
    h, orc = ctx["h"], ctx["orc"]
    rng = np.random.default_rng(seed)
    b = rng.uniform(-4, 4, (nb, 3))
    # a: noisy copies of a subset of b (distances below and above the cutoff) plus strays
    a = (b[rng.permutation(nb)[:na]] if nb else np.zeros((0, 3)))
    if len(a) < na:
        a = np.concatenate([a, rng.uniform(-4, 4, (na - len(a), 3))])
    a = a + rng.normal(size=(na, 3)) * rng.choice([0.02, 0.3, 2.0], size=(na, 1))
    want = orc.ospa(a, b, c, p)
    got = h.ospa(a, b, c, p)
    np.testing.assert_allclose(got, want, rtol=RTOL, atol=1e-12)


def test_ospa_against_scipy(ctx):
    """Independent check of both: the metric is the optimum of a dense assignment problem."""
    from scipy.optimize import linear_sum_assignment
    h, orc = ctx["h"], ctx["orc"]
    rng = np.random.default_rng(21)
    for na, nb in [(12, 20), (150, 150), (500, 520)]:
        a, b = rng.uniform(-2, 2, (na, 3)), rng.uniform(-2, 2, (nb, 3))
        C, P = 1.0, 2.0
        cost = np.full((nb, nb), C ** P)
        cost[:na] = np.minimum(C, np.linalg.norm(a[:, None] - b[None], axis=2)) ** P
        r, k = linear_sum_assignment(cost)
        ref = (cost[r, k].sum() / nb) ** (1 / P)
        assert abs(h.ospa(a, b, C, P)[0] - ref) <= 1e-6      # (entries within 1e-5 of C^P are dropped, Plot.cs:559)
        assert abs(orc.ospa(a, b, C, P)[0] - ref) <= 1e-6


def test_ospa_of_the_filter_estimate(ctx):
    """End of a SLAM frame: OSPA between the best particle's map estimate and the scene's landmarks."""
    h, sc, orc = ctx["h"], ctx["sc"], ctx["orc"]
    w, m, _ = h.get_map(0)
    est = m[w > 0.5]
    got = h.ospa(est, sc.map_m, 1.0, 2.0)
    want = orc.ospa(est, sc.map_m, 1.0, 2.0)
    np.testing.assert_allclose(got, want, rtol=RTOL, atol=1e-12)


def test_headless_run_measured_and_scored_on_device(ctx):
    """A small simulated run whose measurements come from rbphd_generate_measurements and whose map error comes from
    rbphd_ospa: the error matches the oracle's on the same sets and ends below the empty-map error it starts from."""
    from monorfs_b200 import simulation, synth
    orc = ctx["orc"]
    pose0, measurer, lm = simulation.synthetic_scene(12, seed=4)
    run = simulation.HeadlessRun(pose0, measurer, lm, simulation.synthetic_commands(25), particles=8, seed=4,
                                 device_measure=True)
    try:
        rec = run.run()
        err = simulation.map_error(rec, run.h, 1.0, 2.0)
        visited = simulation.visited_map(rec)
        assert len(visited) > 0 and len(err) == 25
        for (t, o, _), (_, (w, m, _)) in zip(err, rec.maps):
            want = orc.ospa(visited, simulation.best_map_estimate(w, m), 1.0, 2.0)[0]
            assert o == pytest.approx(want, rel=RTOL, abs=1e-12)
        assert err[-1][1] < 1.0
    finally:
        run.close()
