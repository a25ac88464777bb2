"""Boundary entry points added in round 2, through the C ABI against the oracle: the static likelihood functions
of PHDNavigator (SetLikelihood, SetLogLikeMatrix, QuasiSetLogLikelihood: PHD:395-460, 526-713) and the two-step
SlamUpdate that lets the host draw the wheel's uniform only when the reference does (PHD:355-357, 727)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
RTOL = 1e-9


@pytest.fixture(scope="module")
def capi():
    from monorfs_b200 import capi as c
    return c


@pytest.fixture(scope="module")
def orc():
    from oracle import orc as o
    return o


@pytest.fixture(scope="module")
def synth():
    from monorfs_b200 import synth as s
    return s


@pytest.fixture(scope="module")
def ctx(capi, orc, synth):
    sc = synth.make_scene(4, 40, 16, seed=3)
    h = capi.Handle(sc.params, max_particles=4, max_components=256, max_measurements=64, max_pairs=2048)
    h.reset(4, sc.poses[0], sc.map_w, sc.map_m, sc.map_P)
    yield dict(h=h, sc=sc, ocfg=orc.make_config(sc.params))
    h.close()


def scenes(ctx, orc):
    """(pose, landmarks, measurements) cases: the scene's own frame, competing landmarks, dense random blocks."""
    sc, ocfg = ctx["sc"], ctx["ocfg"]
    fr = sc.next_frame()
    yield sc.poses[0], sc.map_m, fr.z
    yield sc.poses[1], sc.map_m[:7], fr.z
    yield sc.poses[0], sc.map_m[:0], fr.z
    pose = [0, 0, 0, 1, 0, 0, 0]
    lm = np.array([[0.2, 0.1, 3.0], [0.21, 0.1, 3.0], [-1.0, 0.5, 5.0]])
    z = np.array([orc.measure_perfect(ocfg, pose, l) for l in lm])
    yield pose, lm, np.concatenate([z + [0.5, -0.5, 0.01], [[100.0, 100.0, 4.0]]])
    for seed in range(4):
        rng = np.random.default_rng(300 + seed)
        J, M = 18, 14
        zc = np.stack([rng.uniform(-60, 60, J), rng.uniform(-40, 40, J), rng.uniform(2.0, 2.8, J)], axis=1)
        lms = np.array([orc.measure_to_map(ocfg, pose, zz) for zz in zc])
        zs = zc[rng.integers(0, J, M)] + rng.normal(size=(M, 3)) * [2.5, 2.5, 0.05]
        yield pose, lms, zs


def test_quasi_set_loglikelihood(ctx, orc):
    """Full visibility, gate d < 12: wider blocks than SetLogLikelihood on the same inputs (PHD:561-713)."""
    h, ocfg = ctx["h"], ctx["ocfg"]
    n = 0
    for pose, lm, z in scenes(ctx, orc):
        exp = orc.quasi_set_loglikelihood(ocfg, pose, lm, z)
        got = h.quasi_set_loglikelihood(pose, lm, z)
        assert np.isfinite(exp)
        assert abs(got - exp) <= RTOL * max(1.0, abs(exp)), (n, got, exp)
        n += 1
    assert n >= 8


def test_set_likelihood_is_exp_of_the_log(ctx, orc):
    h, ocfg = ctx["h"], ctx["ocfg"]
    for pose, lm, z in scenes(ctx, orc):
        exp = np.exp(orc.set_loglikelihood(ocfg, pose, lm, z))
        got = h.set_likelihood(pose, lm, z)
        assert abs(got - exp) <= 1e-9 * abs(exp) + 1e-300, (got, exp)


def test_set_loglike_matrix(ctx, orc):
    """Same entries at the same (row, column) positions as the reference's SparseMatrix (PHD:415-460)."""
    h, ocfg = ctx["h"], ctx["ocfg"]
    for pose, lm, z in scenes(ctx, orc):
        er, ec, ev = orc.set_loglike_matrix(ocfg, pose, lm, z)
        gr, gc, gv = h.set_loglike_matrix(pose, lm, z)
        assert np.array_equal(gr, er) and np.array_equal(gc, ec)
        assert np.allclose(gv, ev, rtol=RTOL, atol=0)
        assert len(gr) >= len(lm) + len(z)   # one miss entry per landmark, one clutter entry per measurement


@pytest.mark.parametrize("meff", [0.95, 0.05])
def test_two_step_slam_update_equals_one_step(capi, synth, meff):
    """_begin reports ParticleDepleted; _finish(u) runs the wheel: identical to rbphd_slam_update(u)."""
    P, N, M = 12, 50, 20
    sc = synth.make_scene(P, N, M, seed=11, min_effective_particle=meff)
    frames = [sc.next_frame() for _ in range(5)]
    a = capi.Handle(sc.params, max_particles=P, max_components=256, max_measurements=M)
    b = capi.Handle(sc.params, max_particles=P, max_components=256, max_measurements=M)
    for h in (a, b):
        h.reset(P, sc.poses[0], sc.map_w, sc.map_m, sc.map_P)
        h.set_poses(sc.poses)
    ndep = 0
    for fr in frames:
        a.update(fr.reading, synth.DT, fr.gauss)
        b.update(fr.reading, synth.DT, fr.gauss)
        best1, res1 = a.slam_update(fr.z, fr.u)
        best2, dep = b.slam_update_begin(fr.z)
        assert dep == res1
        if dep:
            ndep += 1
            best2 = b.slam_update_finish(fr.u)
        else:
            assert b.slam_update_finish(fr.u) == best2     # no-op
        assert best2 == best1
        assert np.allclose(a.get_weights(), b.get_weights(), rtol=1e-12, atol=0)   # (Map.Evaluate sums use atomics)
        assert np.array_equal(a.get_ancestors(), b.get_ancestors())
        assert np.array_equal(a.get_poses(), b.get_poses())
        assert np.array_equal(a.get_map_counts(), b.get_map_counts())
        for i in range(P):
            for x, y in zip(a.get_map(i), b.get_map(i)):
                assert np.array_equal(x, y)
    assert (ndep > 0) == (meff > 0.5)
    a.close()
    b.close()


def test_filter_missing_batch_equals_single_filters(capi, orc, synth):
    """LoopyPHDNavigator.FilterMissing for every held-out index as one batch of "particles" (LOOPY:729-763)."""
    T, N, M = 7, 30, 14
    sc = synth.make_scene(1, N, M, seed=21)
    frames = [sc.next_frame() for _ in range(T)]
    traj = [fr.true_pose for fr in frames]
    factors = [fr.z for fr in frames]
    factors[3] = np.zeros((0, 3))                       # a frame without measurements
    h = capi.Handle(sc.params, max_particles=T + 1, max_components=256, max_measurements=M)
    h.reset(T + 1, traj[0], np.zeros(0), np.zeros((0, 3)), np.zeros((0, 3, 3)))
    h.filter_missing_batch(traj + [traj[-1]], factors + [factors[-1]], to=T)
    ocfg = orc.make_config(sc.params)
    for j in range(T + 1):                              # particle T holds out nothing: the plain filter
        nav = orc.Navigator(ocfg, 1, traj[0], only_mapping=True)
        for i in range(T):
            if i == j:
                continue
            nav.set_pose(0, traj[i])
            nav.slam_update(factors[i], 0.0)
        ow, om, oP = nav.get_map(0)
        gw, gm, gP = h.get_map(j)
        assert len(gw) == len(ow), (j, len(gw), len(ow))
        assert np.allclose(gw, ow, rtol=RTOL, atol=1e-300)
        assert np.allclose(gm, om, rtol=RTOL, atol=1e-12)
        # merged covariances: raw-moment conditioning floor, as in tests/test_gpu_parity.py
        floor = 64 * np.finfo(float).eps * np.sum(om ** 2, axis=1)[:, None, None]
        assert np.all(np.abs(gP - oP) <= RTOL * np.abs(oP) + floor + 1e-15)
        nav.close()
    counts = h.get_map_counts()
    assert len(set(counts.tolist())) > 1                # the held-out frame matters
    h.close()


def test_kinect_depth_visibility(capi, orc, synth):
    """KinectMeasurer.FuzzyVisibleM (KinectMeasurer.cs:151-173): with a depth frame attached, landmarks behind the
    measured surface lose their detection probability -- SLAM frames and the per-stage calls against the oracle."""
    P, N, M = 4, 80, 30
    border = 8
    meas = list(synth.MEASURER)
    meas[3], meas[4], meas[5], meas[6] = -320 + border, -240 + border, 640 - 2 * border, 480 - 2 * border   # deflated film
    sc = synth.make_scene(P, N, M, seed=51, measurer=meas, min_effective_particle=0.5)
    # a wall at 6 m over the left half of the image, open space on the right, a band without readings
    depth = np.full((640, 480), 9.5, dtype=np.float32)
    depth[:300, :] = 6.0
    depth[300:330, :] = np.nan
    xs = np.arange(640)[:, None]
    depth[:300, :] += (0.002 * xs[:300]).astype(np.float32)
    ocfg = orc.make_config(sc.params)
    h = capi.Handle(sc.params, max_particles=P, max_components=512, max_measurements=M)
    nav = orc.Navigator(ocfg, P, sc.poses[0])
    h.reset(P, sc.poses[0], sc.map_w, sc.map_m, sc.map_P)
    h.set_poses(sc.poses)
    for i in range(P):
        nav.set_pose(i, sc.poses[i])
        nav.set_map(i, sc.map_w, sc.map_m, sc.map_P)
    try:
        orc.set_depth_frame(depth)
        h.set_depth_frame(depth)
        fr = sc.next_frame()
        # the stage entry points see the same measurer
        exp = orc.correct(ocfg, sc.poses[0], sc.map_w, sc.map_m, sc.map_P, fr.z)
        got = h.stage_correct(sc.poses[0], sc.map_w, sc.map_m, sc.map_P, fr.z)
        assert len(got[0]) == len(exp[0]) and np.allclose(got[0], exp[0], rtol=RTOL, atol=1e-300)
        plain = None
        for f in range(4):
            h.update(fr.reading, synth.DT, fr.gauss)
            nav.update(fr.reading, synth.DT, fr.gauss)
            gbest, gres = h.slam_update(fr.z, fr.u)
            obest, ores, oanc = nav.slam_update(fr.z, fr.u)
            assert (gbest, gres) == (obest, ores), f
            assert h.get_ancestors().tolist() == oanc.tolist()
            with np.errstate(divide="ignore"):
                assert np.allclose(np.log(h.get_alphas()), np.log(nav.get_alphas()), rtol=RTOL, atol=0)
            for i in range(P):
                gw, gm, gP = h.get_map(i)
                ow, om, oP = nav.get_map(i)
                assert len(gw) == len(ow), (f, i)
                assert np.allclose(gw, ow, rtol=RTOL, atol=1e-300) and np.allclose(gm, om, rtol=RTOL, atol=1e-12)
            if f == 0:
                plain = h.get_map(0)[0].copy()
            fr = sc.next_frame()
        # occlusion changes the result: the same first frame without the depth frame gives other weights
        orc.set_depth_frame(None)
        h.set_depth_frame(None)
        h.reset(P, sc.poses[0], sc.map_w, sc.map_m, sc.map_P)
        h.set_poses(sc.poses)
        sc2 = synth.make_scene(P, N, M, seed=51, measurer=meas, min_effective_particle=0.5)
        fr0 = sc2.next_frame()
        h.update(fr0.reading, synth.DT, fr0.gauss)
        h.slam_update(fr0.z, fr0.u)
        w0 = h.get_map(0)[0]
        assert len(w0) != len(plain) or not np.allclose(w0, plain, rtol=1e-6)
    finally:
        orc.set_depth_frame(None)
        h.close()


@pytest.mark.parametrize("sum_normalised", [False, True])
def test_quasi_set_loglikelihood_gradient(ctx, orc, sum_normalised):
    """QuasiSetLogLikelihood(..., out gradient) (PHD:544-549, 561-713) incl. the in-place TemperedAverage on the
    shared logcomp buffer, in both normalisations (oracle/README.md D10)."""
    h, ocfg = ctx["h"], ctx["ocfg"]
    orc.lib().orc_set_tempered_norm(1 if sum_normalised else 0)
    try:
        n = 0
        for pose, lm, z in scenes(ctx, orc):
            ev, eg = orc.quasi_set_loglikelihood_gradient(ocfg, pose, lm, z)
            gv, gg = h.quasi_set_loglikelihood_gradient(pose, lm, z, sum_normalised=sum_normalised)
            assert abs(gv - ev) <= RTOL * max(1.0, abs(ev)), (n, gv, ev)
            assert np.allclose(gg, eg, rtol=1e-9, atol=1e-9 * max(1.0, float(np.max(np.abs(eg))))), (n, gg, eg)
            n += 1
        assert n >= 8
    finally:
        orc.lib().orc_set_tempered_norm(0)


def test_quasi_gradient_is_the_derivative_in_the_sum_normalised_variant(ctx, orc):
    """With sum normalisation the gradient is the derivative of the value along MeasurementJacobianP's
    parametrisation (translation in the world frame, PRM:200-206): central differences on the translation part."""
    h, ocfg = ctx["h"], ctx["ocfg"]
    pose = np.array([0.02, -0.01, 0.03, 1, 0, 0, 0.0])
    rng = np.random.default_rng(77)
    zc = np.stack([rng.uniform(-60, 60, 6), rng.uniform(-40, 40, 6), rng.uniform(2.0, 2.8, 6)], axis=1)
    lms = np.array([orc.measure_to_map(ocfg, pose, zz) for zz in zc])
    zs = zc + rng.normal(size=(6, 3)) * [1.0, 1.0, 0.02]
    _, g = h.quasi_set_loglikelihood_gradient(pose, lms, zs, sum_normalised=True)
    eps = 1e-6
    for a in range(3):
        pp, pm = pose.copy(), pose.copy()
        pp[a] += eps
        pm[a] -= eps
        num = (h.quasi_set_loglikelihood(pp, lms, zs) - h.quasi_set_loglikelihood(pm, lms, zs)) / (2 * eps)
        assert abs(num - g[a]) <= 1e-4 * max(1.0, abs(g[a])), (a, num, g[a])


def test_two_step_update_must_be_finished(capi, synth):
    """A frame that found the particles depleted has to be finished before the next one starts."""
    P, N, M = 8, 30, 12
    sc = synth.make_scene(P, N, M, seed=12, min_effective_particle=0.99)
    h = capi.Handle(sc.params, max_particles=P, max_components=128, max_measurements=M)
    h.reset(P, sc.poses[0], sc.map_w, sc.map_m, sc.map_P)
    h.set_poses(sc.poses)
    fr = sc.next_frame()
    h.update(fr.reading, synth.DT, fr.gauss)
    _, dep = h.slam_update_begin(fr.z)
    assert dep
    with pytest.raises(capi.RbphdError):
        h.slam_update(fr.z, fr.u)
    h.slam_update_finish(fr.u)
    fr = sc.next_frame()
    h.update(fr.reading, synth.DT, fr.gauss)
    h.slam_update(fr.z, fr.u)
    h.close()
