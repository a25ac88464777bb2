"""Parity of the CUDA path (through the C ABI of librbphd.so) against the CPU oracle.

Bar: component counts, ancestors, best index and resampling decisions bit-exact;
means / covariances / weights within 1e-9 relative (relative to the vector / matrix norm).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-9


@pytest.fixture(scope="module")
def capi():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from monorfs_b200 import capi as _capi
    _capi.load()
    return _capi


@pytest.fixture(scope="module")
def synth():
    from monorfs_b200 import synth as _s
    return _s


def close_rel(a, b, rtol=RTOL):
    a, b = np.asarray(a, float), np.asarray(b, float)
    if a.shape != b.shape:
        return False
    scale = max(np.max(np.abs(b)) if b.size else 0.0, 1e-300)
    both_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    diff = np.where(both_inf, 0.0, np.abs(a - b))
    return bool(np.all(diff <= rtol * scale))


def assert_maps_equal(got, exp, what="", merge_floor=0.0):
    """Counts exact, values to RTOL.  merge_floor > 0 (long runs only) adds merge_floor * eps * |m|^2 of absolute
    slack on covariance entries: the reference merges with raw moments, sum w (P + m m^T) / w - m m^T
    (GAUSS:329-344), which cancels |m|^2 / |P| ~ 1e7 leading digits, so one ulp of difference in a weight (CUDA's
    exp/log vs glibc's) moves a merged covariance by ~1e-9 relative and the difference feeds the next frames."""
    gw, gm, gP = got
    ew, em, eP = exp
    assert len(gw) == len(ew), f"{what}: component count {len(gw)} != {len(ew)}"
    assert close_rel(gw, ew), f"{what}: weights differ, max abs {np.max(np.abs(gw - ew)):.3e}"
    eps = np.finfo(float).eps
    for i in range(len(ew)):
        assert close_rel(gm[i], em[i]), f"{what}: mean {i}: {gm[i]} vs {em[i]}"
        if merge_floor:
            tol = RTOL * np.max(np.abs(eP[i])) + merge_floor * eps * float(np.dot(em[i], em[i]))
            assert np.all(np.abs(gP[i] - eP[i]) <= tol), f"{what}: cov {i}: {np.max(np.abs(gP[i] - eP[i])):.3e} > {tol:.3e}"
        else:
            assert close_rel(gP[i], eP[i]), f"{what}: cov {i}"


def small_scene(synth, P=8, N=40, M=16, seed=3, **over):
    return synth.make_scene(P, N, M, seed=seed, **over)


@pytest.fixture(scope="module")
def ctx(capi, orc, synth):
    sc = small_scene(synth)
    h = capi.Handle(sc.params, max_particles=64, max_components=256, max_measurements=64, max_pairs=4096)
    yield dict(h=h, sc=sc, ocfg=orc.make_config(sc.params))
    h.close()


# ------------------------------------------------------------------ per-particle stages
def test_stage_predict(ctx, orc):
    h, sc, ocfg = ctx["h"], ctx["sc"], ctx["ocfg"]
    fr = sc.next_frame()
    pose = sc.poses[1]
    exp = orc.predict(ocfg, pose, sc.map_w, sc.map_m, sc.map_P, fr.z)
    got = h.stage_predict(pose, sc.map_w, sc.map_m, sc.map_P, fr.z)
    assert exp[3] > 0, "scene must produce births"
    assert_maps_equal(got, exp[:3], "predict")


def test_stage_predict_empty_map(ctx, orc):
    """PHDNavigatorTest.PredictInitial (PHDNavigatorTest.cs:85-104) on the pixel-range model."""
    h, sc, ocfg = ctx["h"], ctx["sc"], ctx["ocfg"]
    pose = [0.1, -0.2, 0.05, 1, 0, 0, 0]
    z = [[10.0, -20.0, 2.5]]
    got = h.stage_predict(pose, np.zeros(0), np.zeros((0, 3)), np.zeros((0, 3, 3)), z)
    exp = orc.predict(ocfg, pose, np.zeros(0), np.zeros((0, 3)), np.zeros((0, 3, 3)), z)
    assert len(got[0]) == 1
    assert_maps_equal(got, exp[:3], "predict-empty")
    assert close_rel(got[1][0], orc.measure_to_map(ocfg, pose, z[0]))
    assert np.allclose(got[2][0], np.eye(3) * 1e-2) and got[0][0] == 0.05


def test_stage_predict_known(ctx, orc):
    """PHDNavigatorTest.PredictKnown (PHDNavigatorTest.cs:106-126): an explored point gives no birth."""
    h, ocfg = ctx["h"], ctx["ocfg"]
    pose = [0, 0, 0, 1, 0, 0, 0]
    z = [[10.0, -20.0, 2.5]]
    c = orc.measure_to_map(ocfg, pose, z[0])
    got = h.stage_predict(pose, [1.0], [c], [np.eye(3)], z)
    assert len(got[0]) == 1 and got[0][0] == 1.0


def test_stage_correct_gated_and_ungated(ctx, orc):
    h, sc, ocfg = ctx["h"], ctx["sc"], ctx["ocfg"]
    fr = sc.next_frame()
    pose = sc.poses[2]
    pw, pm, pP, _ = orc.predict(ocfg, pose, sc.map_w, sc.map_m, sc.map_P, fr.z)
    exp = orc.correct(ocfg, pose, pw, pm, pP, fr.z)
    got = h.stage_correct(pose, pw, pm, pP, fr.z)
    assert len(exp[0]) > len(pw), "scene must produce detections"
    assert_maps_equal(got, exp, "correct-gated")
    # ungated form = what PHDNavigatorTest.Correct pins (PHDNavigatorTest.cs:128-193); small sizes
    exp = orc.correct(ocfg, pose, pw[:6], pm[:6], pP[:6], fr.z[:5], gate_radius=-1.0)
    got = h.stage_correct(pose, pw[:6], pm[:6], pP[:6], fr.z[:5], gate_radius=-1.0)
    assert len(got[0]) == 6 + 6 * 5
    assert_maps_equal(got, exp, "correct-ungated")


def test_stage_prune_reference_case(ctx, orc):
    """PHDNavigatorTest.Prune (PHDNavigatorTest.cs:195-265): 13 components -> 5."""
    h, ocfg = ctx["h"], ctx["ocfg"]
    I = np.eye(3)
    mw, md = 1e-3, 0.3
    comps = [([-12, -24, -54], I, 23.0), ([-80, -22, -12], 4.0 * I, 1.0), ([-63, -11, -95], 0.1 * I, 6.0),
             ([12, 24, 54], I, 0.3 * mw), ([80, 22, 12], 4.0 * I, 0.8 * mw), ([63, 11, 95], 0.1 * I, 0.99 * mw),
             ([23, 19, 73], I, 0.0),
             ([0, 0, 0], I, 1.0), ([0, md, 0], I, 0.6), ([0, md / 2, 0], I, 1.2),
             ([99 - md / 6, 99, 99], I, 0.9), ([99, 99 - md / 6, 99], I, 0.5), ([99, 99, 99 - md / 6], I, 1.1)]
    w = [c[2] for c in comps]
    m = [c[0] for c in comps]
    P = [c[1] for c in comps]
    exp = orc.prune(ocfg, w, m, P)
    got = h.stage_prune(w, m, P)
    assert len(got[0]) == 5
    assert_maps_equal(got, exp, "prune-reference")


def test_stage_prune_scene(ctx, orc):
    h, sc, ocfg = ctx["h"], ctx["sc"], ctx["ocfg"]
    fr = sc.next_frame()
    pose = sc.poses[3]
    pw, pm, pP, _ = orc.predict(ocfg, pose, sc.map_w, sc.map_m, sc.map_P, fr.z)
    cw, cm, cP = orc.correct(ocfg, pose, pw, pm, pP, fr.z)
    exp = orc.prune(ocfg, cw, cm, cP)
    got = h.stage_prune(cw, cm, cP)
    assert_maps_equal(got, exp, "prune-scene")
    # idempotence on a pruned map whose components are not mutually close
    again = h.stage_prune(*got)
    exp2 = orc.prune(ocfg, *exp)
    assert_maps_equal(again, exp2, "prune-twice")


def test_stage_prune_ties_and_chains(ctx, orc):
    """Equal weights (stable order) and a chain a~b~c where b is absorbed by a and c survives."""
    h, ocfg = ctx["h"], ctx["ocfg"]
    I = np.eye(3) * 1e-2
    t = 0.3 * 0.1   # merge radius for cov 1e-2 I
    m = [[0, 0, 0], [0.9 * t, 0, 0], [1.8 * t, 0, 0], [5, 5, 5], [5, 5, 5 + 0.5 * t], [9, 9, 9]]
    w = [0.5, 0.5, 0.5, 0.2, 0.2, 0.2]
    P = [I] * 6
    exp = orc.prune(ocfg, w, m, P)
    got = h.stage_prune(w, m, P)
    assert_maps_equal(got, exp, "prune-ties")


def test_stage_set_loglikelihood(ctx, orc):
    h, sc, ocfg = ctx["h"], ctx["sc"], ctx["ocfg"]
    fr = sc.next_frame()
    pose = sc.poses[0]
    for J in (0, 1, 7, len(sc.map_m)):
        jm = sc.map_m[:J]
        exp = orc.set_loglikelihood(ocfg, pose, jm, fr.z)
        got = h.stage_set_loglikelihood(pose, jm, fr.z)
        assert close_rel(got, exp), (J, got, exp)


def test_stage_set_loglikelihood_shared_measurements(ctx, orc):
    """Two landmarks competing for the same measurements: 4x4 blocks through the lexicographic lane."""
    h, ocfg = ctx["h"], ctx["ocfg"]
    pose = [0, 0, 0, 1, 0, 0, 0]
    lm = np.array([[0.2, 0.1, 3.0], [0.21, 0.1, 3.0], [-1.0, 0.5, 5.0]])
    z = np.array([orc.measure_perfect(ocfg, pose, l) for l in lm])
    z = np.concatenate([z + [0.5, -0.5, 0.01], [[100.0, 100.0, 4.0]]])
    exp = orc.set_loglikelihood(ocfg, pose, lm, z)
    got = h.stage_set_loglikelihood(pose, lm, z)
    assert close_rel(got, exp), (got, exp)


def _cluster(orc, ocfg, pose, centre, n_lm, n_z, rng, spread_px=2.5, spread_r=0.04):
    """n_lm landmarks whose predicted measurements fall within a few pixels / centimetres of each other,
    and n_z measurements around them: one association block of n_lm + n_z rows (PHD:435-436 gate d < 5)."""
    z0 = orc.measure_perfect(ocfg, pose, centre)
    lms, zs = [], []
    for _ in range(n_lm):
        zz = z0 + rng.normal(size=3) * [spread_px, spread_px, spread_r]
        lms.append(orc.measure_to_map(ocfg, pose, zz))
    for _ in range(n_z):
        zs.append(z0 + rng.normal(size=3) * [spread_px, spread_px, spread_r])
    return lms, zs


def test_stage_set_loglikelihood_murty_lane(ctx, orc):
    """Blocks with more than five rows: Hungarian + Murty (GC:64-272) and the stale-buffer early exit
    (PHD:503) across a mix of 2x2, 3-5 row and large blocks."""
    h, ocfg = ctx["h"], ctx["ocfg"]
    pose = [0.05, -0.02, 0.01, 0.999, 0.01, -0.02, 0.015]
    pose = list(pose[:3]) + list(np.array(pose[3:]) / np.linalg.norm(pose[3:]))
    checked = 0
    for seed in range(8):
        rng = np.random.default_rng(100 + seed)
        lms, zs = [], []
        layout = [((0.3, 0.2, 3.0), 1, 1), ((-0.8, 0.4, 4.0), 2, 1), ((1.2, -0.6, 6.0), 3 + seed % 3, 3 + seed % 2),
                  ((-1.5, -0.9, 7.5), 1, 1), ((0.1, 0.9, 5.0), 4, 5), ((2.0, 1.0, 8.0), 2, 2)]
        rng.shuffle(layout)
        for centre, nl, nz in layout:
            a, b = _cluster(orc, ocfg, pose, np.array(centre), nl, nz, rng)
            lms += a
            zs += b
        zs.append([250.0, -200.0, 3.3])   # isolated clutter
        lms.append([5.0, 5.0, -3.0])      # landmark behind the camera
        lms, zs = np.array(lms), np.array(zs)
        exp = orc.set_loglikelihood(ocfg, pose, lms, zs)
        got = h.stage_set_loglikelihood(pose, lms, zs)
        assert np.isfinite(exp)
        assert close_rel(got, exp), (seed, got, exp)
        checked += 1
    assert checked == 8


def test_stage_set_loglikelihood_dense_random(ctx, orc):
    h, ocfg = ctx["h"], ctx["ocfg"]
    pose = [0, 0, 0, 1, 0, 0, 0]
    for seed in range(6):
        rng = np.random.default_rng(200 + seed)
        J, M = 24, 20
        zc = np.stack([rng.uniform(-40, 40, J), rng.uniform(-30, 30, J), rng.uniform(2.0, 2.6, J)], axis=1)
        lms = np.array([orc.measure_to_map(ocfg, pose, z) for z in zc])
        zs = zc[rng.integers(0, J, M)] + rng.normal(size=(M, 3)) * [1.5, 1.5, 0.03]
        exp = orc.set_loglikelihood(ocfg, pose, lms, zs)
        got = h.stage_set_loglikelihood(pose, lms, zs)
        assert close_rel(got, exp), (seed, got, exp)


def test_stage_weight_alpha(ctx, orc):
    h, sc, ocfg = ctx["h"], ctx["sc"], ctx["ocfg"]
    fr = sc.next_frame()
    for pi in (0, 5):
        pose = sc.poses[pi]
        pw, pm, pP, _ = orc.predict(ocfg, pose, sc.map_w, sc.map_m, sc.map_P, fr.z)
        cw, cm, cP = orc.correct(ocfg, pose, pw, pm, pP, fr.z)
        qw, qm, qP = orc.prune(ocfg, cw, cm, cP)
        exp = orc.weight_alpha(ocfg, pose, fr.z, (pw, pm, pP), (qw, qm, qP))
        got = h.stage_weight_alpha(pose, fr.z, (pw, pm, pP), (qw, qm, qP))
        assert got["J"] == exp["J"]
        for k in ("pcount", "ccount", "ploglik", "cloglik", "setloglik"):
            assert close_rel(got[k], exp[k]), (k, got[k], exp[k])
        assert close_rel(np.log(got["alpha"]), np.log(exp["alpha"]), 1e-9), (got["alpha"], exp["alpha"])


def test_best_map_estimate_heavy_weights(ctx, orc):
    """Weights above 1 are picked repeatedly (MAP:131-138)."""
    h, sc, ocfg = ctx["h"], ctx["sc"], ctx["ocfg"]
    fr = sc.next_frame()
    pose = sc.poses[0]
    n = 12
    w = np.array([2.5, 0.4, 1.5, 1.5, 0.9, 3.2, 0.05, 1.0, 0.3, 0.7, 2.5, 0.6])
    pred = (w, sc.map_m[:n], sc.map_P[:n])
    exp = orc.weight_alpha(ocfg, pose, fr.z, pred, pred)
    got = h.stage_weight_alpha(pose, fr.z, pred, pred)
    assert got["J"] == exp["J"] == int(w.sum())
    for k in ("ploglik", "cloglik", "setloglik"):
        assert close_rel(got[k], exp[k]), (k, got[k], exp[k])


# ------------------------------------------------------------------ particle set
def test_pose_update(ctx, orc):
    h, sc = ctx["h"], ctx["sc"]
    P = sc.P
    h.reset(P, sc.poses[0], sc.map_w, sc.map_m, sc.map_P)
    h.set_poses(sc.poses)
    nav = orc.Navigator(ctx["ocfg"], P, sc.poses[0])
    for i in range(P):
        nav.set_pose(i, sc.poses[i])
    rng = np.random.default_rng(5)
    for _ in range(3):
        g = rng.normal(size=(P, 6))
        reading = [0.01, -0.02, 0.03, 0.004, -0.002, 0.001]
        h.update(reading, 1 / 30.0, g)
        nav.update(reading, 1 / 30.0, g)
        assert np.allclose(h.get_poses(), nav.get_poses(), rtol=0, atol=1e-13)
    # PerfectStill with a zero reading: no noise (TRK:94)
    before = h.get_poses()
    h.update([0] * 6, 1 / 30.0, rng.normal(size=(P, 6)), perfect_still=True)
    nav.update([0] * 6, 1 / 30.0, rng.normal(size=(P, 6)), perfect_still=True)
    assert np.allclose(h.get_poses(), nav.get_poses(), rtol=0, atol=1e-13)
    assert np.allclose(h.get_poses(), before, rtol=0, atol=1e-12)


def test_resample_reference_invariants(ctx, orc):
    """SimulationTest.resample (SimulationTest.cs:225-270) through rbphd_resample."""
    h, sc, ocfg = ctx["h"], ctx["sc"], ctx["ocfg"]
    rng = np.random.default_rng(11)
    seen0 = seen3 = 0
    iters = 60
    for it in range(iters):
        h.reset(5, sc.poses[0], sc.map_w[:3], sc.map_m[:3], sc.map_P[:3])
        poses = np.tile(np.array([0, 0, 0, 1, 0, 0, 0], float), (5, 1))
        poses[:, 0] = np.arange(5)
        h.set_poses(poses)
        for i in range(5):
            h.set_map(i, sc.map_w[:i + 1], sc.map_m[:i + 1], sc.map_P[:i + 1])
        wts = [0.11, 0.28, 0.31, 0.01, 0.29]
        h.set_weights(wts)
        u = float(np.float32(rng.random()))
        h.resample(u)
        anc = h.get_ancestors()
        _, ebest, eanc, _ = orc.normalize_resample(ocfg, wts, u, force=2)
        assert anc.tolist() == eanc.tolist()
        assert h.get_best() == ebest and anc[h.get_best()] == 2
        assert {1, 2, 4} <= set(anc.tolist())
        assert np.allclose(h.get_weights(), 0.2)
        assert h.get_poses()[:, 0].tolist() == anc.astype(float).tolist()
        assert h.get_map_counts().tolist() == (anc + 1).tolist()
        seen0 += 0 in anc
        seen3 += 3 in anc
    assert seen0 < iters and seen3 < iters


def run_both(capi, orc, synth, P, N, M, frames, seed, only_mapping=False, merge_floor=0.0, max_pairs=0,
             final_counts=None, **over):
    sc = synth.make_scene(P, N, M, seed=seed, **over)
    ocfg = orc.make_config(sc.params)
    h = capi.Handle(sc.params, max_particles=P, max_components=max(2 * N, 64), max_measurements=M,
                    max_pairs=max_pairs or max(8 * M, 256))
    nav = orc.Navigator(ocfg, P, sc.poses[0], only_mapping=only_mapping)
    h.reset(P, sc.poses[0], sc.map_w, sc.map_m, sc.map_P)
    h.set_poses(sc.poses)
    for i in range(P):
        nav.set_pose(i, sc.poses[i])
        nav.set_map(i, sc.map_w, sc.map_m, sc.map_P)
    nres = 0
    for f in range(frames):
        fr = sc.next_frame()
        if only_mapping:
            for i in range(P):
                h.set_pose(i, fr.true_pose)
                nav.set_pose(i, fr.true_pose)
        else:
            h.update(fr.reading, synth.DT, fr.gauss)
            nav.update(fr.reading, synth.DT, fr.gauss)
        gbest, gres = h.slam_update(fr.z, fr.u, only_mapping=only_mapping)
        obest, ores, oanc = nav.slam_update(fr.z, fr.u)
        tag = f"frame {f}"
        assert gres == ores, tag
        if not only_mapping:
            with np.errstate(divide="ignore"):
                assert close_rel(np.log(h.get_alphas()), np.log(nav.get_alphas()), 1e-9), tag
            assert h.get_ancestors().tolist() == oanc.tolist(), tag
            assert gbest == obest, tag
            assert close_rel(h.get_weights(), nav.get_weights()), tag
        assert np.allclose(h.get_poses(), nav.get_poses(), rtol=0, atol=1e-12), tag
        counts = h.get_map_counts()
        for i in range(P):
            assert_maps_equal(h.get_map(i), nav.get_map(i), f"{tag} particle {i}", merge_floor)
            assert counts[i] == len(nav.get_map(i)[0])
        nres += int(gres)
    if final_counts is not None:
        final_counts.extend(h.get_map_counts().tolist())
    h.close()
    return nres


def test_slam_frames_small(capi, orc, synth):
    nres = run_both(capi, orc, synth, P=12, N=40, M=16, frames=8, seed=21, min_effective_particle=0.6)
    assert nres >= 1, "the run must exercise resampling"


def test_slam_frames_medium(capi, orc, synth):
    run_both(capi, orc, synth, P=24, N=150, M=48, frames=5, seed=22, min_effective_particle=0.3)


def test_slam_long_run(capi, orc, synth):
    """Thirty consecutive SLAM frames against the oracle (maps grow to MaxQuantity, resampling fires several
    times): drift between the two implementations would show up as a count or ancestor mismatch."""
    nres = run_both(capi, orc, synth, P=16, N=120, M=40, frames=30, seed=27, min_effective_particle=0.5,
                    merge_floor=64.0)
    assert nres >= 3


@pytest.mark.parametrize("nequal", [900, 300, 60])
def test_slam_degenerate_weights(capi, orc, synth, nequal):
    """Hundreds of exactly equal weights and a very low MinWeight: the prune candidate sort must order ties by
    list position like the stable reference sort.  900 equal keys take the radix-sort fallback of the bucket
    sort, 300 the CTA-wide bucket finish, 60 the warp finish."""
    P, N, M = 3, 30, 64
    sc = synth.make_scene(P, N, M, seed=29, min_weight=1e-9, birth_weight=0.05)
    ocfg = orc.make_config(sc.params)
    h = capi.Handle(sc.params, max_particles=P, max_components=2048, max_measurements=M, max_pairs=64 * M)
    nav = orc.Navigator(ocfg, P, sc.poses[0])
    rng = np.random.default_rng(5)
    w = np.concatenate([np.full(nequal, 0.25), rng.uniform(0.3, 0.9, 900 - nequal)])[rng.permutation(900)]
    m = rng.random((900, 3)) * np.array([4.0, 3.0, 6.0]) + np.array([-2.0, -1.5, 1.0])
    Pm = np.tile(np.eye(3) * 1e-3, (900, 1, 1))
    h.reset(P, sc.poses[0], w, m, Pm)
    for i in range(P):
        nav.set_pose(i, sc.poses[0])
        nav.set_map(i, w, m, Pm)
    for f in range(3):
        fr = sc.next_frame()
        gbest, gres = h.slam_update(fr.z, fr.u)
        obest, ores, oanc = nav.slam_update(fr.z, fr.u)
        assert (gbest, gres) == (obest, ores)
        counts = h.get_map_counts()
        for i in range(P):
            assert counts[i] == len(nav.get_map(i)[0]), f"frame {f} particle {i}"
            assert_maps_equal(h.get_map(i), nav.get_map(i), f"degenerate frame {f} particle {i}")
    h.close()


def test_mapping_only_frames(capi, orc, synth):
    run_both(capi, orc, synth, P=3, N=80, M=30, frames=6, seed=23, only_mapping=True)


def test_squared_gate_metric(capi, orc, synth):
    """The Accord squared-Euclidean KD-tree hypothesis as a switch."""
    run_both(capi, orc, synth, P=6, N=50, M=20, frames=3, seed=24, gate_metric=1)


def test_empty_inputs(capi, orc, synth):
    sc = small_scene(synth)
    h = capi.Handle(sc.params, max_particles=4, max_components=128, max_measurements=32)
    h.reset(4, sc.poses[0], np.zeros(0), np.zeros((0, 3)), np.zeros((0, 3, 3)))
    best, res = h.slam_update(np.zeros((0, 3)), 0.5)          # empty map, no measurements
    assert h.get_map_counts().tolist() == [0, 0, 0, 0]
    fr = sc.next_frame()
    h.slam_update(fr.z, 0.5)                                   # empty map, births only
    nav = orc.Navigator(orc.make_config(sc.params), 4, sc.poses[0])
    nav.slam_update(np.zeros((0, 3)), 0.5)
    nav.slam_update(fr.z, 0.5)
    for i in range(4):
        assert_maps_equal(h.get_map(i), nav.get_map(i), f"empty particle {i}")
    h.close()


def test_capacity_error_is_loud(capi, synth):
    sc = small_scene(synth, N=40)
    h = capi.Handle(sc.params, max_particles=2, max_components=64, max_measurements=16, max_pairs=4)
    h.reset(2, sc.poses[0], sc.map_w, sc.map_m, sc.map_P)
    fr = sc.next_frame()
    with pytest.raises(capi.RbphdError) as e:
        h.slam_update(fr.z, 0.5)
    assert e.value.code == capi.ERR_CAPACITY
    h.close()


def test_c2_shape_sampled_parity(capi, orc, synth):
    """BASELINE config 2 shape (2000 particles x 500 components x 100 measurements): the map update of
    the first particles against the oracle, WeightAlpha of one particle at that size, and
    size-independent properties over all particles."""
    P, N, M = 2000, 500, 100
    sc = synth.make_scene(P, N, M, seed=31)
    h = capi.Handle(sc.params, max_particles=P, max_components=2 * N, max_measurements=M)
    h.reset(P, sc.poses[0], sc.map_w, sc.map_m, sc.map_P)
    h.set_poses(sc.poses)
    S = 5
    ocfg = orc.make_config(sc.params)
    nav = orc.Navigator(ocfg, S, sc.poses[0], only_mapping=True)
    for i in range(S):
        nav.set_pose(i, sc.poses[i])
        nav.set_map(i, sc.map_w, sc.map_m, sc.map_P)
    fr = sc.next_frame()
    h.slam_update(fr.z, fr.u, only_mapping=True)
    nav.slam_update(fr.z, fr.u)
    counts = h.get_map_counts()
    assert counts.min() > 0 and counts.max() <= sc.params["max_quantity"]
    for i in range(S):
        assert_maps_equal(h.get_map(i), nav.get_map(i), f"c2 particle {i}")
    # WeightAlpha at this size through the stage entry point
    pose = sc.poses[0]
    pred = orc.predict(ocfg, pose, sc.map_w, sc.map_m, sc.map_P, fr.z)[:3]
    corr = nav.get_map(0)
    exp = orc.weight_alpha(ocfg, pose, fr.z, pred, corr)
    got = h.stage_weight_alpha(pose, fr.z, pred, corr)
    assert got["J"] == exp["J"] and exp["J"] > 100
    for k in ("pcount", "ccount", "ploglik", "cloglik", "setloglik"):
        assert close_rel(got[k], exp[k]), (k, got[k], exp[k])
    # a full SLAM frame over all particles: weights normalised, ancestors a valid non-decreasing wheel
    h.update(fr.reading, synth.DT, fr.gauss)
    fr2 = sc.next_frame()
    best, res = h.slam_update(fr2.z, fr2.u)
    w = h.get_weights()
    assert np.all(np.isfinite(w)) and abs(w.sum() - 1.0) < 1e-9
    anc = h.get_ancestors()
    assert np.all(np.diff(anc) >= 0) and 0 <= anc.min() and anc.max() < P
    assert 0 <= best < P
    if res:
        assert np.allclose(w, 1.0 / P)
    h.close()


def test_c4_shape_sampled_parity(capi, orc, synth):
    """BASELINE config 4's per-particle shape (2000 components x 500 measurements, MaxQuantity 4000): a few
    particles against the oracle over consecutive SLAM frames.  Exercises the large-size lanes of the
    kernel (radix-select before the candidate sort, counting sort of > 4096 pairs, overflow lists,
    thousands of map-estimate points) that the small tests never reach."""
    P, N, M = 4, 2000, 500
    sc = synth.make_scene(P, N, M, seed=33)
    h = capi.Handle(sc.params, max_particles=P, max_components=2 * N, max_measurements=M, max_pairs=16 * M)
    nav = orc.Navigator(orc.make_config(sc.params), P, sc.poses[0])
    h.reset(P, sc.poses[0], sc.map_w, sc.map_m, sc.map_P)
    h.set_poses(sc.poses)
    for i in range(P):
        nav.set_pose(i, sc.poses[i])
        nav.set_map(i, sc.map_w, sc.map_m, sc.map_P)
    for f in range(3):
        fr = sc.next_frame()
        h.update(fr.reading, synth.DT, fr.gauss)
        nav.update(fr.reading, synth.DT, fr.gauss)
        # the Parallel.For body alone (maps + alphas); with 500 measurements exp() of the log-likelihood
        # underflows to 0 in FP64 on both sides, so compare the parts of WeightAlpha through the stage API too
        h.upload_frame_inputs(None, fr.z)
        gbest, gres = h.slam_update(fr.z, fr.u)
        obest, ores, oanc = nav.slam_update(fr.z, fr.u)
        assert (gbest, gres) == (obest, ores), f"frame {f}"
        assert h.get_ancestors().tolist() == oanc.tolist()
        assert h.get_alphas().tolist() == nav.get_alphas().tolist() or close_rel(h.get_alphas(), nav.get_alphas())
        counts = h.get_map_counts()
        for i in range(P):
            om = nav.get_map(i)
            assert counts[i] == len(om[0]), f"frame {f} particle {i}: {counts[i]} vs {len(om[0])}"
            assert_maps_equal(h.get_map(i), om, f"frame {f} particle {i}")
    # WeightAlpha parts at this size (J ~ 1600 points, thousands of components)
    pose = nav.get_poses()[0]
    prior = nav.get_map(0)
    fr = sc.next_frame()
    ocfg = orc.make_config(sc.params)
    pred = orc.predict(ocfg, pose, *prior, fr.z)[:3]
    corr = orc.prune(ocfg, *orc.correct(ocfg, pose, *pred, fr.z))
    exp = orc.weight_alpha(ocfg, pose, fr.z, pred, corr)
    got = h.stage_weight_alpha(pose, fr.z, pred, corr)
    assert got["J"] == exp["J"] and exp["J"] > 1000
    for k in ("pcount", "ccount", "ploglik", "cloglik", "setloglik"):
        assert close_rel(got[k], exp[k]), (k, got[k], exp[k])
    h.close()


def test_c4_shape_saturated_parity(capi, orc, synth):
    """The regime bench.py times: config 4's per-particle shape followed for 14 consecutive frames, by which
    time the map has grown from 2000 to ~3850 components and the MaxQuantity cut (4000 heaviest candidates,
    PHD:924-927) is active on every frame.  Two particles against the oracle."""
    counts = []
    nres = run_both(capi, orc, synth, P=2, N=2000, M=500, frames=14, seed=35, merge_floor=64.0, max_pairs=16 * 500,
                    final_counts=counts)
    assert nres == 0    # (the set likelihood underflows at 500 measurements: weights 0, never depleted)
    assert min(counts) > 3500, counts   # saturated: the MaxQuantity cut has been deciding the survivors


def test_big_particle_fallback_lanes(capi, orc, synth):
    """One particle with a map far beyond the shared-memory capacities of the fused kernel (more than 8192
    gated pairs, tens of thousands of pre-prune candidates, thousands of map-estimate points): the
    slab-resident (global memory) lanes must give the same answer.  Shape is towards BASELINE config 3."""
    P, N, M = 2, 9000, 1000
    sc = synth.make_scene(P, N, M, seed=35)
    h = capi.Handle(sc.params, max_particles=P, max_components=2 * N, max_measurements=M, max_pairs=24 * M)
    nav = orc.Navigator(orc.make_config(sc.params), P, sc.poses[0], only_mapping=True)
    h.reset(P, sc.poses[0], sc.map_w, sc.map_m, sc.map_P)
    h.set_poses(sc.poses)
    for i in range(P):
        nav.set_pose(i, sc.poses[i])
        nav.set_map(i, sc.map_w, sc.map_m, sc.map_P)
    for f in range(4):
        fr = sc.next_frame()
        h.counters(reset=True)
        h.slam_update(fr.z, fr.u, only_mapping=True)
        nav.slam_update(fr.z, fr.u)
        for i in range(P):
            assert_maps_equal(h.get_map(i), nav.get_map(i), f"frame {f} particle {i}")
    ctr = h.counters()
    assert ctr["pairs"] / ctr["particle_frames"] > 8192, ctr   # last frame: beyond the shared-memory pair lists
    h.close()


def test_c3_shape_sampled_parity(capi, orc, synth):
    """BASELINE config 3's per-particle shape (mapping-only, 50 000 components x 1000 measurements per frame,
    box scaled x25): two filters on the GPU, one of them against the oracle over two frames (the oracle needs
    ~25 s per particle-frame at this size), the other checked through size-independent properties."""
    P, N, M = 2, 50000, 1000
    sc = synth.make_workload("c3", P=P, N=N, M=M, seed=35)
    h = capi.Handle(sc.params, max_particles=P, max_components=2 * N, max_measurements=M, max_pairs=16 * M)
    h.reset(P, sc.poses[0], sc.map_w, sc.map_m, sc.map_P)
    h.set_poses(sc.poses)
    nav = orc.Navigator(orc.make_config(sc.params), 1, sc.poses[0], only_mapping=True)
    nav.set_pose(0, sc.poses[0])
    nav.set_map(0, sc.map_w, sc.map_m, sc.map_P)
    for f in range(2):
        fr = sc.next_frame()
        h.slam_update(fr.z, fr.u, only_mapping=True)
        nav.slam_update(fr.z, fr.u)
        counts = h.get_map_counts()
        om = nav.get_map(0)
        assert counts[0] == len(om[0]) and counts[0] > N
        assert_maps_equal(h.get_map(0), om, f"c3 frame {f}")
        w1, m1, P1 = h.get_map(1)
        assert np.all(np.isfinite(w1)) and np.all(np.isfinite(m1)) and np.all(np.isfinite(P1))
        assert abs(len(w1) - counts[0]) < 0.01 * N
    assert np.array_equal(h.get_weights(), np.full(P, 1.0 / P))   # OnlyMapping leaves the weights alone (PHD:334)
    h.close()
