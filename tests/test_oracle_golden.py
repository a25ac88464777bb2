"""Pins the CPU oracle against the reference's own NUnit known-answer tests.

Each test restates one test of /root/reference/mono-rfs-lib/Test/*.cs (file:line cited) with the
same inputs and the same tolerance.  Nothing here reads /root/reference at run time.
"""
import math

import numpy as np
import pytest


# ----------------------------------------------------------------------------------------------
# helpers restating Gaussian.Fuse / Canonical / Multiply (GAUSS:165-180, 253-288) with numpy's
# SVD pseudo-inverse / pseudo-determinant -- i.e. what Accord does -- for the EXPECTED values only
# ----------------------------------------------------------------------------------------------
def _pdet(a):
    s = np.linalg.svd(a, compute_uv=False)
    tol = np.finfo(float).eps * max(a.shape) * s.max()
    return float(np.prod(s[s > tol]))


class G:
    def __init__(self, mean, cov, w):
        self.mean = np.asarray(mean, float)
        self.cov = np.asarray(cov, float)
        self.w = w
        self.det = _pdet(self.cov)
        self.inv = np.linalg.pinv(self.cov)
        self.mult = (2 * math.pi) ** int(-len(self.mean) / 2) / math.sqrt(self.det)   # C# integer division truncates toward zero
        self.vec = self.inv @ self.mean

    @property
    def bias(self):
        return math.log(self.mult) - 0.5 * self.mean @ (self.inv @ self.mean)

    @staticmethod
    def canonical(vec, mat, w):
        cov = np.linalg.pinv(mat)
        g = G(cov @ vec, cov, w)
        g.inv, g.vec = mat, vec
        return g

    @staticmethod
    def multiply(a, b):
        fused = G.canonical(a.vec + b.vec, a.inv + b.inv, 1.0)
        logscale = a.bias + b.bias - fused.bias
        fused.w = math.exp(logscale + math.log(a.w) + math.log(b.w))
        return fused


def _match_multiset(expected, got, tol):
    """Order-insensitive comparison as in PHDNavigatorTest.cs:178-192."""
    gw, gm, gP = [list(x) for x in got]
    for (w, m, P) in expected:
        found = None
        for i in range(len(gw)):
            if abs(gw[i] - w) <= tol and np.all(np.abs(gm[i] - m) <= tol) and np.all(np.abs(gP[i] - P) <= tol):
                found = i
                break
        assert found is not None, f"component not found: w={w} m={m}"
        del gw[found], gm[found], gP[found]


@pytest.fixture()
def lin(orc):
    """PHDNavigatorTest.cs:58-77: Linear2D, pose (1,2), measurer range 6.5."""
    p = orc.linear2d_params(measurer=[6.5, 0, 0, 0, 0, 0, 0])
    return orc.make_config(p), np.array([1.0, 2.0])


# ------------------------------------------------------------------ PHDNavigatorTest.cs
def test_predict_initial(orc, lin):   # PHDNavigatorTest.cs:85-104
    cfg, pose = lin
    w, m, P, nb = orc.predict(cfg, pose, np.zeros(0), np.zeros((0, 3)), np.zeros((0, 3, 3)), [[2, 3, 0]])
    assert len(w) == 1 and nb == 1
    assert np.allclose(m[0], [3, 5, 0], atol=1e-5)
    assert np.allclose(P[0], np.eye(3) * 1e-2, atol=1e-5)
    assert abs(w[0] - 0.05) <= 1e-5


def test_predict_known(orc, lin):   # PHDNavigatorTest.cs:106-126
    cfg, pose = lin
    w, m, P, nb = orc.predict(cfg, pose, [1.0], [[3, 5, 0]], [np.eye(3)], [[2, 3, 0]])
    assert len(w) == 1 and nb == 0
    assert np.allclose(m[0], [3, 5, 0], atol=1e-5) and np.allclose(P[0], np.eye(3), atol=1e-5)
    assert abs(w[0] - 1.0) <= 1e-5


def test_correct(orc, lin):   # PHDNavigatorTest.cs:128-193 (pins the UNGATED per-pair formulas)
    cfg, pose = lin
    PD, clutter = 0.9, 3e-7
    I = np.eye(3)
    comp1 = G([3, 5, 0], I, 0.8)
    comp2 = G([7, 5, 0], 4.0 * I, 1.4)
    mcov = np.zeros((3, 3))
    mcov[0, 0] = mcov[1, 1] = 5e-4
    gz1 = G([1 + 2, 2 + 3, 0], mcov, 1.0)
    gz2 = G([1 + 5, 2 + 3, 0], mcov, 1.0)
    z11, z12 = G.multiply(gz1, comp1), G.multiply(gz1, comp2)
    z21, z22 = G.multiply(gz2, comp1), G.multiply(gz2, comp2)
    s1, s2 = z11.w + z12.w, z21.w + z22.w
    expected = [
        (0.8 * (1 - PD), comp1.mean, comp1.cov),
        (1.4 * (1 - PD), comp2.mean, comp2.cov),
        (z11.w * PD / (clutter + PD * s1), z11.mean, z11.cov),
        (z12.w * PD / (clutter + PD * s1), z12.mean, z12.cov),
        (z21.w * PD / (clutter + PD * s2), z21.mean, z21.cov),
        (z22.w * PD / (clutter + PD * s2), z22.mean, z22.cov),
    ]
    got = orc.correct(cfg, pose, [0.8, 1.4], [[3, 5, 0], [7, 5, 0]], [I, 4.0 * I],
                      [[2, 3, 0], [5, 3, 0]], gate_radius=-1.0)
    assert len(got[0]) == 6
    _match_multiset(expected, got, 1e-5)


def test_correct_gated_head_behaviour(orc, lin):
    """With DensityDistanceThreshold = 0.5 (CFG:74, PHD:882) only components within the gate update."""
    cfg, pose = lin
    I = np.eye(3)
    got = orc.correct(cfg, pose, [0.8, 1.4], [[3, 5, 0], [7, 5, 0]], [I, 4.0 * I], [[2, 3, 0], [5, 3, 0]])
    # z1 maps to (3,5,0): only comp1 is within 0.5; z2 maps to (6,5,0): nothing within 0.5
    assert len(got[0]) == 3


def test_prune(orc, lin):   # PHDNavigatorTest.cs:195-265
    cfg, _ = lin
    I = np.eye(3)
    mw, md = 1e-3, 0.3
    big = [([-12, -24, -54], I, 23.0), ([-80, -22, -12], 4.0 * I, 1.0), ([-63, -11, -95], 0.1 * I, 6.0)]
    irr = [([12, 24, 54], I, 0.3 * mw), ([80, 22, 12], 4.0 * I, 0.8 * mw), ([63, 11, 95], 0.1 * I, 0.99 * mw),
           ([23, 19, 73], I, 0.0)]
    mg1 = [([0, 0, 0], I, 1.0), ([0, md, 0], I, 0.6), ([0, md / 2, 0], I, 1.2)]
    mg2 = [([99 - md / 6, 99, 99], I, 0.9), ([99, 99 - md / 6, 99], I, 0.5), ([99, 99, 99 - md / 6], I, 1.1)]
    allc = big + irr + mg1 + mg2
    w = [c[2] for c in allc]
    m = [c[0] for c in allc]
    P = [c[1] for c in allc]
    expected = [(c[2], np.asarray(c[0], float), c[1]) for c in big]
    for grp in (mg1, mg2):
        ow, om, oP = orc.gaussian_merge([c[2] for c in grp], [c[0] for c in grp], [c[1] for c in grp])
        # cross-check Merge itself against the textbook moment match (GAUSS:308-327 comment)
        ws = np.array([c[2] for c in grp])
        ms = np.array([c[0] for c in grp], float)
        mu = (ws[:, None] * ms).sum(0) / ws.sum()
        cov = sum(wi * (I + np.outer(mi - mu, mi - mu)) for wi, mi in zip(ws, ms)) / ws.sum()
        assert abs(ow - ws.sum()) < 1e-12 and np.allclose(om, mu, atol=1e-9) and np.allclose(oP, cov, atol=1e-7)
        expected.append((ow, om, oP))
    got = orc.prune(cfg, w, m, P)
    assert len(got[0]) == 5
    _match_multiset(expected, got, 1e-5)


# ------------------------------------------------------------------ GraphCombinatoricsTest.cs
def _threecomp():   # GraphCombinatoricsTest.cs:50-64
    d = np.zeros((6, 6), dtype=np.uint8)
    for (i, k) in [(0, 0), (1, 0), (1, 1), (2, 2), (2, 3), (2, 4), (3, 3), (4, 4), (5, 5)]:
        d[i, k] = 1
    return d


def test_connected_components_empty(orc):   # :66-76
    assert orc.connected_components(np.zeros((100, 100), dtype=np.uint8)) == 0


def test_connected_components_one(orc):   # :78-94
    assert orc.connected_components(np.ones((10, 10), dtype=np.uint8)) == 1


def test_connected_components_count(orc):   # :96-126
    d = _threecomp()
    assert orc.connected_components(d) == 3
    d[1, 2] = 1
    assert orc.connected_components(d) == 2
    d[5, 4] = 1
    assert orc.connected_components(d) == 1


def test_assignment_values(orc):   # :174-198 through the enumerators' value column
    d = _threecomp()
    v = d.astype(float)
    # identity assignment value 6 (AssignmentValue1)
    assert sum(v[i, i] for i in range(6)) == 6
    v[1, 0] = 100
    matches = [1, 0, 4, 0, 4, 5]
    assert sum(v[i, matches[i]] for i in range(6)) == 103


def test_linear_assignment_unique(orc):   # :200-214
    v = np.zeros((10, 10))
    d = np.zeros((10, 10), dtype=np.uint8)
    for i in range(10):
        v[i, i] = (i + 1) / 2.0
        d[i, i] = 1
    assert list(orc.hungarian(v, d, 0.0)) == list(range(10))


PROFIT3 = np.array([[6, 8, 5], [7, 3, 4], [9, 8, 7]], float)


def test_linear_assignment_1(orc):   # :216-230
    assert list(orc.hungarian(PROFIT3)) == [1, 0, 2]


def test_linear_assignment_2(orc):   # :232-246
    v = np.array([[0, 2, 5], [3, 0, 6], [1, 2, 0]], float)
    d = np.array([[0, 1, 1], [1, 0, 1], [1, 1, 0]], dtype=np.uint8)
    assert list(orc.hungarian(v, d, 0.0)) == [2, 0, 1]


def test_linear_assignment_3(orc):   # :248-256
    d = _threecomp()
    v = d.astype(float)
    v[4, 2] = 3
    d[4, 2] = 1
    d[2, 2] = 0
    v[2, 2] = 0
    assert list(orc.hungarian(v, d, 0.0)) == [0, 1, 4, 3, 2, 5]


def test_lexicographical_full(orc):   # :258-281
    perms, vals = orc.lexicographical(PROFIT3, 3)
    assert perms.tolist() == [[0, 1, 2], [0, 2, 1], [1, 0, 2], [1, 2, 0], [2, 0, 1], [2, 1, 0]]
    assert vals.tolist() == [sum(PROFIT3[i, p[i]] for i in range(3)) for p in perms.tolist()]


def test_lexicographical_no_duplicates(orc):   # :283-306
    perms, _ = orc.lexicographical(PROFIT3, 1)
    assert perms.tolist() == [[0, 2, 1], [1, 2, 0], [2, 1, 0]]


def test_murty_children_with_duplicates(orc):   # :308-334
    ch = orc.murty_children([0, 1, 2, 3, 4], [(1, 1)], [(0, 2)])
    exp = [([(1, 1)], [(0, 2), (0, 0)]),
           ([(1, 1), (0, 0)], [(0, 2), (2, 2)]),
           ([(1, 1), (0, 0), (2, 2)], [(0, 2), (3, 3)])]
    assert len(ch) == 3
    for e in exp:
        assert e in ch


def test_murty_children_none(orc):   # :336-355
    ch = orc.murty_children([0, 1, 2, 3, 4], [(i, i) for i in range(5)], [(1, 2)])
    assert ch == []


def test_murty_full_small(orc):   # :357-382
    perms, vals = orc.murty(PROFIT3)
    assert perms.tolist() == [[1, 0, 2], [1, 2, 0], [2, 0, 1], [0, 2, 1], [2, 1, 0], [0, 1, 2]]
    assert all(vals[i] >= vals[i + 1] for i in range(len(vals) - 1))


def test_murty_unique(orc):   # :384-404
    v = np.zeros((5, 5))
    d = np.eye(5, dtype=np.uint8)
    v[d == 1] = 1
    perms, _ = orc.murty(v, d, 0.0)
    assert perms.tolist() == [[0, 1, 2, 3, 4]]


# ------------------------------------------------------------------ SimulationTest.cs:225-270
def test_resample_invariants(orc):
    cfg = orc.make_config(orc.prm3d_params())
    rng = np.random.default_rng(7)
    seen0 = seen3 = 0
    iters = 10000
    for _ in range(iters):
        u = float(np.float32(rng.random()))
        w, best, anc, res = orc.normalize_resample(cfg, [0.11, 0.28, 0.31, 0.01, 0.29], u, force=2)   # ResampleParticles() alone
        assert res and np.allclose(w, 0.2)
        assert anc[best] == 2                       # best estimate is particle 2
        assert {1, 2, 4} <= set(anc.tolist())       # weights > 0.2 always survive
        assert all(anc[i] <= anc[i + 1] for i in range(4))
        seen0 += 0 in anc
        seen3 += 3 in anc
    assert seen0 < iters and seen3 < iters


def test_ess_threshold(orc):   # PHD:768-777
    cfg = orc.make_config(orc.prm3d_params(min_effective_particle=0.3))
    _, _, anc, res = orc.normalize_resample(cfg, [1, 1, 1, 1], 0.5)
    assert not res and anc.tolist() == [0, 1, 2, 3]
    w, best, anc, res = orc.normalize_resample(cfg, [100, 1e-3, 1e-3, 1e-3], 0.5)
    assert res and anc.tolist() == [0, 0, 0, 0] and best == 0 and np.allclose(w, 0.25)


# ------------------------------------------------------------------ LoopyPHDNavigatorTest.cs fixtures
IDENT = [0, 0, 0, 1, 0, 0, 0]


def test_pixel_range_fixtures(orc):   # LoopyPHDNavigatorTest.cs:169-175
    cfg = orc.make_config(orc.prm3d_params())
    for lm, z in [([0.2, 0, 1], [115.16312, 0, 1.019803903]),
                  ([0, 0.1, 1], [0, 57.58156, 1.004987562]),
                  ([0.1, 0, 2], [28.79078, 0, 2.002498439])]:
        assert np.allclose(orc.measure_perfect(cfg, IDENT, lm), z, rtol=0, atol=1e-6)
        assert np.allclose(orc.measure_to_map(cfg, IDENT, z), lm, rtol=0, atol=1e-6)


def test_information_matrix_at_identity(orc):   # LoopyPHDNavigatorTest.cs:90-123 (translation block)
    cfg = orc.make_config(orc.prm3d_params())
    H = orc.measurement_jacobian_l(cfg, IDENT, [0, 0, 1])
    info = H.T @ np.diag([1 / 2.0, 1 / 2.0, 1 / 1e-3]) @ H
    f2 = 575.8156 ** 2
    assert np.allclose(info, np.diag([f2 / 2.0, f2 / 2.0, 1 / 1e-3]), rtol=1e-3)


def _pose_close(a, b, tol):
    return np.all(np.abs(np.asarray(a) - np.asarray(b)) < tol)


def test_fit_measurement_already_fine(orc):   # :194-208
    cfg = orc.make_config(orc.prm3d_params())
    assert _pose_close(orc.fit_to_measurement(cfg, IDENT, [0, 0, 1], [0, 0, 1]), IDENT, 1e-5)


def test_fit_measurement_only_translation(orc):   # :210-224
    cfg = orc.make_config(orc.prm3d_params())
    assert _pose_close(orc.fit_to_measurement(cfg, IDENT, [0, 0, 1], [0, 0, 2.5]), [0, 0, 1.5, 1, 0, 0, 0], 1e-5)


def test_fit_measurement_only_rotation(orc):   # :226-243
    cfg = orc.make_config(orc.prm3d_params())
    s = 1 / math.sqrt(2)
    exp = [0, 0, 0, math.cos(math.pi / 8), -math.sin(math.pi / 8), 0, 0]
    assert _pose_close(orc.fit_to_measurement(cfg, IDENT, [0, 0, 1], [0, s, s]), exp, 1e-5)


def test_fit_measurement_unmeasurable(orc):   # :245-262
    cfg = orc.make_config(orc.prm3d_params())
    exp = [0, 0, 0, math.cos(math.pi / 4), -math.sin(math.pi / 4), 0, 0]
    assert _pose_close(orc.fit_to_measurement(cfg, IDENT, [0, 0, 1], [0, 1, 0]), exp, 1e-5)


def test_fit_measurement_general(orc):   # :264-278
    cfg = orc.make_config(orc.prm3d_params())
    q = np.array([1, 2, 3, 4.0])
    q /= np.linalg.norm(q)
    pose0 = [2.0, 0.2, 0.1, *q]
    z = [-120, 50, 1.3]
    lm = [0.1, -1.0, 1.2]
    fitted = orc.fit_to_measurement(cfg, pose0, z, lm)
    assert np.allclose(orc.measure_perfect(cfg, fitted, lm), z, atol=1e-5)


# ------------------------------------------------------------------ Pose3DTest.cs / QuaternionTest.cs
def _poses(orc):   # Pose3DTest.cs:48-57
    def n(q):
        return q / np.linalg.norm(q)
    a = np.array([0.1, 0.3, 0.2, *n(orc.quat_from_ypr(0.4, 1.6, 0.1))])
    b = np.array([0.5, -0.4, 0.7, *n(orc.quat_from_ypr(0.4, 0.6, 0.5))])
    return a, b


def test_pose_add_subtract(orc):   # Pose3DTest.cs:65-76
    a, _ = _poses(orc)
    odo = [0.12, 2.17, 1.03, 0.21, 0.05, 1.05]
    rec = orc.pose_diff_odometry(orc.pose_add_odometry(a, odo), a)
    assert np.allclose(rec, odo, atol=1e-3)


def test_pose_subtract_add(orc):   # Pose3DTest.cs:78-90
    a, b = _poses(orc)
    rec = orc.pose_add_odometry(b, orc.pose_diff_odometry(a, b))
    assert np.allclose(rec, a, atol=1e-3)


def test_quat_exp_log(orc):   # QuaternionTest.cs:52-62
    q = orc.quat_from_ypr(0.5, 0.2, 0.3)
    assert np.allclose(orc.quat_exp(orc.quat_log(q)), q, atol=1e-3)


def test_quat_log_exp(orc):   # QuaternionTest.cs:64-73
    lie = [0.5, 0.2, 0.3]
    assert np.allclose(orc.quat_log(orc.quat_exp(lie)), lie, atol=1e-3)


def test_quat_add_subtract(orc):   # QuaternionTest.cs:75-85
    q = orc.quat_from_ypr(0.4, 0.6, 0.1)
    lie = np.array([0.5, 0.2, 0.3])
    added = orc.quat_mul(q, orc.quat_exp(0.5 * lie))                       # QUAT:165-168
    conj = q * np.array([1, -1, -1, -1])
    rec = 2 * orc.quat_log(orc.quat_mul(conj, added))                      # QUAT:175-178
    assert np.allclose(rec, lie, atol=1e-3)


def test_quat_vector_rotator(orc):   # QuaternionTest.cs:100-131
    f = np.array([1, 2.3, 3.0])
    f /= np.linalg.norm(f)
    t = np.array([4.8, 3, 2.0])
    t /= np.linalg.norm(t)
    assert np.allclose(orc.quat_to_matrix(orc.quat_vector_rotator(f, t)) @ f, t, atol=1e-5)
    assert np.allclose(orc.quat_to_matrix(orc.quat_vector_rotator(f, f)) @ f, f, atol=1e-5)


def test_jacobian_matches_independent_restatement(orc):
    """isam2/PixelRangeFactor.cpp:76-105 states H = Jproj(local) * R(q)^T independently; check it
    numerically by central differences of MeasurePerfect (PRM:138-177)."""
    cfg = orc.make_config(orc.prm3d_params())
    q = np.array([0.9, 0.1, -0.3, 0.2])
    q /= np.linalg.norm(q)
    pose = [0.3, -0.2, 0.1, *q]
    lm = np.array([0.5, 0.4, 1.7])
    H = orc.measurement_jacobian_l(cfg, pose, lm)
    num = np.zeros((3, 3))
    for j in range(3):
        d = np.zeros(3)
        d[j] = 1e-6
        num[:, j] = (orc.measure_perfect(cfg, pose, lm + d) - orc.measure_perfect(cfg, pose, lm - d)) / 2e-6
    assert np.allclose(H, num, rtol=1e-6, atol=1e-5)


def test_loglike2d_gradient_reference_test_pins_the_tempered_average():
    """LoopyPHDNavigatorTest.LogLike2D (LoopyPHDNavigatorTest.cs:352-425) restated: the analytic pose gradient of
    QuasiSetLogLikelihood against central differences of its value on a grid of poses, tolerance 0.5.  It passes when
    TemperedAverage (MX:400-440) divides by the SUM of the tempered weights; with the Euclidean norm -- what Accord's
    Normalize() does elsewhere in the reference -- a fifth of the grid fails (oracle/README.md D10)."""
    from oracle import orc
    p = orc.linear2d_params()
    p.update(R=np.diag([5e-2, 5e-2, 1.0]), pd=0.9, clutter=0.0)      # Setup(): only the measurement covariance is set
    cfg = orc.make_config(p)
    lm = np.array([[0, 1.45, 0], [0, 0.65, 0], [1.0, 0, 0]])
    z = np.array([[0, 1, 0], [0.2, 0.6, 0]])
    n = 101
    xs = (np.arange(n) / (n - 1) - 0.5) / 0.5
    frac = {}
    try:
        for mode in (1, 0):
            orc.lib().orc_set_tempered_norm(mode)
            L, G = np.zeros((n, n)), np.zeros((n, n, 2))
            for i, x in enumerate(xs):
                for k, y in enumerate(xs):
                    L[i, k], G[i, k] = orc.quasi_set_loglikelihood_gradient(cfg, [x, y], lm, z)
            num_x = (L[2:, 1:-1] - L[:-2, 1:-1]) / (xs[2:] - xs[:-2])[:, None]
            num_y = (L[1:-1, 2:] - L[1:-1, :-2]) / (xs[2:] - xs[:-2])[None, :]
            ok = (np.abs(num_x - G[1:-1, 1:-1, 0]) < 0.5) & (np.abs(num_y - G[1:-1, 1:-1, 1]) < 0.5)
            frac[mode] = float(ok.mean())
    finally:
        orc.lib().orc_set_tempered_norm(0)
    assert frac[1] == 1.0, frac
    assert frac[0] < 0.95, frac
