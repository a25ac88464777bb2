"""ctypes binding of the CPU oracle (oracle/_build/liborc.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under monorfs_b200/ imports this module.
"""
import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liborc.so")

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_u8_p = C.POINTER(C.c_uint8)


class OrcConfig(C.Structure):
    _fields_ = [
        ("model", C.c_int32),
        ("max_quantity", C.c_int32),
        ("gate_metric", C.c_int32),
        ("nthreads", C.c_int32),
        ("R", C.c_double * 9),
        ("Q", C.c_double * 36),
        ("pd", C.c_double),
        ("clutter", C.c_double),
        ("birth_cov", C.c_double * 9),
        ("birth_weight", C.c_double),
        ("min_weight", C.c_double),
        ("merge_threshold", C.c_double),
        ("exploration_threshold", C.c_double),
        ("density_distance_threshold", C.c_double),
        ("min_effective_particle", C.c_double),
        ("visibility_ramp", C.c_double * 3),
        ("measurer", C.c_double * 7),
    ]


def build(force=False):
    if force or not os.path.exists(_LIB_PATH) or (
        os.path.getmtime(_LIB_PATH) < max(
            os.path.getmtime(os.path.join(_HERE, f)) for f in ("rbphd_oracle.cpp", "rbphd_oracle.h"))
        and os.path.exists("/usr/bin/make")
    ):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_fuzzy_visible.restype = C.c_double
        _lib.orc_gaussian_evaluate.restype = C.c_double
        _lib.orc_set_loglikelihood.restype = C.c_double
        _lib.orc_quasi_set_loglikelihood.restype = C.c_double
        _lib.orc_quasi_set_loglikelihood_gradient.restype = C.c_double
        _lib.orc_ospa.restype = C.c_double
        _lib.orc_nav_new.restype = C.c_void_p
    return _lib


def prm3d_params(**over):
    """Reference defaults for the pixel-range 3-D model (CFG:77-91, 238-263; PRM:70-73)."""
    R = np.diag([2.0, 2.0, 1e-3])
    p = dict(
        model=0, max_quantity=600, gate_metric=0, nthreads=8,
        R=R, Q=np.diag([5e-3] * 3 + [2e-4] * 3), pd=0.9, clutter=3e-7,
        birth_cov=np.eye(3) * 1e-2, birth_weight=0.05, min_weight=1e-3, merge_threshold=0.3,
        exploration_threshold=1e-5, density_distance_threshold=0.5, min_effective_particle=0.1,
        visibility_ramp=None,
        measurer=[575.8156, 0.1, 2.0, -320, -240, 640, 480],
    )
    p.update(over)
    if p["visibility_ramp"] is None:
        Rm = np.asarray(p["R"], dtype=np.float64).reshape(3, 3)
        p["visibility_ramp"] = [3 * math.sqrt(Rm[0, 0]), 3 * math.sqrt(Rm[1, 1]), 3 * math.sqrt(Rm[2, 2])]
    return p


def linear2d_params(**over):
    """Config.SetLinear2DDefaults (CFG:210-232) + Linear2DMeasurer(range) (Linear2DMeasurer.cs:60-70)."""
    R = np.zeros((3, 3))
    R[0, 0] = R[1, 1] = 5e-4
    Q = np.zeros((6, 6))
    Q[0, 0] = Q[1, 1] = 2.0
    p = prm3d_params(model=1, R=R, Q=Q,
                     visibility_ramp=[3 * math.sqrt(5e-4), 3 * math.sqrt(5e-4), 0.0],
                     measurer=[2.0, 0, 0, 0, 0, 0, 0])
    p.update(over)
    return p


def make_config(p):
    c = OrcConfig()
    for k in ("model", "max_quantity", "gate_metric", "nthreads"):
        setattr(c, k, int(p[k]))
    for k in ("pd", "clutter", "birth_weight", "min_weight", "merge_threshold", "exploration_threshold",
              "density_distance_threshold", "min_effective_particle"):
        setattr(c, k, float(p[k]))
    for k, n in (("R", 9), ("Q", 36), ("birth_cov", 9), ("visibility_ramp", 3), ("measurer", 7)):
        arr = np.asarray(p[k], dtype=np.float64).reshape(-1)
        assert arr.size == n, (k, arr.size)
        getattr(c, k)[:] = arr.tolist()
    return c


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(c_double_p)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(c_int_p)


def _pose(pose):
    p = np.zeros(7)
    pose = np.asarray(pose, dtype=np.float64).reshape(-1)
    p[:pose.size] = pose
    return _d(p)


def _map_in(w, m, P):
    w = np.ascontiguousarray(w, dtype=np.float64).reshape(-1)
    m = np.ascontiguousarray(m, dtype=np.float64).reshape(-1, 3)
    P = np.ascontiguousarray(P, dtype=np.float64).reshape(-1, 3, 3)
    assert w.shape[0] == m.shape[0] == P.shape[0]
    return w, m, P


def _z(z):
    z = np.ascontiguousarray(z, dtype=np.float64).reshape(-1, 3)
    return z


def _map_out(cap):
    return np.zeros(cap), np.zeros((cap, 3)), np.zeros((cap, 3, 3))


def _p(a):
    return a.ctypes.data_as(c_double_p)


# ---------------------------------------------------------------- geometry
def quat_mul(a, b):
    out = np.zeros(4)
    lib().orc_quat_mul(_d(a)[1], _d(b)[1], _p(out))
    return out


def quat_exp(lie):
    out = np.zeros(4)
    lib().orc_quat_exp(_d(lie)[1], _p(out))
    return out


def quat_log(q):
    out = np.zeros(3)
    lib().orc_quat_log(_d(q)[1], _p(out))
    return out


def quat_sqrt(q):
    out = np.zeros(4)
    lib().orc_quat_sqrt(_d(q)[1], _p(out))
    return out


def quat_from_ypr(yaw, pitch, roll):
    out = np.zeros(4)
    lib().orc_quat_from_ypr(C.c_double(yaw), C.c_double(pitch), C.c_double(roll), _p(out))
    return out


def quat_to_matrix(q):
    out = np.zeros(9)
    lib().orc_quat_to_matrix(_d(q)[1], _p(out))
    return out.reshape(3, 3)


def quat_vector_rotator(a, b):
    out = np.zeros(4)
    lib().orc_quat_vector_rotator(_d(a)[1], _d(b)[1], _p(out))
    return out


def pose_from_state(s):
    out = np.zeros(7)
    lib().orc_pose_from_state(_d(s)[1], _p(out))
    return out


def pose_add_odometry(pose, delta):
    out = np.zeros(7)
    lib().orc_pose_add_odometry(_d(pose)[1], _d(delta)[1], _p(out))
    return out


def pose_diff_odometry(pose, origin):
    out = np.zeros(6)
    lib().orc_pose_diff_odometry(_d(pose)[1], _d(origin)[1], _p(out))
    return out


def measure_perfect(cfg, pose, m):
    out = np.zeros(3)
    lib().orc_measure_perfect(C.byref(cfg), _pose(pose)[1], _d(m)[1], _p(out))
    return out


def measurement_jacobian_l(cfg, pose, m):
    out = np.zeros(9)
    lib().orc_measurement_jacobian_l(C.byref(cfg), _pose(pose)[1], _d(m)[1], _p(out))
    return out.reshape(3, 3)


def measure_to_map(cfg, pose, z):
    out = np.zeros(3)
    lib().orc_measure_to_map(C.byref(cfg), _pose(pose)[1], _d(z)[1], _p(out))
    return out


def fuzzy_visible(cfg, z):
    return lib().orc_fuzzy_visible(C.byref(cfg), _d(z)[1])


def fit_to_measurement(cfg, pose0, z, landmark):
    out = np.zeros(7)
    lib().orc_fit_to_measurement(C.byref(cfg), _d(pose0)[1], _d(z)[1], _d(landmark)[1], _p(out))
    return out


def gaussian_evaluate(m, P, x):
    return lib().orc_gaussian_evaluate(_d(m)[1], _d(P)[1], _d(x)[1])


def gaussian_merge(w, m, P):
    w, m, P = _map_in(w, m, P)
    ow = C.c_double()
    om, oP = np.zeros(3), np.zeros(9)
    lib().orc_gaussian_merge(len(w), _p(w), _p(m), _p(P), C.byref(ow), _p(om), _p(oP))
    return ow.value, om, oP.reshape(3, 3)


# ---------------------------------------------------------------- stages
def predict(cfg, pose, w, m, P, z):
    w, m, P = _map_in(w, m, P)
    z = _z(z)
    cap = len(w) + len(z) + 1
    ow, om, oP = _map_out(cap)
    nb = C.c_int()
    n = lib().orc_predict(C.byref(cfg), _pose(pose)[1], len(w), _p(w), _p(m), _p(P), len(z), _p(z), cap,
                          _p(ow), _p(om), _p(oP), C.byref(nb))
    return ow[:n], om[:n], oP[:n], nb.value


def correct(cfg, pose, w, m, P, z, gate_radius=None):
    w, m, P = _map_in(w, m, P)
    z = _z(z)
    cap = len(w) * (len(z) + 1) + 1
    ow, om, oP = _map_out(cap)
    if gate_radius is None:
        gate_radius = cfg.density_distance_threshold
    n = lib().orc_correct(C.byref(cfg), _pose(pose)[1], len(w), _p(w), _p(m), _p(P), len(z), _p(z),
                          C.c_double(gate_radius), cap, _p(ow), _p(om), _p(oP))
    return ow[:n], om[:n], oP[:n]


def prune(cfg, w, m, P):
    w, m, P = _map_in(w, m, P)
    cap = len(w) + 1
    ow, om, oP = _map_out(cap)
    n = lib().orc_prune(C.byref(cfg), len(w), _p(w), _p(m), _p(P), cap, _p(ow), _p(om), _p(oP))
    return ow[:n], om[:n], oP[:n]


def best_map_estimate(w):
    w = np.ascontiguousarray(w, dtype=np.float64)
    cap = int(max(0.0, float(np.sum(w)))) + 8
    picks = np.zeros(cap, dtype=np.int32)
    n = lib().orc_best_map_estimate(len(w), _p(w), cap, picks.ctypes.data_as(c_int_p))
    return picks[:n]


def set_loglikelihood(cfg, pose, jm, z):
    jm = np.ascontiguousarray(jm, dtype=np.float64).reshape(-1, 3)
    z = _z(z)
    return lib().orc_set_loglikelihood(C.byref(cfg), _pose(pose)[1], len(jm), _p(jm), len(z), _p(z))


def set_depth_frame(depth_xy):
    """KinectMeasurer variant: depth_xy[x, y] float32 metres (NaN = no reading); None detaches."""
    if depth_xy is None:
        lib().orc_set_depth_frame(None, 0, 0)
        return
    d = np.ascontiguousarray(depth_xy, dtype=np.float32)
    lib().orc_set_depth_frame(d.ctypes.data_as(C.POINTER(C.c_float)), d.shape[0], d.shape[1])


def quasi_set_loglikelihood(cfg, pose, jm, z):
    jm = np.ascontiguousarray(jm, dtype=np.float64).reshape(-1, 3)
    z = _z(z)
    return lib().orc_quasi_set_loglikelihood(C.byref(cfg), _pose(pose)[1], len(jm), _p(jm), len(z), _p(z))


def quasi_set_loglikelihood_gradient(cfg, pose, jm, z):
    """(value, gradient[OdoSize]) of PHD:544-549."""
    jm = np.ascontiguousarray(jm, dtype=np.float64).reshape(-1, 3)
    z = _z(z)
    g = np.zeros(6)
    v = lib().orc_quasi_set_loglikelihood_gradient(C.byref(cfg), _pose(pose)[1], len(jm), _p(jm), len(z), _p(z), _p(g))
    return v, g[:(6 if cfg.model == 0 else 2)]


def measurement_jacobian_p(cfg, pose, m):
    out = np.zeros(18)
    lib().orc_measurement_jacobian_p(C.byref(cfg), _pose(pose)[1], _p(np.ascontiguousarray(m, dtype=np.float64)), _p(out))
    return out.reshape(3, 6)


def set_loglike_matrix(cfg, pose, jm, z, quasi=False):
    """(rows, cols, vals) of SetLogLikeMatrix, sorted by (row, col)."""
    jm = np.ascontiguousarray(jm, dtype=np.float64).reshape(-1, 3)
    z = _z(z)
    cap = len(jm) * len(z) + len(jm) + len(z) + 1
    rows, cols, vals = np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap)
    n = lib().orc_set_loglike_matrix(C.byref(cfg), _pose(pose)[1], len(jm), _p(jm), len(z), _p(z), int(quasi), cap,
                                     rows.ctypes.data_as(c_int_p), cols.ctypes.data_as(c_int_p), _p(vals))
    order = np.lexsort((cols[:n], rows[:n]))
    return rows[:n][order], cols[:n][order], vals[:n][order]


def weight_alpha(cfg, pose, z, pred, corr):
    pw, pm, pP = _map_in(*pred)
    cw, cm, cP = _map_in(*corr)
    z = _z(z)
    out = np.zeros(7)
    lib().orc_weight_alpha(C.byref(cfg), _pose(pose)[1], len(z), _p(z), len(pw), _p(pw), _p(pm), _p(pP),
                           len(cw), _p(cw), _p(cm), _p(cP), _p(out))
    return dict(alpha=out[0], setloglik=out[1], ploglik=out[2], cloglik=out[3], pcount=out[4],
                ccount=out[5], J=int(out[6]))


def normalize_resample(cfg, weights, u, best=0, force=False):
    w = np.array(weights, dtype=np.float64)
    anc = np.zeros(len(w), dtype=np.int32)
    b = C.c_int(best)
    res = lib().orc_normalize_resample(C.byref(cfg), len(w), _p(w), C.c_double(u), C.byref(b),
                                       anc.ctypes.data_as(c_int_p), int(force))
    return w, b.value, anc, bool(res)


# ---------------------------------------------------------------- graph
def _dense(val, defined):
    val = np.ascontiguousarray(val, dtype=np.float64)
    if defined is None:
        defined = np.ones(val.shape, dtype=np.uint8)
    defined = np.ascontiguousarray(defined, dtype=np.uint8)
    return val, defined


def hungarian(val, defined=None, defval=0.0):
    val, defined = _dense(val, defined)
    n = val.shape[0]
    match = np.zeros(n, dtype=np.int32)
    ok = lib().orc_hungarian(n, _p(val), defined.ctypes.data_as(c_u8_p), C.c_double(defval),
                             match.ctypes.data_as(c_int_p))
    return match if ok else None


def ospa(a, b, c=1.0, p=2.0):
    """(OSPA, cardinality error) between two landmark position sets (postanalysis/Plot.cs:531-581)."""
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1, 3))
    b = np.ascontiguousarray(np.asarray(b, dtype=np.float64).reshape(-1, 3))
    card = C.c_double(0)
    v = lib().orc_ospa(len(a), _p(a), len(b), _p(b), C.c_double(c), C.c_double(p), C.byref(card))
    return float(v), float(card.value)


def generate_measurements(cfg, pose, landmarks, uniforms, gauss, chol, clutter_u):
    """SimulatedVehicle.Measure (SIMV:243-295) with the caller's random numbers -> (z, assoc)."""
    lm = np.ascontiguousarray(np.asarray(landmarks, dtype=np.float64).reshape(-1, 3))
    un = np.ascontiguousarray(uniforms, dtype=np.float64)
    ga = np.ascontiguousarray(np.asarray(gauss, dtype=np.float64).reshape(-1, 3))
    ch = np.ascontiguousarray(chol, dtype=np.float64).reshape(3, 3)
    cu = np.ascontiguousarray(np.asarray(clutter_u, dtype=np.float64).reshape(-1, 3))
    z = np.zeros((len(lm) + len(cu), 3))
    assoc = np.zeros(len(lm) + len(cu), dtype=np.int32)
    n = lib().orc_generate_measurements(C.byref(cfg), _pose(pose)[1], len(lm), _p(lm), _p(un), _p(ga), _p(ch), len(cu),
                                        _p(cu), _p(z), assoc.ctypes.data_as(c_int_p))
    return z[:n].copy(), assoc[:n].copy()


def connected_components(defined):
    defined = np.ascontiguousarray(defined, dtype=np.uint8)
    return lib().orc_connected_components(defined.shape[0], defined.shape[1], defined.ctypes.data_as(c_u8_p))


def lexicographical(val, modelsize, defined=None, defval=0.0, cap=1024):
    val, defined = _dense(val, defined)
    n = val.shape[0]
    perms = np.zeros((cap, n), dtype=np.int32)
    values = np.zeros(cap)
    cnt = lib().orc_lexicographical(n, _p(val), defined.ctypes.data_as(c_u8_p), C.c_double(defval),
                                    int(modelsize), cap, perms.ctypes.data_as(c_int_p), _p(values))
    return perms[:cnt], values[:cnt]


def murty(val, defined=None, defval=0.0, cap=1024):
    val, defined = _dense(val, defined)
    n = val.shape[0]
    perms = np.zeros((cap, n), dtype=np.int32)
    values = np.zeros(cap)
    cnt = lib().orc_murty(n, _p(val), defined.ctypes.data_as(c_u8_p), C.c_double(defval), cap,
                          perms.ctypes.data_as(c_int_p), _p(values))
    return perms[:cnt], values[:cnt]


def murty_children(assignment, forced, eliminated):
    a, ap = _i(assignment)
    f, fp = _i(np.asarray(forced, dtype=np.int32).reshape(-1))
    e, ep = _i(np.asarray(eliminated, dtype=np.int32).reshape(-1))
    out = np.zeros(4096, dtype=np.int32)
    cnt = lib().orc_murty_children(len(a), ap, len(f) // 2, fp, len(e) // 2, ep, len(out),
                                   out.ctypes.data_as(c_int_p))
    res, pos = [], 0
    for _ in range(cnt):
        nf, ne = out[pos], out[pos + 1]
        pos += 2
        fo = [tuple(out[pos + 2 * i: pos + 2 * i + 2]) for i in range(nf)]
        pos += 2 * nf
        el = [tuple(out[pos + 2 * i: pos + 2 * i + 2]) for i in range(ne)]
        pos += 2 * ne
        res.append((fo, el))
    return res


# ---------------------------------------------------------------- navigator
class Navigator:
    """Whole-filter oracle (PHD:192-362)."""

    def __init__(self, cfg, P, pose, only_mapping=False):
        self.cfg = cfg
        self._h = C.c_void_p(lib().orc_nav_new(C.byref(cfg), int(P), _pose(pose)[1], int(only_mapping)))
        self.P = lib().orc_nav_particle_count(self._h)

    def close(self):
        if self._h:
            lib().orc_nav_delete(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_map(self, i, w, m, P):
        w, m, P = _map_in(w, m, P)
        lib().orc_nav_set_map(self._h, int(i), len(w), _p(w), _p(m), _p(P))

    def get_map(self, i, cap=1 << 16):
        ow, om, oP = _map_out(cap)
        n = lib().orc_nav_get_map(self._h, int(i), cap, _p(ow), _p(om), _p(oP))
        if n > cap:
            return self.get_map(i, n)
        return ow[:n].copy(), om[:n].copy(), oP[:n].copy()

    def set_pose(self, i, pose):
        lib().orc_nav_set_pose(self._h, int(i), _pose(pose)[1])

    def get_poses(self):
        out = np.zeros((self.P, 7))
        lib().orc_nav_get_poses(self._h, _p(out))
        return out

    def set_weights(self, w):
        w, wp = _d(w)
        lib().orc_nav_set_weights(self._h, wp)

    def get_weights(self):
        out = np.zeros(self.P)
        lib().orc_nav_get_weights(self._h, _p(out))
        return out

    def get_alphas(self):
        out = np.zeros(self.P)
        lib().orc_nav_get_alphas(self._h, _p(out))
        return out

    def update(self, reading, dt, gauss, perfect_still=False):
        g, gp = _d(np.asarray(gauss).reshape(self.P, 6))
        lib().orc_nav_update(self._h, _d(reading)[1], C.c_double(dt), gp, int(perfect_still))

    def slam_update(self, z, u):
        z = _z(z)
        best, res = C.c_int(), C.c_int()
        anc = np.zeros(self.P, dtype=np.int32)
        lib().orc_nav_slam_update(self._h, len(z), _p(z), C.c_double(u), C.byref(best), C.byref(res),
                                  anc.ctypes.data_as(c_int_p))
        return best.value, bool(res.value), anc

    def map_update_range(self, z, first, last):
        z = _z(z)
        lib().orc_nav_map_update_range(self._h, len(z), _p(z), int(first), int(last))
