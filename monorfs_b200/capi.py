"""ctypes binding of librbphd.so -- the same C ABI the C# GpuPHDNavigator P/Invokes (include/rbphd.h).

There is no CPU fallback: loading fails loudly when the library has not been built, and every
computing call returns an error when no sm_100a device is present.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# RBPHD_LIB selects an experiment build of the same library (monorfs_b200.build.build_variant)
LIB_PATH = os.environ.get("RBPHD_LIB") or os.path.join(_HERE, "_build", "librbphd.so")

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)

OK, ERR_GENERIC, ERR_CUDA, ERR_CAPACITY, ERR_ARGUMENT, ERR_NO_DEVICE = 0, -1, 1, 2, 3, 4


class RbphdConfig(C.Structure):
    _fields_ = [
        ("model", C.c_int32), ("max_quantity", C.c_int32), ("gate_metric", C.c_int32), ("nthreads", C.c_int32),
        ("R", C.c_double * 9), ("Q", C.c_double * 36), ("pd", C.c_double), ("clutter", C.c_double),
        ("birth_cov", C.c_double * 9), ("birth_weight", C.c_double), ("min_weight", C.c_double),
        ("merge_threshold", C.c_double), ("exploration_threshold", C.c_double),
        ("density_distance_threshold", C.c_double), ("min_effective_particle", C.c_double),
        ("visibility_ramp", C.c_double * 3), ("measurer", C.c_double * 7),
    ]


class RbphdLimits(C.Structure):
    _fields_ = [
        ("device", C.c_int32), ("max_particles", C.c_int32), ("max_components", C.c_int32),
        ("max_measurements", C.c_int32), ("max_pairs", C.c_int32), ("resident_frames", C.c_int32),
        ("reserved", C.c_int32 * 2),
    ]


# every symbol include/rbphd.h declares (tests/test_capi_exports.py checks the header against this list)
EXPORTS = [
    "rbphd_new", "rbphd_delete", "rbphd_last_error", "rbphd_reset", "rbphd_clear_maps", "rbphd_particle_count",
    "rbphd_update", "rbphd_set_pose", "rbphd_set_poses", "rbphd_get_poses", "rbphd_slam_update",
    "rbphd_frame_async", "rbphd_upload_frame_inputs", "rbphd_update_async", "rbphd_synchronize", "rbphd_resample",
    "rbphd_particle_depleted", "rbphd_get_weights", "rbphd_set_weights", "rbphd_get_alphas", "rbphd_get_best",
    "rbphd_get_ancestors", "rbphd_get_map_counts", "rbphd_get_map", "rbphd_set_map", "rbphd_stage_predict",
    "rbphd_stage_correct", "rbphd_stage_prune", "rbphd_stage_weight_alpha", "rbphd_stage_set_loglikelihood",
    "rbphd_comm_unique_id", "rbphd_comm_init_rank", "rbphd_comm_destroy", "rbphd_frame_result", "rbphd_comm_stats",
    "rbphd_debug_migration_plan", "rbphd_kernel_launches", "rbphd_profile_enable",
    "rbphd_profile_read", "rbphd_get_counters", "rbphd_get_phase_cycles", "rbphd_stream",
    "rbphd_launch_shape", "rbphd_bench_fp64", "rbphd_slam_update_begin", "rbphd_slam_update_finish",
    "rbphd_set_likelihood", "rbphd_quasi_set_loglikelihood", "rbphd_set_loglike_matrix", "rbphd_set_holdout", "rbphd_set_depth_frame",
    "rbphd_quasi_set_loglikelihood_gradient", "rbphd_generate_measurements", "rbphd_ospa",
]

_lib = None


class RbphdError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("librbphd error %d: %s" % (code, msg))
        self.code = code


def load():
    """dlopen librbphd.so; raises if it has not been built (python -m monorfs_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("librbphd.so is not built (run `python -m monorfs_b200.build`); "
                               "monorfs_b200 has no CPU fallback")
        lib = C.CDLL(LIB_PATH)
        lib.rbphd_new.restype = C.c_void_p
        lib.rbphd_new.argtypes = [C.POINTER(RbphdConfig), C.POINTER(RbphdLimits)]
        lib.rbphd_delete.argtypes = [C.c_void_p]
        lib.rbphd_delete.restype = None
        lib.rbphd_last_error.restype = C.c_char_p
        lib.rbphd_last_error.argtypes = [C.c_void_p]
        lib.rbphd_kernel_launches.restype = C.c_int64
        lib.rbphd_kernel_launches.argtypes = [C.c_void_p]
        lib.rbphd_stream.restype = C.c_void_p
        lib.rbphd_stream.argtypes = [C.c_void_p]
        _lib = lib
    return _lib


def comm_unique_id():
    """A fresh NCCL unique id (128 bytes) for rbphd_comm_init_rank; call on one rank and distribute."""
    lib = load()
    buf = (C.c_ubyte * 128)()
    rc = lib.rbphd_comm_unique_id(buf)
    if rc != OK:
        raise RbphdError(rc, lib.rbphd_last_error(None).decode())
    return bytes(buf)


def debug_migration_plan(ancestors, counts, world, rank, device=0):
    """The exchange plan the device derives for `rank` (test hook, rbphd_debug_migration_plan)."""
    lib = load()
    anc = np.ascontiguousarray(ancestors, dtype=np.int32)
    cnt = np.ascontiguousarray(counts, dtype=np.int32)
    total = len(anc)
    lo, hi = (total * rank) // world, (total * (rank + 1)) // world
    pl = hi - lo
    local_src = np.zeros(max(pl, 1), np.int32)
    rec_off = np.zeros(max(pl, 1), np.int64)
    send_idx = np.zeros(pl + world, np.int32)
    send_off = np.zeros(pl + world, np.int64)
    hdr = np.zeros(2 * world + 3, np.int64)
    ip = lambda a: a.ctypes.data_as(c_int_p)
    lp = lambda a: a.ctypes.data_as(C.POINTER(C.c_int64))
    rc = lib.rbphd_debug_migration_plan(int(device), ip(anc), ip(cnt), total, int(world), int(rank), ip(local_src),
                                        lp(rec_off), ip(send_idx), lp(send_off), lp(hdr))
    if rc != OK:
        raise RbphdError(rc, lib.rbphd_last_error(None).decode())
    nsend = int(hdr[2 * world])
    return dict(local_src=local_src[:pl], rec_off=rec_off[:pl], send_idx=send_idx[:nsend], send_off=send_off[:nsend],
                send_doubles=hdr[:world].copy(), recv_doubles=hdr[world:2 * world].copy(), n_send=nsend,
                n_recv=int(hdr[2 * world + 1]), sorted=bool(hdr[2 * world + 2]))


def bench_fp64(device=0, outer=600):
    """FP64 pipe microbenchmark (rbphd_microbench.cu): dict of measured peaks on `device`."""
    lib = load()
    out = (C.c_double * 6)()
    rc = lib.rbphd_bench_fp64(int(device), int(outer), out)
    if rc != OK:
        raise RbphdError(rc, "rbphd_bench_fp64 failed")
    keys = ("dfma_tflops", "dmul_dadd_tflops", "dadd_tflops", "fp64_tinst_per_s_fused", "fp64_tinst_per_s_unfused", "sms")
    return dict(zip(keys, [float(x) for x in out]))


def make_config(p):
    c = RbphdConfig()
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
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(c_double_p)


def _view(ptr, n, dtype=np.float64):
    if n <= 0:
        return np.zeros(0, dtype=dtype)
    ct = C.c_double if dtype == np.float64 else C.c_int
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(n,)).copy()


class Handle:
    """Thin RAII wrapper over rbphd_navigator* (mirrors the HandleRef + Dispose pattern of ISAM2Navigator.cs:446-452)."""

    def __init__(self, params, max_particles, max_components=0, max_measurements=0, max_pairs=0, device=0,
                 resident_frames=1):
        self.lib = load()
        self.cfg = make_config(params)
        lim = RbphdLimits()
        lim.device = device
        lim.max_particles = int(max_particles)
        lim.max_components = int(max_components)
        lim.max_measurements = int(max_measurements)
        lim.max_pairs = int(max_pairs)
        lim.resident_frames = int(resident_frames)
        self._h = self.lib.rbphd_new(C.byref(self.cfg), C.byref(lim))
        if not self._h:
            raise RbphdError(ERR_NO_DEVICE, self.lib.rbphd_last_error(None).decode())
        self._h = C.c_void_p(self._h)

    def close(self):
        if getattr(self, "_h", None):
            self.lib.rbphd_delete(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, code):
        if code != OK:
            raise RbphdError(code, self.lib.rbphd_last_error(self._h).decode())

    # ---- state
    def reset(self, particles, pose, w, m, P):
        w, m, P = _d(w).reshape(-1), _d(m).reshape(-1, 3), _d(P).reshape(-1, 9)
        self._ck(self.lib.rbphd_reset(self._h, int(particles), _p(_d(pose)), len(w), _p(w), _p(m), _p(P)))

    @property
    def particles(self):
        return self.lib.rbphd_particle_count(self._h)

    def set_pose(self, i, pose):
        self._ck(self.lib.rbphd_set_pose(self._h, int(i), _p(_d(pose))))

    def set_poses(self, poses):
        poses = _d(poses).reshape(-1, 7)
        assert len(poses) == self.particles
        self._ck(self.lib.rbphd_set_poses(self._h, _p(poses)))

    def get_poses(self):
        ptr, n = c_double_p(), C.c_int()
        self._ck(self.lib.rbphd_get_poses(self._h, C.byref(ptr), C.byref(n)))
        return _view(ptr, 7 * n.value).reshape(-1, 7)

    def get_weights(self):
        ptr, n = c_double_p(), C.c_int()
        self._ck(self.lib.rbphd_get_weights(self._h, C.byref(ptr), C.byref(n)))
        return _view(ptr, n.value)

    def get_alphas(self):
        ptr, n = c_double_p(), C.c_int()
        self._ck(self.lib.rbphd_get_alphas(self._h, C.byref(ptr), C.byref(n)))
        return _view(ptr, n.value)

    def set_weights(self, w):
        w = _d(w)
        assert len(w) == self.particles
        self._ck(self.lib.rbphd_set_weights(self._h, _p(w)))

    def get_best(self):
        b = C.c_int()
        self._ck(self.lib.rbphd_get_best(self._h, C.byref(b)))
        return b.value

    def get_ancestors(self):
        ptr, n = c_int_p(), C.c_int()
        self._ck(self.lib.rbphd_get_ancestors(self._h, C.byref(ptr), C.byref(n)))
        return _view(ptr, n.value, np.int32)

    def get_map_counts(self):
        ptr, n = c_int_p(), C.c_int()
        self._ck(self.lib.rbphd_get_map_counts(self._h, C.byref(ptr), C.byref(n)))
        return _view(ptr, n.value, np.int32)

    def _maps_out(self, pw, pm, pP, n):
        n = n.value
        return _view(pw, n), _view(pm, 3 * n).reshape(-1, 3), _view(pP, 9 * n).reshape(-1, 3, 3)

    def get_map(self, i):
        pw, pm, pP, n = c_double_p(), c_double_p(), c_double_p(), C.c_int()
        self._ck(self.lib.rbphd_get_map(self._h, int(i), C.byref(pw), C.byref(pm), C.byref(pP), C.byref(n)))
        return self._maps_out(pw, pm, pP, n)

    def set_map(self, i, w, m, P):
        w, m, P = _d(w).reshape(-1), _d(m).reshape(-1, 3), _d(P).reshape(-1, 9)
        self._ck(self.lib.rbphd_set_map(self._h, int(i), len(w), _p(w), _p(m), _p(P)))

    def clear_maps(self):
        self._ck(self.lib.rbphd_clear_maps(self._h))

    # ---- frame
    def update(self, reading, dt, gauss, perfect_still=False):
        g = _d(gauss).reshape(-1, 6)
        assert len(g) == self.particles
        self._ck(self.lib.rbphd_update(self._h, _p(_d(reading)), C.c_double(dt), _p(g), int(perfect_still)))

    def slam_update(self, z, u, only_mapping=False):
        z = _d(z).reshape(-1, 3)
        best, res = C.c_int(), C.c_int()
        self._ck(self.lib.rbphd_slam_update(self._h, _p(z), len(z), int(only_mapping), C.c_double(u),
                                            C.byref(best), C.byref(res)))
        return best.value, bool(res.value)

    def set_depth_frame(self, depth_xy):
        """Attach a Kinect depth frame, depth_xy[x, y] in metres (float32, NaN = no reading); None detaches it."""
        if depth_xy is None:
            self._ck(self.lib.rbphd_set_depth_frame(self._h, None, 0, 0))
            return
        d = np.ascontiguousarray(depth_xy, dtype=np.float32)
        self._ck(self.lib.rbphd_set_depth_frame(self._h, d.ctypes.data_as(C.POINTER(C.c_float)), d.shape[0], d.shape[1]))

    def set_holdout(self, particle):
        self._ck(self.lib.rbphd_set_holdout(self._h, int(particle)))

    def filter_missing_batch(self, trajectory, factors, to=None):
        """All leave-one-out maps of LoopyPHDNavigator.FilterMissing (LOOPY:729-763) in one pass: particle j ends
        with the map filtered over frames 0..to-1 except frame j (j >= to: nothing skipped).  The navigator must
        hold len(trajectory) particles with empty (or common prior) maps."""
        T = len(trajectory)
        to = T if to is None else min(T, to)
        assert self.particles >= T
        for i in range(to):
            self.set_poses(np.tile(_d(trajectory[i]).reshape(1, 7), (self.particles, 1)))
            self.set_holdout(i)
            z = _d(factors[i]).reshape(-1, 3)
            self.upload_frame_inputs(None, z, slot=0)
            self.frame_async(None, 0.0, len(z), 0.0, only_mapping=True, slot=0)
        self.set_holdout(-1)
        self.synchronize()

    def slam_update_begin(self, z, only_mapping=False):
        z = _d(z).reshape(-1, 3)
        best, dep = C.c_int(), C.c_int()
        self._ck(self.lib.rbphd_slam_update_begin(self._h, _p(z), len(z), int(only_mapping), C.byref(best), C.byref(dep)))
        return best.value, bool(dep.value)

    def slam_update_finish(self, u):
        best = C.c_int()
        self._ck(self.lib.rbphd_slam_update_finish(self._h, C.c_double(u), C.byref(best)))
        return best.value

    def upload_frame_inputs(self, gauss, z, slot=0):
        g = _d(gauss).reshape(-1, 6) if gauss is not None else None
        zz = _d(z).reshape(-1, 3) if z is not None else None
        self._ck(self.lib.rbphd_upload_frame_inputs(self._h, int(slot), _p(g) if g is not None else None,
                                                    _p(zz) if zz is not None else None,
                                                    len(zz) if zz is not None else 0))

    def frame_async(self, reading, dt, m, u, only_mapping=False, perfect_still=False, slot=0):
        self._ck(self.lib.rbphd_frame_async(self._h, int(slot), _p(_d(reading)) if reading is not None else None,
                                            C.c_double(dt), int(perfect_still), int(m), int(only_mapping),
                                            C.c_double(u)))

    def update_async(self, reading, dt, slot=0, perfect_still=False):
        self._ck(self.lib.rbphd_update_async(self._h, int(slot), _p(_d(reading)), C.c_double(dt), int(perfect_still)))

    def synchronize(self):
        self._ck(self.lib.rbphd_synchronize(self._h))

    def resample(self, u):
        self._ck(self.lib.rbphd_resample(self._h, C.c_double(u)))

    def particle_depleted(self):
        d = C.c_int()
        self._ck(self.lib.rbphd_particle_depleted(self._h, C.byref(d)))
        return bool(d.value)

    # ---- stages
    def stage_predict(self, pose, w, m, P, z):
        w, m, P, z = _d(w).reshape(-1), _d(m).reshape(-1, 3), _d(P).reshape(-1, 9), _d(z).reshape(-1, 3)
        pw, pm, pP, n = c_double_p(), c_double_p(), c_double_p(), C.c_int()
        self._ck(self.lib.rbphd_stage_predict(self._h, _p(_d(pose)), len(w), _p(w), _p(m), _p(P), _p(z), len(z),
                                              C.byref(pw), C.byref(pm), C.byref(pP), C.byref(n)))
        return self._maps_out(pw, pm, pP, n)

    def stage_correct(self, pose, w, m, P, z, gate_radius=None):
        w, m, P, z = _d(w).reshape(-1), _d(m).reshape(-1, 3), _d(P).reshape(-1, 9), _d(z).reshape(-1, 3)
        if gate_radius is None:
            gate_radius = self.cfg.density_distance_threshold
        pw, pm, pP, n = c_double_p(), c_double_p(), c_double_p(), C.c_int()
        self._ck(self.lib.rbphd_stage_correct(self._h, _p(_d(pose)), len(w), _p(w), _p(m), _p(P), _p(z), len(z),
                                              C.c_double(gate_radius), C.byref(pw), C.byref(pm), C.byref(pP),
                                              C.byref(n)))
        return self._maps_out(pw, pm, pP, n)

    def stage_prune(self, w, m, P):
        w, m, P = _d(w).reshape(-1), _d(m).reshape(-1, 3), _d(P).reshape(-1, 9)
        pw, pm, pP, n = c_double_p(), c_double_p(), c_double_p(), C.c_int()
        self._ck(self.lib.rbphd_stage_prune(self._h, len(w), _p(w), _p(m), _p(P), C.byref(pw), C.byref(pm),
                                            C.byref(pP), C.byref(n)))
        return self._maps_out(pw, pm, pP, n)

    def stage_weight_alpha(self, pose, z, pred, corr):
        pw, pm, pP = _d(pred[0]).reshape(-1), _d(pred[1]).reshape(-1, 3), _d(pred[2]).reshape(-1, 9)
        cw, cm, cP = _d(corr[0]).reshape(-1), _d(corr[1]).reshape(-1, 3), _d(corr[2]).reshape(-1, 9)
        z = _d(z).reshape(-1, 3)
        out = np.zeros(7)
        self._ck(self.lib.rbphd_stage_weight_alpha(self._h, _p(_d(pose)), _p(z), len(z), len(pw), _p(pw), _p(pm),
                                                   _p(pP), len(cw), _p(cw), _p(cm), _p(cP), _p(out)))
        return dict(alpha=out[0], setloglik=out[1], ploglik=out[2], cloglik=out[3], pcount=out[4],
                    ccount=out[5], J=int(out[6]))

    def stage_set_loglikelihood(self, pose, jm, z):
        jm, z = _d(jm).reshape(-1, 3), _d(z).reshape(-1, 3)
        out = C.c_double()
        self._ck(self.lib.rbphd_stage_set_loglikelihood(self._h, _p(_d(pose)), len(jm), _p(jm), _p(z), len(z),
                                                        C.byref(out)))
        return out.value

    def set_likelihood(self, pose, jm, z):
        jm, z = _d(jm).reshape(-1, 3), _d(z).reshape(-1, 3)
        out = C.c_double()
        self._ck(self.lib.rbphd_set_likelihood(self._h, _p(_d(pose)), len(jm), _p(jm), _p(z), len(z), C.byref(out)))
        return out.value

    def quasi_set_loglikelihood(self, pose, jm, z):
        jm, z = _d(jm).reshape(-1, 3), _d(z).reshape(-1, 3)
        out = C.c_double()
        self._ck(self.lib.rbphd_quasi_set_loglikelihood(self._h, _p(_d(pose)), len(jm), _p(jm), _p(z), len(z),
                                                        C.byref(out)))
        return out.value

    def quasi_set_loglikelihood_gradient(self, pose, jm, z, sum_normalised=False):
        """(value, gradient[6]) of QuasiSetLogLikelihood(..., out gradient) (PHD:544-549)."""
        jm, z = _d(jm).reshape(-1, 3), _d(z).reshape(-1, 3)
        out, g = C.c_double(), np.zeros(6)
        self._ck(self.lib.rbphd_quasi_set_loglikelihood_gradient(self._h, _p(_d(pose)), len(jm), _p(jm), _p(z), len(z),
                                                                 int(sum_normalised), C.byref(out), _p(g)))
        return out.value, g

    def generate_measurements(self, pose, landmarks, uniforms, gauss, clutter_u, chol=None):
        """SimulatedVehicle.Measure (SIMV:243-295) with the host's random numbers -> (z, assoc)."""
        lm, ga, cu = _d(landmarks).reshape(-1, 3), _d(gauss).reshape(-1, 3), _d(clutter_u).reshape(-1, 3)
        un = _d(uniforms).reshape(-1)
        if len(un) != len(lm) or len(ga) != len(lm):
            raise ValueError("one uniform and three gaussian draws per landmark")
        z = np.zeros((len(lm) + len(cu), 3))
        assoc = np.zeros(len(lm) + len(cu), dtype=np.int32)
        n = C.c_int()
        ch = _p(_d(chol).reshape(9)) if chol is not None else None
        self._ck(self.lib.rbphd_generate_measurements(self._h, _p(_d(pose)), _p(lm), len(lm), _p(un), _p(ga), ch, _p(cu),
                                                      len(cu), _p(z), assoc.ctypes.data_as(c_int_p), C.byref(n)))
        return z[:n.value].copy(), assoc[:n.value].copy()

    def ospa(self, a, b, c=1.0, p=2.0):
        """(OSPA, cardinality error) between two landmark position sets (postanalysis/Plot.cs:531-581)."""
        a, b = _d(a).reshape(-1, 3), _d(b).reshape(-1, 3)
        out, card = C.c_double(), C.c_double()
        self._ck(self.lib.rbphd_ospa(self._h, _p(a), len(a), _p(b), len(b), C.c_double(c), C.c_double(p), C.byref(out),
                                     C.byref(card)))
        return out.value, card.value

    def set_loglike_matrix(self, pose, jm, z):
        """SetLogLikeMatrix (PHD:415-460) as sorted (row, col, value) triplets."""
        jm, z = _d(jm).reshape(-1, 3), _d(z).reshape(-1, 3)
        rows, cols, vals, n = c_int_p(), c_int_p(), c_double_p(), C.c_int()
        self._ck(self.lib.rbphd_set_loglike_matrix(self._h, _p(_d(pose)), len(jm), _p(jm), _p(z), len(z),
                                                   C.byref(rows), C.byref(cols), C.byref(vals), C.byref(n)))
        return _view(rows, n.value, np.int32), _view(cols, n.value, np.int32), _view(vals, n.value)

    # ---- multi-GPU (collectives inside the library)
    def comm_init(self, unique_id, rank, world, total_particles):
        """Join the NCCL communicator (unique_id: 128 bytes from comm_unique_id() on one rank)."""
        buf = (C.c_ubyte * 128).from_buffer_copy(bytes(unique_id))
        self._ck(self.lib.rbphd_comm_init_rank(self._h, buf, int(rank), int(world), int(total_particles)))

    def comm_destroy(self):
        self._ck(self.lib.rbphd_comm_destroy(self._h))

    def comm_stats(self):
        out = (C.c_int64 * 4)()
        self._ck(self.lib.rbphd_comm_stats(self._h, out))
        return dict(resampling_frames=int(out[0]), sent_bytes=int(out[1]), recv_bytes=int(out[2]), records=int(out[3]))

    def frame_result(self):
        best, res = C.c_int(), C.c_int()
        self._ck(self.lib.rbphd_frame_result(self._h, C.byref(best), C.byref(res)))
        return best.value, bool(res.value)

    # ---- instrumentation
    STAGES = ("pose", "prep", "particle_update", "normalize_resample", "copy_particles")

    def profile_enable(self, max_frames):
        self._ck(self.lib.rbphd_profile_enable(self._h, int(max_frames)))

    def profile_read(self, max_frames):
        ms = np.zeros((max_frames, 5))
        n = C.c_int()
        self._ck(self.lib.rbphd_profile_read(self._h, _p(ms), int(max_frames), C.byref(n)))
        return ms[:n.value]

    def counters(self, reset=False):
        out = (C.c_int64 * 4)()
        self._ck(self.lib.rbphd_get_counters(self._h, out, int(reset)))
        return dict(comps_in=out[0], comps_out=out[1], pairs=out[2], particle_frames=out[3])

    PHASES = ("A1 meas->map", "A2 prior comps+gate", "A3 prior pairs", "A4-5 explore+births", "A6-7 birth comps+pairs",
              "A8-9 pair sort+weights", "B1 candidate sort", "B2 materialise", "B3 grid+edges", "B4-5 resolve",
              "B6 merge+write", "C1-3 map estimate", "C4 eval predicted", "C4 eval corrected", "C5 set likelihood",
              "tail", "A4a undecided list", "A4b density accumulate", "A4c decide", "B3a merge grid build",
              "C4a zero", "C4b accumulate", "C4c log-sum", "C3b query grid build", "C5a gate edges", "C5b components",
              "C5c edge sort", "A4b1 gather points", "A4b2 mini grid build", "C4b enumerate", "C4b dense process", "A3a gated component update",
              "B3b cell-ordered copies", "B3c exact edge tests", "B3d edge sort", "B6a survivor scan", "B1a candidate compaction",
              "C1 expected size", "C2 multiset expand", "C3 multiset sort", "C5d block enumeration") + tuple("p%d" % i for i in range(41, 64))

    DEBUG_COUNTERS = ("undecided measurements", "explore hits", "gated components", "merge edges", "W0",
                      "candidates", "eval pairs", "wide components", "eval cell rows", "likelihood edges", "J",
                      "murty blocks", "d12", "d13", "d14", "d15")

    def phase_cycles(self):
        out = (C.c_int64 * 80)()
        self._ck(self.lib.rbphd_get_phase_cycles(self._h, out))
        return {k: int(out[i]) for i, k in enumerate(self.PHASES)}

    def debug_counters(self):
        out = (C.c_int64 * 80)()
        self._ck(self.lib.rbphd_get_phase_cycles(self._h, out))
        return {k: int(out[64 + i]) for i, k in enumerate(self.DEBUG_COUNTERS)}

    @property
    def kernel_launches(self):
        return int(self.lib.rbphd_kernel_launches(self._h))

    def launch_shape(self):
        out = (C.c_int64 * 5)()
        self._ck(self.lib.rbphd_launch_shape(self._h, out))
        return dict(zip(("block", "ctas_per_sm", "smem_bytes", "slabs", "slab_bytes"), [int(x) for x in out]))

    @property
    def stream(self):
        return self.lib.rbphd_stream(self._h)
