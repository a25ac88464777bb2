"""Seeded synthetic 3-D pixel-range scenes.

Host-side workload generator only (numpy).  It produces the inputs a MonoRFS simulation would hand
to the navigator: landmarks, a steady-state Gaussian-mixture map, particle poses, and per frame the
odometry reading, the N(0,1) draws of TrackVehicle.UpdateNoisy (TRK:95-97), the measurement list
(SIMV:243-295: detections with probability PD*fuzzy plus uniform clutter) and the resampling
uniform (PHD:727).  The same arrays feed the CUDA engine, the oracle and the CPU baseline.
"""
import math
from dataclasses import dataclass, field

import numpy as np

SEED = 20261018

# workloads of BASELINE.json "configs" (P particles x N components x M measurements/frame)
WORKLOADS = {
    "c2": dict(P=2000, N=500, M=100),
    "c3": dict(P=200, N=50000, M=1000),
    "c4": dict(P=20000, N=2000, M=500),
    "tiny": dict(P=64, N=60, M=24),
    "c4s": dict(P=1480, N=2000, M=500),   # config 4's per-particle shape on 1480 particles (profiling runs)
    "c4m": dict(P=2960, N=2000, M=500),   # the same on 2960 particles (10 per CTA at two CTAs per SM)
    "c3l": dict(P=200, N=50000, M=1000),  # config 3 as Loopy's leave-one-out batch: 200 mapping-only filters that share
                                          # the trajectory pose of each frame, one of them skipping it (LOOPY:729-763)
    "c2x": dict(P=20000, N=500, M=100),   # config 2's per-particle shape on config 4's particle count: resamples
                                          # every frame, so it exercises the multi-GPU map migration
}

MEASURER = [575.8156, 0.1, 10.0, -320, -240, 640, 480]   # range clip widened to 10 m (section 8d)
R_DIAG = (2.0, 2.0, 1e-3)
Q_DIAG = (5e-3, 5e-3, 5e-3, 2e-4, 2e-4, 2e-4)
ODOMETRY = (0.0, 0.0, 0.01, 0.0, 0.002, 0.0)
DT = 1.0 / 30.0
BOX = ((-6.0, 6.0), (-4.5, 4.5), (0.0, 11.0))


def params(N, **over):
    """Navigator parameters of the synthetic configs (reference defaults, MaxQuantity = 2N)."""
    p = dict(
        model=0, max_quantity=2 * N, gate_metric=0, nthreads=8,
        R=np.diag(R_DIAG), Q=np.diag(Q_DIAG), pd=0.9, clutter=3e-7,
        birth_cov=np.eye(3) * 1e-2, birth_weight=0.05, min_weight=1e-3, merge_threshold=0.3,
        exploration_threshold=1e-5, density_distance_threshold=0.5, min_effective_particle=0.1,
        visibility_ramp=[3 * math.sqrt(r) for r in R_DIAG], measurer=list(MEASURER),
    )
    p.update(over)
    return p


# --------------------------------------------------------------------------- geometry (inputs only)
def qmul(a, b):
    aw, ax, ay, az = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bw, bx, by, bz = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([aw * bw - (ax * bx + ay * by + az * bz),
                     aw * bx + ax * bw + ay * bz - az * by,
                     aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx], axis=-1)


def qconj(q):
    return q * np.array([1.0, -1.0, -1.0, -1.0])


def qexp(lie):
    lie = np.asarray(lie, float)
    phi = np.linalg.norm(lie, axis=-1, keepdims=True)
    safe = np.where(phi < 1e-12, 1.0, phi)
    q = np.concatenate([np.cos(phi), np.sin(phi) * lie / safe], axis=-1)
    ident = np.zeros_like(q)
    ident[..., 0] = 1.0
    return np.where(phi < 1e-12, ident, q)


def add_odometry(pose, delta):
    """POSE:314-333 (vectorised; used for the true trajectory and the initial particle spread)."""
    pose = np.asarray(pose, float)
    delta = np.asarray(delta, float)
    q = pose[..., 3:7]
    dq = qexp(0.5 * delta[..., 3:6])
    newq = qmul(q, dq)
    rw = np.sqrt(0.5 * (1 + dq[..., 0:1]))
    mid = np.concatenate([rw, dq[..., 1:4] / (2 * rw)], axis=-1)
    midr = qmul(q, mid)
    v = np.concatenate([np.zeros_like(delta[..., 0:1]), delta[..., 0:3]], axis=-1)
    dl = qmul(qmul(midr, v), qconj(midr))
    newq = newq / np.linalg.norm(newq, axis=-1, keepdims=True)
    return np.concatenate([pose[..., 0:3] + dl[..., 1:4], newq], axis=-1)


def measure_perfect(pose, lm, focal):
    """PRM:138-149, vectorised over landmarks."""
    q = np.asarray(pose[3:7], float)
    diff = np.asarray(lm, float) - np.asarray(pose[0:3], float)
    v = np.concatenate([np.zeros((len(diff), 1)), diff], axis=1)
    loc = qmul(qmul(qconj(q)[None, :], v), q[None, :])[:, 1:4]
    rng = np.sign(loc[:, 2]) * np.linalg.norm(diff, axis=1)
    with np.errstate(divide="ignore", invalid="ignore"):
        px = focal * loc[:, 0] / loc[:, 2]
        py = focal * loc[:, 1] / loc[:, 2]
    return np.stack([px, py, rng], axis=1)


def fuzzy_visible(z, measurer, ramp):
    """PRM:277-291."""
    f, rmin, rmax, fx, fy, fw, fh = measurer
    rmin, rmax = float(np.float32(rmin)), float(np.float32(rmax))
    d = np.stack([(z[:, 0] - fx) / ramp[0], (fx + fw - z[:, 0]) / ramp[0],
                  (z[:, 1] - fy) / ramp[1], (fy + fh - z[:, 1]) / ramp[1],
                  (z[:, 2] - rmin) / ramp[2], (rmax - z[:, 2]) / ramp[2]], axis=1)
    d = np.where(np.isnan(d), -np.inf, d)
    return np.clip(d.min(axis=1), 0.0, 1.0)


def random_rotations(rng, n):
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q.T
    return np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                     2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                     2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], axis=1).reshape(n, 3, 3)


@dataclass
class Frame:
    reading: np.ndarray      # (6,)  odometry
    gauss: np.ndarray        # (P,6) N(0,1) draws for the particle noise
    z: np.ndarray            # (M,3) measurements (px, py, range)
    u: float                 # resampling uniform (float-valued, like AForge's generator)
    true_pose: np.ndarray    # (7,)


@dataclass
class Scene:
    P: int
    N: int
    M: int
    params: dict
    landmarks: np.ndarray                 # (N,3)
    map_w: np.ndarray                     # (N,)   steady-state map shared by all particles at t0
    map_m: np.ndarray                     # (N,3)
    map_P: np.ndarray                     # (N,3,3)
    poses: np.ndarray                     # (P,7)  initial particle poses
    true_pose: np.ndarray                 # (7,)
    rng: np.random.Generator = field(repr=False, default=None)
    box_scale: float = 1.0

    def next_frame(self):
        """Advance the true vehicle and draw one frame of navigator inputs."""
        p = self.params
        self.true_pose = add_odometry(self.true_pose, np.array(ODOMETRY))
        gauss = self.rng.normal(size=(self.P, 6)).astype(np.float32).astype(np.float64)
        meas, ramp = p["measurer"], p["visibility_ramp"]
        zp = measure_perfect(self.true_pose, self.landmarks, meas[0])
        pdet = p["pd"] * fuzzy_visible(zp, meas, ramp)
        hit = self.rng.random(len(zp)) < pdet
        rdiag = np.sqrt(np.diag(np.asarray(p["R"]).reshape(3, 3)))
        det = zp[hit] + self.rng.normal(size=(int(hit.sum()), 3)) * rdiag
        if len(det) > self.M:
            det = det[self.rng.permutation(len(det))[: self.M]]
        nclutter = self.M - len(det)
        rmin, rmax = float(np.float32(meas[1])), float(np.float32(meas[2]))
        clutter = np.stack([self.rng.random(nclutter) * meas[5] + meas[3],
                            self.rng.random(nclutter) * meas[6] + meas[4],
                            self.rng.random(nclutter) * (rmax - rmin) + rmin], axis=1)
        z = np.concatenate([det, clutter], axis=0)
        z = z[self.rng.permutation(len(z))]
        u = float(np.float32(self.rng.random()))
        return Frame(np.array(ODOMETRY), gauss, np.ascontiguousarray(z), u, self.true_pose.copy())


def make_scene(P, N, M, seed=SEED, box_scale=1.0, **over):
    """Steady-state start of section 8d: every particle carries the same N-component map
    (mean = landmark + N(0,1e-4 I), cov = rotated diag(U[0.5,2]*1e-3), weight U[0.3,1]) -- the state
    PHDNavigator.CollapseParticles (PHD:233-266) produces -- and its own pose = truth (+) N(0, Q dt^2)."""
    rng = np.random.default_rng(seed)
    p = params(N, **over)
    s = box_scale ** (1.0 / 3.0)
    lo = np.array([b[0] for b in BOX]) * np.array([s, s, 1.0])
    hi = np.array([b[1] for b in BOX]) * np.array([s, s, 1.0])
    hi[2] = BOX[2][1] * s
    if box_scale != 1.0:
        p["measurer"][2] = MEASURER[2] * s
    landmarks = rng.random((N, 3)) * (hi - lo) + lo
    map_m = landmarks + rng.normal(size=(N, 3)) * 1e-2
    rot = random_rotations(rng, N)
    d = rng.uniform(0.5, 2.0, size=(N, 3)) * 1e-3
    map_P = np.einsum("nij,nj,nkj->nik", rot, d, rot)
    map_P = 0.5 * (map_P + np.transpose(map_P, (0, 2, 1)))
    map_w = rng.uniform(0.3, 1.0, size=N)
    true_pose = np.array([0, 0, 0, 1, 0, 0, 0], float)
    spread = rng.normal(size=(P, 6)) * np.sqrt(np.array(Q_DIAG)) * DT
    poses = add_odometry(np.broadcast_to(true_pose, (P, 7)), spread)
    return Scene(P, N, M, p, landmarks, map_w, map_m, np.ascontiguousarray(map_P), np.ascontiguousarray(poses),
                 true_pose, rng, box_scale)


def make_workload(name, seed=SEED, **over):
    w = dict(WORKLOADS[name])
    w.update({k: over.pop(k) for k in ("P", "N", "M") if k in over})
    scale = 25.0 if name in ("c3", "c3l") else 1.0
    return make_scene(w["P"], w["N"], w["M"], seed=seed, box_scale=over.pop("box_scale", scale), **over)
