"""Headless simulation run and recorded-run replay around the GPU navigator (BASELINE.json configs 1 and 5).

`HeadlessRun` is the drive loop of `monorfs -i=simulation -f=map.world -c=movements.in -p=N -x`
(UI/Simulation.cs:560-683) without graphics: per command the true vehicle moves (Vehicle.Update,
Vehicle.cs:325-336), its corrupted odometry is read (ReadOdometry, Vehicle.cs:342-353) and handed to
Navigator.Update, the vehicle measures (SimulatedVehicle.Measure, SimulatedVehicle.cs:243-295: detection with
probability PD * fuzzy visibility, measurement noise, Poisson clutter capped at 10 lambda) and the measurements go
to Navigator.SlamUpdate -- which here is librbphd.so through the C ABI.  Everything on this side of the ABI is
the host's business in the reference as well (its C#): the random streams, the histories (WayPoints, WayOdometry,
WayMeasurements, WayTrajectories, WayMaps; Navigator.cs:258-272) and the file formats (recordio).
`replay` feeds a recorded run (data.zip: odometry + measurements) through the navigator the way
RecordVehicle does for `-i=record` (RecordVehicle.cs:132-200).
"""
import time as _time

import numpy as np

from . import capi, recordio, synth


class TrueVehicle:
    """SimulatedVehicle: groundtruth pose, odometry corruption and the measurement process (host-side numpy)."""

    def __init__(self, pose, landmarks, params, rng):
        self.pose = np.asarray(pose, float).copy()
        self.odometry_pose = self.pose.copy()
        self.ref_odometry = self.pose.copy()
        self.landmarks = np.asarray(landmarks, float).reshape(-1, 3)
        self.p = params
        self.rng = rng
        q = np.asarray(params["Q"], float).reshape(6, 6)
        qq = q.copy()
        for i in range(6):
            qq[i, i] = max(qq[i, i], 1e-40)
        self.chol = np.linalg.cholesky(qq)
        meas = params["measurer"]
        self.range_length = float(np.float32(meas[2]) - np.float32(meas[1]))   # AForge.Range.Length: float arithmetic
        self.volume = meas[5] * meas[6] * self.range_length   # PRM:119-122
        self.clutter_count = params["clutter"] * self.volume

    def update(self, reading, dt):
        """Vehicle.Update (Vehicle.cs:325-336)."""
        self.pose = synth.add_odometry(self.pose, reading)
        self.odometry_pose = synth.add_odometry(self.odometry_pose, reading)
        noise = dt * (self.chol @ self.rng.normal(size=6).astype(np.float32).astype(np.float64))
        self.odometry_pose = synth.add_odometry(self.odometry_pose, noise)

    def read_odometry(self):
        """Vehicle.ReadOdometry (Vehicle.cs:342-353): odometry pose relative to the reference, then both reset."""
        reading = diff_odometry(self.odometry_pose, self.ref_odometry)
        self.odometry_pose = self.pose.copy()
        self.ref_odometry = self.pose.copy()
        return reading

    def measure_on_device(self, handle):
        """SimulatedVehicle.Measure with the measurement model evaluated by librbphd.so (rbphd_generate_measurements);
        the random numbers are drawn here, in the reference's order of kinds (detection, noise, count, clutter)."""
        p, rng = self.p, self.rng
        n = len(self.landmarks)
        un, ga = rng.random(n), rng.normal(size=(n, 3))
        ncl = min(int(rng.poisson(self.clutter_count)), int(self.clutter_count * 10)) if self.clutter_count > 0 else 0
        z, assoc = handle.generate_measurements(self.pose, self.landmarks, un, ga, rng.random((ncl, 3)))
        zp = synth.measure_perfect(self.pose, self.landmarks, p["measurer"][0]) if n else np.zeros((0, 3))
        pdet = p["pd"] * synth.fuzzy_visible(zp, p["measurer"], p["visibility_ramp"]) if n else np.zeros(0)
        hit = np.zeros(n, bool)
        hit[assoc[assoc >= 0]] = True
        return z, self.landmarks[pdet > 0], hit[pdet > 0]

    def measure(self):
        """SimulatedVehicle.Measure (SimulatedVehicle.cs:243-295)."""
        p, rng = self.p, self.rng
        meas, ramp = p["measurer"], p["visibility_ramp"]
        zp = synth.measure_perfect(self.pose, self.landmarks, meas[0]) if len(self.landmarks) else np.zeros((0, 3))
        pdet = p["pd"] * synth.fuzzy_visible(zp, meas, ramp) if len(zp) else np.zeros(0)
        hit = (pdet > 0) & (rng.random(len(zp)) < pdet)
        rdiag = np.sqrt(np.diag(np.asarray(p["R"]).reshape(3, 3)))
        det = zp[hit] + rng.normal(size=(int(hit.sum()), 3)) * rdiag
        ncl = min(int(rng.poisson(self.clutter_count)), int(self.clutter_count * 10)) if self.clutter_count > 0 else 0
        rmin, rmax = float(np.float32(meas[1])), float(np.float32(meas[2]))
        clutter = np.stack([rng.random(ncl) * meas[5] + meas[3], rng.random(ncl) * meas[6] + meas[4],
                            rng.random(ncl) * self.range_length + rmin], axis=1) if ncl else np.zeros((0, 3))
        visible = self.landmarks[pdet > 0]
        return np.ascontiguousarray(np.concatenate([det, clutter], axis=0)), visible, hit[pdet > 0]


def qlog(q):
    """Quaternion.Log (QUAT:155-183): rotation vector of a unit quaternion."""
    v = np.asarray(q[1:4], float)
    n = np.linalg.norm(v)
    if n < 1e-12:
        return np.zeros(3)
    return v / n * np.arctan2(n, q[0])


def diff_odometry(pose, origin):
    """Pose3D.DiffOdometry (POSE:336-350): the odometry that takes `origin` to `pose`."""
    q0, q1 = np.asarray(origin[3:7], float), np.asarray(pose[3:7], float)
    dq = synth.qmul(synth.qconj(q0), q1)
    dq = dq / np.linalg.norm(dq)
    rw = np.sqrt(0.5 * (1 + dq[0]))
    mid = synth.qmul(q0, np.concatenate([[rw], dq[1:4] / (2 * rw)]) if abs(dq[0] + 1) >= 1e-8 else np.array([1.0, 0, 0, 0]))
    d = np.asarray(pose[0:3], float) - np.asarray(origin[0:3], float)
    dl = synth.qmul(synth.qmul(synth.qconj(mid), np.concatenate([[0.0], d])), mid)[1:4]
    return np.concatenate([dl, 2 * qlog(dq)])


class HeadlessRun:
    """One `monorfs -i=simulation ... -x` run with the navigator on the GPU."""

    def __init__(self, pose0, measurer, landmarks, commands, particles, params=None, seed=synth.SEED, device=0,
                 max_components=0, max_measurements=0, device_measure=False):
        self.device_measure = bool(device_measure)
        self.commands = [np.asarray(c, float) for c in commands]
        self.P = int(particles)
        n_lm = max(1, len(landmarks))
        self.params = dict(params) if params is not None else synth.params(n_lm)
        self.params["measurer"] = [float(v) for v in measurer]
        self.rng = np.random.default_rng(seed)
        self.vehicle = TrueVehicle(pose0, landmarks, self.params, self.rng)
        self.pose0 = np.asarray(pose0, float)
        self.measurer = np.asarray(measurer, float)
        self.landmarks = np.asarray(landmarks, float).reshape(-1, 3)
        cap = max_components or max(64, 2 * self.params["max_quantity"])
        mcap = max_measurements or max(64, 4 * n_lm)
        self.h = capi.Handle(self.params, max_particles=self.P, max_components=cap, max_measurements=mcap,
                             max_pairs=16 * mcap, device=device)
        self.h.reset(self.P, self.pose0, np.zeros(0), np.zeros((0, 3)), np.zeros((0, 3, 3)))
        self.only_mapping = False
        self.rec = recordio.Recording(self.pose0, self.measurer, self.landmarks)
        self.way = [[] for _ in range(self.P)]   # per particle WayPoints (follow the ancestors on resampling)
        self.frames = 0
        self.gpu_seconds = 0.0
        self.resamples = 0

    def step(self, command, dt=synth.DT):
        t = (self.frames + 1) * dt
        rec, rng = self.rec, self.rng
        if len(command) > 6 and command[6] > 0:
            rec.tags.append((t, "SLAM mode on"))
            self.only_mapping = False
        elif len(command) > 6 and command[6] < 0:
            rec.tags.append((t, "Mapping mode on"))
            self.only_mapping = True
        self.vehicle.update(command[:6], dt)
        reading = self.vehicle.read_odometry()
        z, visible, detected = (self.vehicle.measure_on_device(self.h) if self.device_measure
                                else self.vehicle.measure())
        rec.trajectory.append((t, self.vehicle.pose.copy()))
        rec.odometry.append((t, reading))
        rec.measurements.append((t, z))
        rec.vismaps.append((t, (detected.astype(float), visible, np.tile(np.eye(3) * 1e-3, (len(visible), 1, 1)))))
        gauss = rng.normal(size=(self.P, 6)).astype(np.float32).astype(np.float64)
        u = float(np.float32(rng.random()))
        t0 = _time.perf_counter()
        if self.only_mapping:
            self.h.set_poses(np.tile(self.vehicle.pose, (self.P, 1)))       # PHD:297-300
        else:
            self.h.update(reading, dt, gauss)
        best, resampled = self.h.slam_update(z, u, only_mapping=self.only_mapping)
        poses = self.h.get_poses()
        bw, bm, bP = self.h.get_map(best)
        self.gpu_seconds += _time.perf_counter() - t0
        if resampled:
            self.resamples += 1
            anc = self.h.get_ancestors()
            self.way = [list(self.way[a]) for a in anc]
        for i in range(self.P):
            self.way[i].append((t, poses[i].copy()))
        rec.estimate.append((t, list(self.way[best])))
        rec.maps.append((t, (bw, bm, bP)))
        self.frames += 1
        return best, resampled

    def run(self):
        for c in self.commands:
            self.step(c)
        self.h.synchronize()
        return self.rec

    def close(self):
        self.h.close()


def replay(rec, particles, params=None, seed=synth.SEED, device=0, max_frames=None, handle=None):
    """Feed a recorded run's odometry and measurements through the navigator (`-i=record`).  Returns
    (estimate history of the best particle, final best map, seconds in the navigator, resampling frames)."""
    n_lm = max(1, len(rec.landmarks))
    prm = dict(params) if params is not None else synth.params(n_lm)
    prm["measurer"] = [float(v) for v in rec.measurer]
    rng = np.random.default_rng(seed)
    mmax = max([len(z) for _, z in rec.measurements] + [8])
    h = handle or capi.Handle(prm, max_particles=particles, max_components=max(64, 2 * prm["max_quantity"]),
                              max_measurements=mmax, max_pairs=16 * mmax, device=device)
    h.reset(particles, rec.pose0, np.zeros(0), np.zeros((0, 3)), np.zeros((0, 3, 3)))
    frames = list(zip(rec.odometry, rec.measurements))[:max_frames]
    way = [[] for _ in range(particles)]
    out, secs, nres, prev_t, best = [], 0.0, 0, 0.0, 0
    for (t, reading), (_, z) in frames:
        gauss = rng.normal(size=(particles, 6)).astype(np.float32).astype(np.float64)
        u = float(np.float32(rng.random()))
        t0 = _time.perf_counter()
        h.update(reading, t - prev_t, gauss)
        best, res = h.slam_update(z, u)
        poses = h.get_poses()
        secs += _time.perf_counter() - t0
        if res:
            nres += 1
            anc = h.get_ancestors()
            way = [list(way[a]) for a in anc]
        for i in range(particles):
            way[i].append((t, poses[i].copy()))
        out.append((t, list(way[best])))
        prev_t = t
    final_map = h.get_map(best)
    if handle is None:
        h.close()
    return out, final_map, secs, nres


def synthetic_scene(n_landmarks, seed=synth.SEED):
    """A map.world for config 1: landmarks uniform in the synthetic box, camera at the origin."""
    rng = np.random.default_rng(seed)
    lo = np.array([b[0] for b in synth.BOX])
    hi = np.array([b[1] for b in synth.BOX])
    landmarks = rng.random((n_landmarks, 3)) * (hi - lo) + lo
    return np.array([0, 0, 0, 1.0, 0, 0, 0]), np.array(synth.MEASURER, float), landmarks


def synthetic_commands(n_frames):
    """A movements.in: constant forward motion with a slow yaw; SLAM switched on by the first command."""
    cmds = []
    for f in range(n_frames):
        c = list(synth.ODOMETRY) + [1.0 if f == 0 else 0.0]
        cmds.append(np.array(c))
    return cmds


def best_map_estimate(w, m):
    """Map.BestMapEstimate (Map.cs:119-140): floor(expected size) picks of the heaviest component, each pick
    re-entering the list with its weight reduced by one.  Returns the picked means (k x 3)."""
    import heapq
    w, m = np.asarray(w, float), np.asarray(m, float).reshape(-1, 3)
    size = int(w.sum()) if len(w) else 0
    heap = [(-wi, i) for i, wi in enumerate(w)]
    heapq.heapify(heap)
    picks = []
    for _ in range(size):
        nw, i = heapq.heappop(heap)
        picks.append(i)
        heapq.heappush(heap, (nw + 1.0, i))
    return m[picks] if picks else np.zeros((0, 3))


def visited_map(rec):
    """Plot.VisitedMap (postanalysis/Plot.cs:230-248): every landmark that was visible AND detected in some frame."""
    seen = []
    for _, (wts, means, _) in rec.vismaps:
        for wt, mean in zip(wts, means):
            if wt > 0 and not any(np.linalg.norm(mean - s) <= 1e-5 for s in seen):
                seen.append(np.asarray(mean, float))
    return np.array(seen).reshape(-1, 3)


def map_error(rec, handle, c=1.0, p=2.0):
    """Plot.MapError / SpatialMapError (postanalysis/Plot.cs:477-527) with RefTime = 0 (estimate and groundtruth share
    the first pose, so the alignment transform is the identity): per recorded frame the OSPA distance between the
    visited groundtruth landmarks and the best map estimate, evaluated by librbphd.so (rbphd_ospa).
    Returns [(time, ospa, spatial)]."""
    visited = visited_map(rec)
    out = []
    for t, (w, m, _) in rec.maps:
        o, card = handle.ospa(visited, best_map_estimate(w, m), c, p)
        out.append((t, o, max(o ** p - card ** p, 0.0) ** (1.0 / p)))
    return out
