"""Host-side pieces of the headless run / replay (monorfs_b200/simulation.py): pure numpy, no GPU."""
import numpy as np

from monorfs_b200 import recordio, simulation, synth


def test_diff_odometry_inverts_add_odometry():
    """Pose3D.DiffOdometry (POSE:336-350) is the inverse of AddOdometry (POSE:314-333)."""
    rng = np.random.default_rng(4)
    pose = np.array([0.3, -0.2, 1.0, 1, 0, 0, 0.0])
    for _ in range(20):
        pose = synth.add_odometry(pose, rng.normal(size=6) * [0.1, 0.1, 0.1, 0.05, 0.05, 0.05])
        d = rng.normal(size=6) * [0.02, 0.02, 0.02, 0.004, 0.004, 0.004]
        q = synth.add_odometry(pose, d)
        assert np.allclose(simulation.diff_odometry(q, pose), d, rtol=0, atol=1e-12)
    assert np.allclose(simulation.diff_odometry(pose, pose), 0, atol=1e-15)


def test_true_vehicle_measures_like_the_reference():
    """SimulatedVehicle.Measure (SimulatedVehicle.cs:243-295): detections only for landmarks with PD > 0, clutter count
    capped at 10 lambda, everything inside the film / range clip; ReadOdometry resets the odometry pose."""
    pose0, measurer, landmarks = simulation.synthetic_scene(300, seed=9)
    prm = synth.params(300)
    prm["measurer"] = [float(v) for v in measurer]
    v = simulation.TrueVehicle(pose0, landmarks, prm, np.random.default_rng(5))
    assert abs(v.clutter_count - prm["clutter"] * 640 * 480 * (10 - 0.1)) < 1e-3
    counts = []
    for _ in range(30):
        v.update(np.array(synth.ODOMETRY), synth.DT)
        reading = v.read_odometry()
        assert np.allclose(v.odometry_pose, v.pose) and np.allclose(v.ref_odometry, v.pose)
        assert np.allclose(reading, synth.ODOMETRY, atol=0.02)     # the command plus N(0, Q dt^2) corruption
        z, visible, detected = v.measure()
        assert z.shape[1] == 3 and len(visible) == len(detected)
        assert np.all((z[:, 0] >= -320) & (z[:, 0] <= 320) | (np.abs(z[:, 0]) < 330))   # noise may cross the border
        assert np.all(z[:, 2] > 0)
        assert len(z) <= int(detected.sum()) + int(v.clutter_count * 10)
        counts.append(len(z))
    assert np.mean(counts) > 5


def test_synthetic_run_files_round_trip():
    pose0, measurer, landmarks = simulation.synthetic_scene(25, seed=2)
    cmds = simulation.synthetic_commands(12)
    p2, m2, l2 = recordio.parse_scene(recordio.scene_to_text(pose0, measurer, landmarks, lossless=True))
    assert np.array_equal(p2, pose0) and np.array_equal(m2, measurer) and np.array_equal(l2, landmarks)
    c2 = recordio.parse_commands(recordio.commands_to_text(cmds))
    assert len(c2) == 12 and c2[0][6] == 1 and all(c[6] == 0 for c in c2[1:])
    assert np.allclose(c2[3][:6], synth.ODOMETRY)


def test_best_map_estimate_matches_the_oracle():
    """Map.BestMapEstimate (Map.cs:119-140): the host-side post-analysis helper picks the oracle's components."""
    from oracle import orc
    rng = np.random.default_rng(3)
    for n in (0, 1, 7, 60):
        w = rng.choice([0.05, 0.4, 0.97, 1.6, 2.3], size=n) + rng.random(n) * 1e-3
        m = rng.normal(size=(n, 3))
        picks = orc.best_map_estimate(w)
        got = simulation.best_map_estimate(w, m)
        assert len(got) == len(picks) == int(w.sum())
        assert np.array_equal(got, m[picks].reshape(-1, 3))


def test_visited_map_keeps_detected_landmarks_once():
    rec = recordio.Recording(np.zeros(7), np.zeros(7), np.zeros((0, 3)))
    a, b, c = np.array([1.0, 0, 0]), np.array([0, 2.0, 0]), np.array([0, 0, 3.0])
    cov = np.tile(np.eye(3) * 1e-3, (2, 1, 1))
    rec.vismaps.append((0.1, (np.array([1.0, 0.0]), np.array([a, b]), cov)))      # b visible, not detected
    rec.vismaps.append((0.2, (np.array([1.0, 1.0]), np.array([a + 1e-7, c]), cov)))
    v = simulation.visited_map(rec)
    assert len(v) == 2 and np.array_equal(v[0], a) and np.array_equal(v[1], c)
