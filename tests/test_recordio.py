"""Record / replay wire format (monorfs_b200.recordio) against the reference's formats as its FileParser reads them
(Util/FileParser.cs:56-340, RecordVehicle.cs:244-347, Vehicle.cs:503-524, Simulation.cs:144-231)."""
import os

import numpy as np
import pytest

from monorfs_b200 import recordio as rio


def test_g6_matches_dotnet():
    assert rio.g6(0.0333333333) == "0.0333333"
    assert rio.g6(1234567.0) == "1.23457e+06"
    assert rio.g6(1e-5) == "1e-05"
    assert rio.g6(575.8156) == "575.816"
    assert rio.g6(-0.0) == "0"
    assert rio.g6(2) == "2"


def test_scene_descriptor_layout_and_round_trip():
    pose = [0, 0, 0, 1, 0, 0, 0]
    measurer = [575.8156, 0.1, 10.0, -320, -240, 640, 480]
    lm = np.array([[1.0, 2.0, 3.0], [-0.5, 0.25, 7.125]])
    text = rio.scene_to_text(pose, measurer, lm)
    assert text == ("pose\n\t0 0 0 1 0 0 0\nparams\n\t575.816 0.1 10 -320 -240 640 480\n"
                    "landmarks\n\t1 2 3\n\t-0.5 0.25 7.125\n")
    p, m, l = rio.parse_scene(text)
    assert np.array_equal(p, pose) and np.array_equal(l, lm) and m[0] == 575.816
    # the deprecated "focal" key and Windows line ends (Util.ParseDictionary)
    p, m, l = rio.parse_scene("pose\r\n\t1 2 3 1 0 0 0\r\nfocal\r\n\t500 0.1 2 -320 -240 640 480\r\nlandmarks\r\n\t0 0 1\r\n")
    assert m[0] == 500 and len(l) == 1
    with pytest.raises(ValueError):
        rio.parse_scene("pose\n\t0 0 0 1 0 0 0\nlandmarks\n\t1 2\n")


def test_reference_style_descriptors_parse():
    g = rio.parse_gaussian("0.5;1 2 3;1 0 0 0 1 0 0 0 1")
    assert g[0] == 0.5 and g[1].tolist() == [1, 2, 3] and np.array_equal(g[2], np.eye(3))
    with pytest.raises(ValueError):
        rio.parse_gaussian("0.5;1 2 3;1 0 0 0 1")
    maps = rio.parse_map_history("0.0333333\n0.5;1 2 3;1 0 0 0 1 0 0 0 1\n|\n0.0666667\n0.25;0 0 1;2 0 0 0 2 0 0 0 2\n0.75;0 1 0;1 0 0 0 1 0 0 0 1")
    assert [len(m[1][0]) for m in maps] == [1, 2] and maps[1][0] == 0.0666667
    meas = rio.parse_measurements("0.0333333:1 2 3;4.5 -6 0.75\n0.0666667:")
    assert meas[0][1].shape == (2, 3) and meas[1][1].shape == (0, 3)
    with pytest.raises(ValueError):
        rio.parse_measurements("0.1 1 2 3")
    traj = rio.parse_trajectory_history("0.1\n0.1 0 0 0 1 0 0 0\n|\n0.2\n0.1 0 0 0 1 0 0 0\n0.2 0 0 0.01 1 0 0 0", 7)
    assert [len(t[1]) for t in traj] == [1, 2]
    with pytest.raises(ValueError):
        rio.parse_timed_array(["0.1 1 2 3"], 7)
    cmds = rio.parse_commands("0 0 0.01 0 0.002 0 0\n0 0 0.01 0 0 0 1\n")
    assert len(cmds) == 2 and cmds[1][6] == 1


@pytest.mark.parametrize("lossless", [False, True])
def test_archive_round_trip(tmp_path, lossless):
    rng = np.random.default_rng(3)
    rec = rio.Recording(np.array([0, 0, 0, 1.0, 0, 0, 0]), np.array([575.8156, 0.1, 10, -320, -240, 640, 480]),
                        rng.normal(size=(5, 3)))
    for f in range(4):
        t = (f + 1) / 30.0
        rec.trajectory.append((t, rng.normal(size=7)))
        rec.odometry.append((t, rng.normal(size=6) * 1e-2))
        rec.measurements.append((t, rng.normal(size=(f, 3))))
        rec.estimate.append((t, [(tt, st) for tt, st in rec.trajectory]))
        n = 2 + f
        rec.maps.append((t, (rng.random(n), rng.normal(size=(n, 3)), np.tile(np.eye(3) * 1e-3, (n, 1, 1)))))
        rec.vismaps.append((t, (np.ones(1), rng.normal(size=(1, 3)), np.tile(np.eye(3) * 1e-3, (1, 1, 1)))))
    rec.tags.append((0.1, "screenshot one"))
    rec.config = "MaxQuantity = 600\n"
    path = os.path.join(tmp_path, "data.zip")
    rio.save(rec, path, lossless=lossless)
    back = rio.load(path)
    tol = 0 if lossless else 1e-5
    assert len(back.trajectory) == 4 and len(back.maps) == 4 and back.tags == [(0.1, "screenshot one")]
    for (ta, a), (tb, b) in zip(rec.trajectory, back.trajectory):
        assert np.allclose(a, b, rtol=tol, atol=tol * 1e-3)
    for (ta, a), (tb, b) in zip(rec.measurements, back.measurements):
        assert np.array_equal(a, b)          # measurements are always written in full precision
    for (ta, a), (tb, b) in zip(rec.maps, back.maps):
        for x, y in zip(a, b):
            assert np.allclose(x, y, rtol=tol, atol=tol * 1e-3)
    assert [len(e[1]) for e in back.estimate] == [1, 2, 3, 4]
    assert back.config == rec.config
    # the mandatory members (RecordVehicle.cs:261-279)
    import zipfile
    bad = os.path.join(tmp_path, "bad.zip")
    with zipfile.ZipFile(bad, "w") as zf:
        zf.writestr("scene.world", rio.scene_to_text(rec.pose0, rec.measurer, rec.landmarks))
    with pytest.raises(ValueError):
        rio.load(bad)
