"""tests/golden/slam_small.npz (made by tests/golden/make_golden.py from the oracle): the oracle must keep
reproducing it bit for bit (CPU), and the CUDA path must match it through the C ABI (GPU)."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIX = os.path.join(ROOT, "tests", "golden", "slam_small.npz")
P, N, M, FRAMES = 6, 30, 12, 5


def test_oracle_reproduces_fixture():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(ROOT, "tests", "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    now = mg.run()
    g = np.load(FIX)
    nres = 0
    for k in g.files:
        assert np.array_equal(np.asarray(now[k]), g[k]), k
        nres += int(k.startswith("res") and bool(g[k]))
    assert nres >= 1, "fixture must contain a resampling frame"


@pytest.mark.gpu
def test_gpu_matches_fixture():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from monorfs_b200 import capi, synth
    g = np.load(FIX)
    sc = synth.make_scene(P, N, M, seed=41, min_effective_particle=0.5)
    h = capi.Handle(sc.params, max_particles=P, max_components=128, max_measurements=M)
    h.reset(P, g["poses0"][0], g["map_w"], g["map_m"], g["map_P"])
    h.set_poses(g["poses0"])
    for f in range(FRAMES):
        h.update(synth.ODOMETRY, synth.DT, g["gauss%d" % f])
        best, res = h.slam_update(g["z%d" % f], float(g["u%d" % f]))
        assert best == int(g["best%d" % f]) and res == bool(g["res%d" % f])
        assert h.get_ancestors().tolist() == g["anc%d" % f].tolist()
        assert h.get_map_counts().tolist() == g["counts%d" % f].tolist()
        assert np.allclose(h.get_weights(), g["w%d" % f], rtol=1e-9, atol=0)
        with np.errstate(divide="ignore"):
            assert np.allclose(np.log(h.get_alphas()), np.log(g["alpha%d" % f]), rtol=1e-9, atol=0)
    assert np.allclose(h.get_poses(), g["final_poses"], rtol=0, atol=1e-12)
    for i in range(P):
        w, m, Pm = h.get_map(i)
        assert np.allclose(w, g["final_w%d" % i], rtol=1e-9, atol=0)
        assert np.allclose(m, g["final_m%d" % i], rtol=1e-9, atol=1e-12)
        assert np.allclose(Pm, g["final_P%d" % i], rtol=1e-9, atol=1e-15)
    h.close()
