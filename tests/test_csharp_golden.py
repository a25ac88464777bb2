"""Oracle (CPU) and CUDA path (GPU) against outputs of the REAL C# reference, when someone has produced them with
tests/golden/csharp/DumpGolden.cs (it cannot run in this image: no Mono / .NET).  Skipped while
tests/golden/csharp_slam_small.txt is absent.  This is the one-command check that turns the "parity unpinned"
items of oracle/README.md (KD-tree metric and order, sort ties, WeightAlpha) into pinned ones.

Maps are compared as multisets (sorted by weight, then mean): the reference enumerates a Map in KD-tree order."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "csharp_slam_small.txt")
INP = os.path.join(ROOT, "tests", "golden", "csharp", "inputs_slam_small.txt")

pytestmark = pytest.mark.skipif(not os.path.exists(OUT), reason="no C# golden output (run tests/golden/csharp/DumpGolden.cs)")


def read(path):
    d = {}
    with open(path) as fh:
        for line in fh:
            tok = line.split()
            if len(tok) < 2 or tok[0].startswith("#"):
                continue
            n = int(tok[1])
            d[tok[0]] = np.array([float(x) for x in tok[2:2 + n]])
    return d


def canon(w, m, P):
    w, m, P = np.asarray(w).reshape(-1), np.asarray(m).reshape(-1, 3), np.asarray(P).reshape(-1, 9)
    order = np.lexsort((m[:, 2], m[:, 1], m[:, 0], w))
    return w[order], m[order], P[order]


def assert_same_map(got, exp, what):
    gw, gm, gP = canon(*got)
    ew, em, eP = canon(*exp)
    assert len(gw) == len(ew), (what, len(gw), len(ew))
    assert np.allclose(gw, ew, rtol=1e-9, atol=1e-300), what
    assert np.allclose(gm, em, rtol=1e-9, atol=1e-12), what
    assert np.allclose(gP, eP, rtol=1e-9, atol=1e-15), what


def golden_map(g, prefix):
    return g[prefix + "_w"], g[prefix + "_m"], g[prefix + "_P"]


def scene_inputs():
    i = read(INP)
    P, M, frames = int(i["P"][0]), int(i["M"][0]), int(i["frames"][0])
    return i, P, M, frames


def params_from(i):
    from monorfs_b200 import synth
    p = synth.params(int(i["N"][0]), min_effective_particle=float(i["min_effective_particle"][0]))
    return p


def test_inputs_match_the_seeded_scene():
    """The committed input file is what export_inputs.py writes today (the C# outputs belong to these inputs)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("export_inputs", os.path.join(ROOT, "tests", "golden", "csharp", "export_inputs.py"))
    ex = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ex)
    i = read(INP)
    sc = ex.scene()
    assert np.array_equal(i["poses0"], sc.poses.reshape(-1))
    assert np.array_equal(i["map_m"], sc.map_m.reshape(-1))


def run(nav_kind):
    from monorfs_b200 import synth
    from oracle import orc
    g = read(OUT)
    i, P, M, frames = scene_inputs()
    p = params_from(i)
    poses0 = i["poses0"].reshape(P, 7)
    mw, mm, mP = i["map_w"], i["map_m"].reshape(-1, 3), i["map_P"].reshape(-1, 3, 3)
    ocfg = orc.make_config(p)
    if nav_kind == "oracle":
        nav = orc.Navigator(ocfg, P, poses0[0])
        for k in range(P):
            nav.set_pose(k, poses0[k])
            nav.set_map(k, mw, mm, mP)
        z0 = i["z0"].reshape(-1, 3)
        pred = orc.predict(ocfg, poses0[0], mw, mm, mP, z0)[:3]
        corr = orc.correct(ocfg, poses0[0], *pred, z0)
        prun = orc.prune(ocfg, *corr)
        assert_same_map(pred, golden_map(g, "stage_predicted"), "PredictConditional")
        assert_same_map(corr, golden_map(g, "stage_corrected"), "CorrectConditional")
        assert_same_map(prun, golden_map(g, "stage_pruned"), "PruneModel")
        wa = orc.weight_alpha(ocfg, poses0[0], z0, pred, prun)
        assert np.isclose(np.log(wa["alpha"]), np.log(g["stage_alpha"][0]), rtol=1e-9)
    else:
        from monorfs_b200 import capi
        nav = capi.Handle(p, max_particles=P, max_components=128, max_measurements=M)
        nav.reset(P, poses0[0], mw, mm, mP)
        nav.set_poses(poses0)
    for f in range(frames):
        u = float(g["u%d" % f][0])
        nav.update(i["reading%d" % f], synth.DT, i["gauss%d" % f].reshape(P, 6))
        out = nav.slam_update(i["z%d" % f].reshape(-1, 3), u)
        best, res = out[0], out[1]
        assert best == int(g["best%d" % f][0]) and bool(res) == bool(g["res%d" % f][0]), f
        assert np.allclose(nav.get_weights(), g["w%d" % f], rtol=1e-9, atol=0), f
        assert np.allclose(nav.get_poses().reshape(-1), g["poses%d" % f], rtol=0, atol=1e-12), f
        counts = [len(nav.get_map(k)[0]) for k in range(P)]
        assert counts == [int(c) for c in g["counts%d" % f]], f
    for k in range(P):
        assert_same_map(nav.get_map(k), golden_map(g, "final%d" % k), "final map %d" % k)
    nav.close()


def test_oracle_matches_csharp():
    run("oracle")


@pytest.mark.gpu
def test_gpu_matches_csharp():
    run("gpu")
