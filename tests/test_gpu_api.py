"""Behaviour of the C ABI beyond the arithmetic: state read/write, lifecycle, errors, re-entrancy."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from monorfs_b200 import capi, synth
    capi.load()
    return capi, synth


def test_map_and_pose_round_trip(env):
    capi, synth = env
    sc = synth.make_scene(5, 20, 8, seed=2)
    h = capi.Handle(sc.params, max_particles=8, max_components=64, max_measurements=16)
    h.reset(5, sc.poses[0], sc.map_w, sc.map_m, sc.map_P)
    assert h.particles == 5
    assert np.array_equal(h.get_weights(), np.full(5, 0.2))            # PHD:261
    assert h.get_best() == 0                                            # PHD:265
    assert h.get_map_counts().tolist() == [20] * 5
    for i in range(5):                                                  # deep copies of the same map (PHD:260)
        w, m, P = h.get_map(i)
        assert np.array_equal(w, sc.map_w) and np.array_equal(m, sc.map_m) and np.array_equal(P, sc.map_P)
    h.set_map(3, sc.map_w[:7], sc.map_m[:7], sc.map_P[:7])
    assert h.get_map_counts().tolist() == [20, 20, 20, 7, 20]
    assert np.array_equal(h.get_map(3)[1], sc.map_m[:7])
    h.set_poses(sc.poses)
    h.set_pose(2, [1, 2, 3, 1, 0, 0, 0])
    poses = h.get_poses()
    assert np.array_equal(poses[2], [1, 2, 3, 1, 0, 0, 0]) and np.array_equal(poses[4], sc.poses[4])
    h.clear_maps()                                                      # ResetMapModel (PHD:271-276)
    assert h.get_map_counts().tolist() == [0] * 5
    h.reset(3, sc.poses[1], sc.map_w[:4], sc.map_m[:4], sc.map_P[:4])   # CollapseParticles(3)
    assert h.particles == 3 and h.get_map_counts().tolist() == [4, 4, 4]
    assert np.array_equal(h.get_poses(), np.tile(sc.poses[1], (3, 1)))
    h.close()


def test_argument_and_capacity_errors(env):
    capi, synth = env
    sc = synth.make_scene(2, 20, 8, seed=2)
    h = capi.Handle(sc.params, max_particles=2, max_components=64, max_measurements=8)
    with pytest.raises(capi.RbphdError) as e:
        h.slam_update(np.zeros((2, 3)), 0.5)                            # not reset yet
    assert e.value.code == capi.ERR_ARGUMENT
    with pytest.raises(capi.RbphdError) as e:
        h.reset(3, sc.poses[0], sc.map_w, sc.map_m, sc.map_P)           # more particles than max_particles
    assert e.value.code == capi.ERR_ARGUMENT
    with pytest.raises(capi.RbphdError) as e:
        h.reset(2, sc.poses[0], np.ones(100), np.zeros((100, 3)), np.tile(np.eye(3), (100, 1, 1)))
    assert e.value.code == capi.ERR_CAPACITY
    h.reset(2, sc.poses[0], sc.map_w, sc.map_m, sc.map_P)
    with pytest.raises(capi.RbphdError) as e:
        h.slam_update(np.zeros((9, 3)), 0.5)                            # m > max_measurements
    assert e.value.code == capi.ERR_ARGUMENT
    with pytest.raises(capi.RbphdError):
        h.get_map(2)
    h.close()


def test_particle_depleted_matches_formula(env):
    capi, synth = env
    sc = synth.make_scene(4, 10, 4, seed=2, min_effective_particle=0.5)
    h = capi.Handle(sc.params, max_particles=4, max_components=32, max_measurements=8)
    h.reset(4, sc.poses[0], sc.map_w, sc.map_m, sc.map_P)
    assert not h.particle_depleted()                                    # ESS = 4 >= 0.5 * 4
    h.set_weights([0.97, 0.01, 0.01, 0.01])
    assert h.particle_depleted()                                        # ESS ~ 1.06 < 2   (PHD:768-777)
    h.close()


def test_handles_are_independent_and_reentrant(env):
    """LoopyPHDNavigator creates one navigator per Parallel.For task (LoopyPHDNavigator.cs:525-551)."""
    capi, synth = env
    frames = 4

    def run(seed, out, idx):
        sc = synth.make_scene(6, 40, 16, seed=seed, min_effective_particle=0.5)
        h = capi.Handle(sc.params, max_particles=6, max_components=128, max_measurements=16)
        h.reset(6, sc.poses[0], sc.map_w, sc.map_m, sc.map_P)
        h.set_poses(sc.poses)
        res = []
        for _ in range(frames):
            fr = sc.next_frame()
            h.update(fr.reading, synth.DT, fr.gauss)
            res.append(h.slam_update(fr.z, fr.u))
        out[idx] = (res, h.get_map_counts().tolist(), h.get_ancestors().tolist(), h.get_map(0))
        h.close()

    serial = [None] * 4
    for i in range(4):
        run(50 + i, serial, i)
    threaded = [None] * 4
    ts = [threading.Thread(target=run, args=(50 + i, threaded, i)) for i in range(4)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for a, b in zip(serial, threaded):
        assert a[0] == b[0] and a[1] == b[1] and a[2] == b[2]
        for x, y in zip(a[3], b[3]):
            assert np.array_equal(x, y)


def test_async_frames_match_synchronous_calls(env):
    """rbphd_frame_async (device-resident loop used by bench.py) == rbphd_update + rbphd_slam_update."""
    capi, synth = env
    sc = synth.make_scene(8, 40, 16, seed=9, min_effective_particle=0.5)
    frames = [sc.next_frame() for _ in range(5)]
    a = capi.Handle(sc.params, max_particles=8, max_components=128, max_measurements=16, resident_frames=5)
    b = capi.Handle(sc.params, max_particles=8, max_components=128, max_measurements=16)
    for h in (a, b):
        h.reset(8, sc.poses[0], sc.map_w, sc.map_m, sc.map_P)
        h.set_poses(sc.poses)
    for f, fr in enumerate(frames):
        a.upload_frame_inputs(fr.gauss, fr.z, slot=f)
    for f, fr in enumerate(frames):
        a.frame_async(fr.reading, synth.DT, 16, fr.u, slot=f)
        b.update(fr.reading, synth.DT, fr.gauss)
        b.slam_update(fr.z, fr.u)
    a.synchronize()
    assert a.get_best() == b.get_best()
    assert a.get_map_counts().tolist() == b.get_map_counts().tolist()
    assert np.allclose(a.get_weights(), b.get_weights(), rtol=1e-12, atol=0)
    assert np.array_equal(a.get_poses(), b.get_poses())
    for i in range(8):
        for x, y in zip(a.get_map(i), b.get_map(i)):
            assert np.array_equal(x, y)
    ctr = a.counters()
    assert ctr["particle_frames"] == 8 * 5 and ctr["comps_in"] > 0 and ctr["pairs"] > 0
    assert a.kernel_launches >= 5 * 5
    a.close()
    b.close()


def test_run_to_run_determinism(env):
    """Two runs on the same inputs: maps, counts, poses, ancestors and decisions are bit-identical (lists built with
    atomics are sorted before use); particle weights agree to 1e-12 (the Map.Evaluate sums use FP64 atomics)."""
    capi, synth = env

    def run():
        sc = synth.make_scene(48, 150, 48, seed=41, min_effective_particle=0.5)
        h = capi.Handle(sc.params, max_particles=48, max_components=300, max_measurements=48, max_pairs=16 * 48)
        h.reset(48, sc.poses[0], sc.map_w, sc.map_m, sc.map_P)
        h.set_poses(sc.poses)
        out = []
        for _ in range(6):
            fr = sc.next_frame()
            h.update(fr.reading, synth.DT, fr.gauss)
            out.append(h.slam_update(fr.z, fr.u))
        res = (out, h.get_map_counts().copy(), h.get_ancestors().copy(), h.get_poses().copy(), h.get_weights().copy(),
               [h.get_map(i) for i in (0, 17, 47)])
        h.close()
        return res

    a, b = run(), run()
    assert a[0] == b[0]
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])
    assert np.allclose(a[4], b[4], rtol=1e-12, atol=0)
    for ma, mb in zip(a[5], b[5]):
        for x, y in zip(ma, mb):
            assert np.array_equal(x, y)
