"""Host-side logic that needs no GPU: the synthetic scene generator and the sharding plan."""
import numpy as np

from monorfs_b200 import sharded, synth


def test_scene_is_seeded_and_shaped():
    a = synth.make_scene(4, 30, 12, seed=5)
    b = synth.make_scene(4, 30, 12, seed=5)
    assert np.array_equal(a.map_m, b.map_m) and np.array_equal(a.poses, b.poses)
    fa, fb = a.next_frame(), b.next_frame()
    assert fa.z.shape == (12, 3) and fa.gauss.shape == (4, 6)
    assert np.array_equal(fa.z, fb.z) and fa.u == fb.u
    assert np.allclose(np.linalg.norm(a.poses[:, 3:], axis=1), 1.0)
    assert np.all(np.linalg.eigvalsh(a.map_P) > 0)
    assert float(np.float32(fa.u)) == fa.u            # AForge's generators return float (PHD:727)
    assert a.params["max_quantity"] == 60


def test_scene_measurement_model_matches_oracle(orc):
    sc = synth.make_scene(2, 20, 8, seed=9)
    cfg = orc.make_config(sc.params)
    pose = sc.poses[1]
    z = synth.measure_perfect(pose, sc.landmarks, sc.params["measurer"][0])
    for i in range(5):
        assert np.allclose(z[i], orc.measure_perfect(cfg, pose, sc.landmarks[i]), rtol=1e-12, atol=1e-12)
    nxt = synth.add_odometry(pose, np.array(synth.ODOMETRY))
    assert np.allclose(nxt, orc.pose_add_odometry(pose, synth.ODOMETRY), rtol=1e-13, atol=1e-15)


def test_block_partition_covers_everything():
    for total, world in [(20000, 8), (2000, 3), (7, 4), (5, 8)]:
        seen = []
        for r in range(world):
            lo, hi = sharded.block_range(r, world, total)
            seen += list(range(lo, hi))
            for i in range(lo, hi):
                assert sharded.owner_of(i, world, total) == r
        assert seen == list(range(total))


def test_migration_plan_is_consistent():
    rng = np.random.default_rng(3)
    total, world = 40, 4
    anc = np.sort(rng.integers(0, total, total))
    plans = [sharded.migration_plan(anc, r, world) for r in range(world)]
    for r, p in enumerate(plans):
        lo, hi = sharded.block_range(r, world, total)
        # every local slot is covered exactly once (local copy or remote record)
        covered = {int(s) for s in np.nonzero(p["local_sources"] >= 0)[0]}
        for src, items in p["recv"].items():
            assert src != r
            for a, slots in items:
                assert sharded.owner_of(a, world, total) == src
                for s in slots:
                    assert s not in covered and anc[lo + s] == a
                    covered.add(s)
        assert covered == set(range(hi - lo))
        # what I expect from `src` is exactly what `src` plans to send me, in the same order
        for src, items in p["recv"].items():
            slo, _ = sharded.block_range(src, world, total)
            assert plans[src]["send"][r] == [a - slo for a, _ in items]
    # nothing is sent that is not awaited
    for r, p in enumerate(plans):
        for dest, idx in p["send"].items():
            assert len(plans[dest]["recv"][r]) == len(idx)


def test_kernel_counter_hash_ignores_comments_only():
    """profiles/kernel_counters.json is tied to the device CODE: comments and white space do not change the hash,
    a changed token does."""
    from monorfs_b200 import build
    a = "int x = 1; // note\n/* block\n comment */ int y = x  +  2;\n"
    b = "int x = 1;\nint y = x + 2; // other words\n"
    c = "int x = 1;\nint y = x + 3;\n"
    assert build._code_only(a) == build._code_only(b)
    assert build._code_only(a) != build._code_only(c)
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    rec = json.load(open(os.path.join(root, "profiles", "kernel_counters.json")))
    assert rec["source_hash"] == build.source_hash(), "profiles/kernel_counters.json was captured on other device code"
