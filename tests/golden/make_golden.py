#!/usr/bin/env python
"""Generates tests/golden/slam_small.npz: a seeded 5-frame RB-PHD SLAM run of the ORACLE (the C# reference
cannot run in this image; the oracle is pinned on the reference's NUnit tests, see oracle/README.md).
The fixture freezes the oracle's outputs so that drift in either the oracle or the CUDA path is caught.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from monorfs_b200 import synth  # noqa: E402
from oracle import orc  # noqa: E402

P, N, M, FRAMES, SEED = 6, 30, 12, 5, 41


def run():
    sc = synth.make_scene(P, N, M, seed=SEED, min_effective_particle=0.5)
    nav = orc.Navigator(orc.make_config(sc.params), P, sc.poses[0])
    for i in range(P):
        nav.set_pose(i, sc.poses[i])
        nav.set_map(i, sc.map_w, sc.map_m, sc.map_P)
    out = dict(poses0=sc.poses, map_w=sc.map_w, map_m=sc.map_m, map_P=sc.map_P)
    for f in range(FRAMES):
        fr = sc.next_frame()
        nav.update(fr.reading, synth.DT, fr.gauss)
        best, res, anc = nav.slam_update(fr.z, fr.u)
        out["z%d" % f], out["gauss%d" % f], out["u%d" % f] = fr.z, fr.gauss, fr.u
        out["best%d" % f], out["res%d" % f], out["anc%d" % f] = best, res, anc
        out["w%d" % f], out["alpha%d" % f] = nav.get_weights(), nav.get_alphas()
        out["counts%d" % f] = np.array([len(nav.get_map(i)[0]) for i in range(P)])
    for i in range(P):
        w, m, Pm = nav.get_map(i)
        out["final_w%d" % i], out["final_m%d" % i], out["final_P%d" % i] = w, m, Pm
    out["final_poses"] = nav.get_poses()
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "slam_small.npz"), **run())
    print("written")
