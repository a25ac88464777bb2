#!/usr/bin/env python
"""Writes tests/golden/csharp/inputs_slam_small.txt: the seeded scene of tests/golden/make_golden.py as the plain
text DumpGolden.cs reads (name, count, values; repr() round-trips doubles).

    python tests/golden/csharp/export_inputs.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, ROOT)
from monorfs_b200 import synth  # noqa: E402

P, N, M, FRAMES, SEED = 6, 30, 12, 5, 41


def put(fh, name, values):
    v = np.asarray(values, dtype=np.float64).reshape(-1)
    fh.write("%s %d %s\n" % (name, len(v), " ".join(repr(float(x)) for x in v)))


def scene():
    return synth.make_scene(P, N, M, seed=SEED, min_effective_particle=0.5)


def main():
    sc = scene()
    path = os.path.join(ROOT, "tests", "golden", "csharp", "inputs_slam_small.txt")
    with open(path, "w") as fh:
        fh.write("# seeded scene of tests/golden/make_golden.py (P=%d N=%d M=%d frames=%d seed=%d)\n" % (P, N, M, FRAMES, SEED))
        for k, v in (("P", P), ("N", N), ("M", M), ("frames", FRAMES), ("dt", synth.DT)):
            put(fh, k, [v])
        p = sc.params
        for k in ("R", "Q", "birth_cov", "visibility_ramp", "measurer"):
            put(fh, k, p[k])
        for k in ("pd", "clutter", "birth_weight", "min_weight", "max_quantity", "merge_threshold",
                  "exploration_threshold", "density_distance_threshold", "min_effective_particle"):
            put(fh, k, [p[k]])
        put(fh, "poses0", sc.poses)
        put(fh, "map_w", sc.map_w)
        put(fh, "map_m", sc.map_m)
        put(fh, "map_P", sc.map_P)
        for f in range(FRAMES):
            fr = sc.next_frame()
            put(fh, "reading%d" % f, fr.reading)
            put(fh, "gauss%d" % f, fr.gauss)
            put(fh, "z%d" % f, fr.z)
    print("written", path)


if __name__ == "__main__":
    main()
