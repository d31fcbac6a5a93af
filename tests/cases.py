"""Shared test cases: handcrafted topologies named after the reference's regression inputs
(test/CMakeLists.txt:91-107) plus scaled-down synthetic configs of BASELINE.json."""
import os

import numpy as np

from itkcontourinterpolation_b200.synthetic import ellipsoid_volume, make_config, sparsify

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def disk(shape, cy, cx, r):
    yy, xx = np.mgrid[:shape[0], :shape[1]]
    return (yy - cy) ** 2 + (xx - cx) ** 2 <= r * r


def handcrafted():
    """name -> volume [z, y, x] (uint16)."""
    cases = {}
    v = np.zeros((9, 40, 48), np.uint16)
    cases["Empty"] = v.copy()
    w = v.copy()
    w[3, 5:20, 5:20] = 2
    w[4, 5:20, 5:20] = 2  # adjacent slices only: nothing to interpolate
    cases["NoSlices"] = w
    w = v.copy()
    w[1][disk(w[1].shape, 20, 20, 9)] = 1
    w[7][disk(w[7].shape, 22, 27, 6)] = 1
    cases["SimplestOneToOne"] = w
    w = v.copy()
    w[0][disk(w[0].shape, 20, 24, 15)] = 3
    w[8][disk(w[8].shape, 10, 12, 5)] = 3
    w[8][disk(w[8].shape, 30, 14, 4)] = 3
    w[8][disk(w[8].shape, 20, 36, 6)] = 3
    cases["OneToThree"] = w
    w = v.copy()
    w[2][disk(w[2].shape, 20, 20, 8)] = 1
    w[6][disk(w[6].shape, 20, 20, 8)] = 1
    w[6][disk(w[6].shape, 8, 40, 4)] = 1  # appears without counterpart: extrapolation
    cases["ExtrapolationAppearing"] = w
    w = v.copy()
    w[0][disk(w[0].shape, 14, 14, 7)] = 1
    w[0][disk(w[0].shape, 14, 34, 7)] = 1
    w[5][disk(w[5].shape, 14, 24, 13)] = 1
    w[0][disk(w[0].shape, 32, 24, 5)] = 2
    w[5][disk(w[5].shape, 33, 20, 3)] = 2
    w[5][disk(w[5].shape, 33, 30, 3)] = 2
    cases["DoubleTwoLabelBranching"] = w
    # ring: centroid outside the shape, extrapolation must walk the BFS to land on it
    w = v.copy()
    ring = disk(w[1].shape, 20, 24, 14) & ~disk(w[1].shape, 20, 24, 9)
    w[1][ring] = 4
    w[6][disk(w[6].shape, 20, 24, 14)] = 4
    w[6][disk(w[6].shape, 20, 24, 9)] = 0
    w[1][disk(w[1].shape, 4, 4, 2)] = 4  # small island with no counterpart
    cases["RingAndIsland"] = w
    # a ring with no counterpart: Extrapolate's Align starts at the centroid (in the hole) with score 0
    w = v.copy()
    w[1][disk(w[1].shape, 20, 30, 12) & ~disk(w[1].shape, 20, 30, 8)] = 6
    w[1][disk(w[1].shape, 8, 6, 4)] = 6
    w[5][disk(w[5].shape, 8, 6, 5)] = 6
    cases["RingExtrapolate"] = w
    # two axes annotated, overlapping labels: cross-axis max rule
    w = np.zeros((24, 24, 24), np.uint16)
    dense = np.zeros_like(w)
    zz, yy, xx = np.mgrid[:24, :24, :24]
    dense[(zz - 12) ** 2 + (yy - 12) ** 2 + (xx - 12) ** 2 <= 81] = 5
    dense[(zz - 12) ** 2 + (yy - 8) ** 2 + (xx - 15) ** 2 <= 16] = 9
    cases["TwoAxisTwoLabel"] = sparsify(dense, (6, 0, 5))
    cases["ThreeAxis"] = sparsify(dense, (7, 6, 5))
    # odd gaps and translation: shapes drift between slices (exercises the carry rule)
    w = np.zeros((14, 60, 64), np.uint16)
    w[0][disk(w[0].shape, 20, 20, 10)] = 7
    w[5][disk(w[5].shape, 27, 31, 12)] = 7
    w[12][disk(w[12].shape, 39, 44, 7)] = 7
    cases["DriftOddGaps"] = w
    # shape pushed against the border so that translated shapes get clipped
    w = np.zeros((8, 30, 30), np.uint16)
    w[0][disk(w[0].shape, 3, 3, 6)] = 1
    w[7][disk(w[7].shape, 10, 12, 9)] = 1
    cases["BorderClip"] = w
    return cases


def synthetic_small():
    out = {}
    for name, scale in (("C3", 0.25), ("C4", 0.125), ("C5", 0.0625)):
        vol, _, cfg = make_config(name, scale)
        out[name] = (vol, cfg)
    # branching variant: two lobes merging along z
    dense = ellipsoid_volume((96, 96, 64), 6, 77, (0.12, 0.25), branching=6)
    out["Branching"] = (sparsify(dense, (0, 0, 6)), dict(axis=2, use_dt=True, ball=False))
    return out


ALGOS = {"T": dict(use_distance_transform=True, use_ball=True),   # test driver: T = DT with ball (8-connected CC)
         "B": dict(use_distance_transform=False, use_ball=True),
         "C": dict(use_distance_transform=False, use_ball=False),
         "D": dict(use_distance_transform=True, use_ball=False)}   # filter defaults

# argument sets of the reference's regression tests on 64816L_amygdala_int (test/CMakeLists.txt:143-160): (axis, label)
AMYGDALA_ARGS = [(-1, 0), (0, 0), (1, 0), (2, 0)] + [(a, l) for a in (0, 1, 2) for l in (1, 2, 3, 4)]


def amygdala():
    return np.load(os.path.join(GOLDEN, "64816L_amygdala_int.npz"))["vol"]
