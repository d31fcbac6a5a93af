"""Oracle against the golden vectors and the invariants the reference's own tests state
(test/CMakeLists.txt:91-125 comments; X:463-466, X:1623-1636).  CPU only."""
import hashlib
import json
import os

import numpy as np
import pytest

import cases
from oracle import oracle_lib as O

GOLD = json.load(open(os.path.join(cases.GOLDEN, "golden.json")))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def all_inputs(seven_labels, many_to_many):
    d = {"SevenLabels": seven_labels, "ManyToMany": many_to_many, "SevenLabels_u8": seven_labels.astype(np.uint8)}
    d.update(cases.handcrafted())
    return d


def test_golden_vectors(seven_labels, many_to_many):
    for name, vol in all_inputs(seven_labels, many_to_many).items():
        for algo, kw in cases.ALGOS.items():
            g = GOLD["%s_%s" % (name, algo)]
            assert sha(vol) == g["input"], name
            assert sha(O.interpolate(vol, **kw)) == g["output"], (name, algo)


def test_golden_synthetic():
    for name, (vol, cfg) in cases.synthetic_small().items():
        g = GOLD["synthetic_%s" % name]
        assert sha(vol) == g["input"], name
        out = O.interpolate(vol, axis=cfg.get("axis", -1), use_distance_transform=cfg["use_dt"], use_ball=cfg["ball"])
        assert sha(out) == g["output"], name


def test_empty_and_no_slices_are_identity():
    h = cases.handcrafted()
    for name in ("Empty", "NoSlices"):
        assert np.array_equal(O.interpolate(h[name]), h[name])


def test_output_contains_input(seven_labels, many_to_many):
    for vol in all_inputs(seven_labels, many_to_many).values():
        out = O.interpolate(vol)
        assert np.array_equal(out[vol != 0], vol[vol != 0])


def test_identical_slices_give_identical_mids():
    v = np.zeros((9, 32, 32), np.uint16)
    m = cases.disk((32, 32), 15, 16, 9)
    v[0][m] = 2
    v[8][m] = 2
    for kw in cases.ALGOS.values():
        out = O.interpolate(v, **kw)
        for z in range(9):
            assert np.array_equal(out[z] == 2, m)


def test_axis_and_label_restrictions(seven_labels):
    # SevenLabels is annotated along y only: axis 0 and 2 leave it untouched, axis 1 equals all-axes
    full = O.interpolate(seven_labels)
    assert np.array_equal(O.interpolate(seven_labels, axis=1), full)
    for ax in (0, 2):
        assert np.array_equal(O.interpolate(seven_labels, axis=ax), seven_labels)
    only3 = O.interpolate(seven_labels, label=3)
    changed = only3 != seven_labels
    assert changed.any() and (only3[changed] == 3).all()
    assert np.array_equal(O.interpolate(seven_labels, label=200), seven_labels)


def test_thread_count_independence(many_to_many):
    vol, cfg = cases.synthetic_small()["C4"]
    a = O.interpolate(vol, use_distance_transform=False, use_ball=True, threads=1)
    b = O.interpolate(vol, use_distance_transform=False, use_ball=True, threads=7)
    assert np.array_equal(a, b)


def test_custom_slice_positions(many_to_many):
    # positions given explicitly must reproduce auto-detection when they are the same positions
    auto = O.interpolate(many_to_many)
    cust = O.interpolate(many_to_many, custom_slices=[(2, 1, 0), (2, 1, 7)])
    assert np.array_equal(auto, cust)


def test_no_extrapolation_writes_less():
    v = cases.handcrafted()["ExtrapolationAppearing"]
    with_x = O.interpolate(v)
    without = O.interpolate(v, use_extrapolation=False)
    assert (with_x != 0).sum() > (without != 0).sum()
    assert ((without != 0) <= (with_x != 0)).all()


def test_detection_matches_brute_force_rule():
    vol, _ = cases.synthetic_small()["C4"]
    trip, boxes = O.detect(vol)
    got = set(map(tuple, trip.tolist()))
    want = set()
    Z, Y, X = vol.shape
    pad = np.pad(vol, 1)
    c = pad[1:-1, 1:-1, 1:-1]
    nb = {0: (pad[1:-1, 1:-1, :-2], pad[1:-1, 1:-1, 2:]), 1: (pad[1:-1, :-2, 1:-1], pad[1:-1, 2:, 1:-1]),
          2: (pad[:-2, 1:-1, 1:-1], pad[2:, 1:-1, 1:-1])}
    empty = {a: (nb[a][0] == 0) & (nb[a][1] == 0) for a in nb}
    same = {a: (nb[a][0] == c) & (nb[a][1] == c) for a in nb}
    for a in range(3):
        o1, o2 = [b for b in range(3) if b != a]
        ok = (c != 0) & empty[a] & same[o1] & same[o2]
        zz, yy, xx = np.nonzero(ok)
        pos = (xx, yy, zz)[a]
        for l, s in zip(c[ok].tolist(), pos.tolist()):
            want.add((a, l, s))
    assert got == want
    for row in boxes:
        l = row[0]
        zz, yy, xx = np.nonzero(vol == l)
        assert list(row[1:4]) == [xx.min(), yy.min(), zz.min()]
        assert list(row[4:7]) == [xx.max() - xx.min() + 1, yy.max() - yy.min() + 1, zz.max() - zz.min() + 1]
