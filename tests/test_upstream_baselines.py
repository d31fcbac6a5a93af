"""Known-answer tests against the reference's OWN regression baselines (test/Baseline/*.nrrd.sha512).

tests/make_upstream_pins.py explains how: the two example volumes are bit-identical to the reference's test inputs
(their raw-encoded NRRD reproduces test/Input/*.sha512), and an output volume written through ITK's NRRD byte layout
must hash to the baseline the real ITK filter produced.  The reference's driver reads the volume as Image<short,3>
and maps algorithm letters as: B = dilations + ball, C = dilations + cross, T = distance transform with the ball flag
left ON (8-connected components) -- test/itkMorphologicalContourInterpolationTest.cxx:69-87.

Reproduced: SevenLabels_B, ManyToMany_B, ManyToMany_C, ManyToMany_T.
Not reproduced: SevenLabels_C and SevenLabels_T.  The B run of the same input matches, and it makes the same calls
(same components, pairs, alignments, extrapolation, splits); only the median differs, and the C / T medians are
pinned by ManyToMany.  ITK's regression compare accepts label differences up to 2 (its default intensity
tolerance), so these two baselines may predate changes the upstream tests could not see; they are kept here as
expected failures so that any future change that reproduces them is noticed.
"""
import hashlib
import json
import os

import numpy as np
import pytest

import cases
from itkcontourinterpolation_b200.nrrd import nrrd_bytes
from oracle import oracle_lib as O

PINS = json.load(open(os.path.join(cases.GOLDEN, "upstream_baselines.json")))
DRIVER_ALGOS = {"B": dict(use_distance_transform=False, use_ball=True),
                "C": dict(use_distance_transform=False, use_ball=False),
                "T": dict(use_distance_transform=True, use_ball=True)}
REPRODUCED = [("SevenLabels", "B"), ("ManyToMany", "B"), ("ManyToMany", "C"), ("ManyToMany", "T")]
NOT_REPRODUCED = [("SevenLabels", "C"), ("SevenLabels", "T")]


def file_hash(name, vol):
    p = PINS[name]
    return hashlib.sha512(nrrd_bytes(vol, p["space_directions"], p["space_origin"], **p["writer"])).hexdigest()


def load_as_short(name):
    return np.load(os.path.join(cases.GOLDEN, name + ".npy")).astype(np.int16)  # doTest<itk::Image<short, 3>>


@pytest.mark.parametrize("name", sorted(PINS))
def test_fixture_is_the_upstream_test_input(name):
    p = PINS[name]
    vol = np.load(os.path.join(cases.GOLDEN, name + ".npy"))
    assert hashlib.sha512(nrrd_bytes(vol, p["space_directions"], p["space_origin"], encoding="raw")).hexdigest() == p["input_sha512"]


@pytest.mark.parametrize("name,algo", REPRODUCED)
def test_oracle_reproduces_upstream_baseline(name, algo):
    out = O.interpolate(load_as_short(name), **DRIVER_ALGOS[algo])
    assert file_hash(name, out) == PINS[name]["baselines"][algo]["sha512"]


@pytest.mark.parametrize("name,algo", NOT_REPRODUCED)
@pytest.mark.xfail(strict=True, reason="baseline not reproduced (see module docstring)")
def test_oracle_vs_unreproduced_upstream_baseline(name, algo):
    out = O.interpolate(load_as_short(name), **DRIVER_ALGOS[algo])
    assert file_hash(name, out) == PINS[name]["baselines"][algo]["sha512"]


@pytest.mark.gpu
@pytest.mark.parametrize("name,algo", REPRODUCED)
def test_cuda_path_reproduces_upstream_baseline(name, algo):
    from itkcontourinterpolation_b200 import MorphologicalContourInterpolator
    f = MorphologicalContourInterpolator(0)
    try:
        kw = DRIVER_ALGOS[algo]
        f.SetUseDistanceTransform(kw["use_distance_transform"])
        f.SetUseBallStructuringElement(kw["use_ball"])
        f.SetInput(load_as_short(name))
        out = f.Update()
    finally:
        f.close()
    assert file_hash(name, out) == PINS[name]["baselines"][algo]["sha512"]
