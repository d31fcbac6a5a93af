"""Small runs of every kernel family for compute-sanitizer:
    compute-sanitizer --tool memcheck  python tools/sanitize_cases.py
    compute-sanitizer --tool racecheck python tools/sanitize_cases.py small
Cases: the reference's ManyToMany (1-to-N splits, Align with several components) under T / B / C, SevenLabels (seven
labels, extrapolation), handcrafted OneToThree / RingExtrapolate / ThreeAxis (slices along x and y), a scaled C3, a
region larger than the shared-memory budgets (global-scratch paths; skipped with `small`), the median smoothing step.
Every output is compared with the oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from itkcontourinterpolation_b200 import MorphologicalContourInterpolator  # noqa: E402
from itkcontourinterpolation_b200.synthetic import make_config  # noqa: E402
from oracle import oracle_lib as O  # noqa: E402


def main():
    small = len(sys.argv) > 1 and sys.argv[1] == "small"
    gold = os.path.join(ROOT, "tests", "golden")
    vols = [("ManyToMany", np.load(os.path.join(gold, "ManyToMany.npy")), "TBC"),
            ("SevenLabels", np.load(os.path.join(gold, "SevenLabels.npy")), "TBCD")]
    h = cases.handcrafted()
    for name in ("OneToThree", "RingExtrapolate", "ThreeAxis", "DriftOddGaps"):
        vols.append((name, h[name], "DC"))
    vols.append(("C3/8", make_config("C3", 0.125)[0], "D"))
    if not small:
        import test_gpu_parity as P
        vols.append(("BigROI", P._big_roi_volume(700), "DC"))
    f = MorphologicalContourInterpolator(0)
    bad = 0
    try:
        for name, vol, algos in vols:
            for a in algos:
                kw = cases.ALGOS[a]
                f.SetUseDistanceTransform(kw["use_distance_transform"])
                f.SetUseBallStructuringElement(kw["use_ball"])
                f.SetInput(vol)
                out = f.Update()
                ok = np.array_equal(out, O.interpolate(vol, **kw))
                bad += not ok
                print("%-14s %s %s launches=%d" % (name, a, "bit-exact" if ok else "MISMATCH", f.last_stats["n_launches"]), flush=True)
    finally:
        f.close()
    from itkcontourinterpolation_b200 import median_image_filter
    vol = make_config("C3", 0.125)[0]
    ok = np.array_equal(median_image_filter(vol, 2), O.median_image_filter(vol, 2))
    bad += not ok
    print("median r=2     %s" % ("bit-exact" if ok else "MISMATCH"), flush=True)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
