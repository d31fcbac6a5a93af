"""For a machine that HAS the reference (pip install itk itk-morphologicalcontourinterpolation): runs the real ITK
filter on this repo's fixtures and synthetic configurations and compares its output, voxel for voxel, with the
oracle (always) and with the CUDA path (when a GPU is present).  Not runnable in the build container or on the GPU
box of this project (ITK is absent there; the script says so and exits 0) -- it exists so that a maintainer can close
what tests/test_upstream_baselines.py leaves open with one command:

    python tools/itk_reference_check.py [--scale 0.25] [--dump out_dir]

Uses the reference's Python wrapping as examples/mciExample.py does (itk.MorphologicalContourInterpolator,
wrapping/itkMorphologicalContourInterpolator.wrap: 3-D UC / SS / US).
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def itk_run(itk, vol, use_distance_transform, use_ball, axis=-1, label=0):
    image = itk.image_from_array(np.ascontiguousarray(vol))  # numpy [z, y, x] -> itk image (x fastest)
    f = itk.MorphologicalContourInterpolator[type(image)].New()
    f.SetInput(image)
    f.SetUseDistanceTransform(bool(use_distance_transform))
    f.SetUseBallStructuringElement(bool(use_ball))
    f.SetAxis(int(axis))
    f.SetLabel(int(label))
    f.Update()
    return itk.array_from_image(f.GetOutput())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=0.25, help="scale of the synthetic C3 / C4 / C5 configurations")
    ap.add_argument("--dump", default=None, help="directory for the ITK outputs (.npy), to freeze as golden vectors")
    args = ap.parse_args()
    try:
        import itk
        itk.MorphologicalContourInterpolator  # noqa: B018  (remote module present?)
    except Exception as exc:  # ImportError / AttributeError
        print("ITK with the MorphologicalContourInterpolation module is not available here (%s); nothing checked." % exc)
        return 0
    import cases
    from oracle import oracle_lib as O
    from itkcontourinterpolation_b200.synthetic import make_config
    gpu = None
    try:
        from itkcontourinterpolation_b200 import morphological_contour_interpolator as gpu
    except Exception as exc:
        print("CUDA path not available (%s): comparing ITK with the oracle only" % exc)
    inputs = {"SevenLabels_u8": np.load(os.path.join(cases.GOLDEN, "SevenLabels.npy")).astype(np.uint8),
              "SevenLabels_short": np.load(os.path.join(cases.GOLDEN, "SevenLabels.npy")).astype(np.int16),
              "ManyToMany_short": np.load(os.path.join(cases.GOLDEN, "ManyToMany.npy")).astype(np.int16),
              "amygdala_short": cases.amygdala()}
    inputs.update({k: v for k, v in cases.handcrafted().items()})
    for name in ("C3", "C4", "C5"):
        inputs["synthetic_" + name] = make_config(name, args.scale)[0]
    bad = 0
    for name, vol in inputs.items():
        for algo, kw in cases.ALGOS.items():
            ref = itk_run(itk, vol, kw["use_distance_transform"], kw["use_ball"])
            if args.dump:
                os.makedirs(args.dump, exist_ok=True)
                np.save(os.path.join(args.dump, "%s_%s.npy" % (name, algo)), ref)
            try:
                ora = O.interpolate(vol, **kw)
                msg = "oracle %s" % ("==" if np.array_equal(ora, ref) else "DIFFERS in %d voxels" % int((ora != ref).sum()))
                bad += not np.array_equal(ora, ref)
            except O.OracleError as exc:
                msg = "oracle raised: %s" % exc
            if gpu is not None:
                out = gpu(vol, use_distance_transform=kw["use_distance_transform"], use_ball_structuring_element=kw["use_ball"])
                msg += "; cuda %s" % ("==" if np.array_equal(out, ref) else "DIFFERS in %d voxels" % int((out != ref).sum()))
                bad += not np.array_equal(out, ref)
            print("%-28s %s  %s" % (name, algo, msg))
    print("all equal" if bad == 0 else "%d comparisons differ" % bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
