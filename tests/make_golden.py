"""Regenerates tests/golden/: the two example label volumes of the reference (read from
/root/reference/examples/*.nrrd, available only in the build container) as .npy inputs, and
golden.json = sha256 of the ORACLE's output for a matrix of inputs x settings.

These goldens freeze the ORACLE restatement, not output of the real ITK filter (ITK is absent here).  The oracle
itself is pinned against four of the reference's own regression baselines by tests/test_upstream_baselines.py
(see DESIGN.md section 0).

    python tests/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import oracle_lib as O  # noqa: E402
import cases  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    gold = os.path.join(ROOT, "tests", "golden")
    os.makedirs(gold, exist_ok=True)
    ref = "/root/reference/examples"
    if os.path.isdir(ref):
        from itkcontourinterpolation_b200.nrrd import read_nrrd
        for f in ("SevenLabels", "ManyToMany"):
            np.save(os.path.join(gold, f + ".npy"), read_nrrd(os.path.join(ref, f + ".nrrd")))
    nii = "/root/reference/wasm/test/data/input/64816L_amygdala_int.nii.gz"
    if os.path.exists(nii):
        from itkcontourinterpolation_b200.nifti import read_nifti
        np.savez_compressed(os.path.join(gold, "64816L_amygdala_int.npz"), vol=read_nifti(nii))  # int16, 170x400x210
    table = {}
    inputs = {"SevenLabels": np.load(os.path.join(gold, "SevenLabels.npy")),
              "ManyToMany": np.load(os.path.join(gold, "ManyToMany.npy"))}
    inputs["SevenLabels_u8"] = inputs["SevenLabels"].astype(np.uint8)  # mciExample.cxx:19 reads it as uchar
    inputs.update(cases.handcrafted())
    for name, vol in inputs.items():
        for algo, kw in cases.ALGOS.items():
            out = O.interpolate(vol, **kw)
            table["%s_%s" % (name, algo)] = dict(input=sha(vol), output=sha(out), shape=list(vol.shape), dtype=str(vol.dtype))
    for name, (vol, cfg) in cases.synthetic_small().items():
        out = O.interpolate(vol, axis=cfg.get("axis", -1), use_distance_transform=cfg["use_dt"], use_ball=cfg["ball"])
        table["synthetic_%s" % name] = dict(input=sha(vol), output=sha(out), shape=list(vol.shape), dtype=str(vol.dtype))
    # the reference's regression matrix for 64816L_amygdala_int (test/CMakeLists.txt:143-160): algorithms B, C, T x
    # {all axes, axis a, axis a + label l}, read as Image<short,3> like its driver does (test/...Test.cxx:110-114)
    amy = np.load(os.path.join(gold, "64816L_amygdala_int.npz"))["vol"]
    for algo in "BCT":
        for axis, label in cases.AMYGDALA_ARGS:
            out = O.interpolate(amy, axis=axis, label=label, threads=os.cpu_count() or 1, **cases.ALGOS[algo])
            table["64816L_amygdala_int_%s_%d_%d" % (algo, axis, label)] = dict(input=sha(amy), output=sha(out),
                                                                                shape=list(amy.shape), dtype=str(amy.dtype))
    with open(os.path.join(gold, "golden.json"), "w") as f:
        json.dump(table, f, indent=1, sort_keys=True)
    print("wrote", len(table), "entries")


if __name__ == "__main__":
    main()
