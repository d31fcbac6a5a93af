"""Collects, from /root/reference (available only in the build container), what is needed to check outputs against
the reference's OWN regression baselines without the reference tree: tests/golden/upstream_baselines.json.

The reference's test inputs and baselines are `.sha512` content links (test/Input, test/Baseline) whose payloads are
not available offline.  For two inputs the payload can be rebuilt: examples/SevenLabels.nrrd and
examples/ManyToMany.nrrd hold the same voxels and geometry as test/Input/{SevenLabels,ManyToMany}.nrrd -- re-encoding
them the way ITK's NRRD writer does ("raw" encoding) reproduces the sha512 of the test inputs exactly (checked
below).  The baselines are what the reference's driver wrote for those inputs (Image<short,3>, UseCompression on,
test/itkMorphologicalContourInterpolationTest.cxx:25-49,110-114): writing OUR output volume through the same
byte layout (itkcontourinterpolation_b200.nrrd.nrrd_bytes) must give the baseline's sha512.  A hash match is a
voxel-exact match of the whole output volume against output of the real ITK filter.

    python tests/make_upstream_pins.py
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from itkcontourinterpolation_b200.nrrd import nrrd_bytes, read_nrrd  # noqa: E402

REF = "/root/reference"
GEOMETRY = {"SevenLabels": ([(2, 0, 0), (0, 5, 0), (0, 0, 3)], (0, 0, 0)),
            "ManyToMany": ([(1, 0, 0), (0, 1, 0), (0, 0, 4)], (0, 0, 0))}


def link(rel):
    return open(os.path.join(REF, rel)).read().strip()


def main():
    table = {}
    for name, (dirs, origin) in GEOMETRY.items():
        vol = read_nrrd(os.path.join(REF, "examples", name + ".nrrd"))
        rebuilt = hashlib.sha512(nrrd_bytes(vol, dirs, origin, encoding="raw")).hexdigest()
        upstream = link("test/Input/%s.nrrd.sha512" % name)
        assert rebuilt == upstream, name  # the example file IS the test input
        table[name] = dict(space_directions=dirs, space_origin=origin,
                           input_link="test/Input/%s.nrrd.sha512" % name, input_sha512=upstream,
                           # test/CMakeLists.txt:40-61: algorithms B, C, T, all axes, all labels
                           baselines={a: dict(link="test/Baseline/%s_%s.nrrd.sha512" % (name, a),
                                              sha512=link("test/Baseline/%s_%s.nrrd.sha512" % (name, a))) for a in "BCT"},
                           # how the baseline files were written (found by reproducing them): teem gzip framing,
                           # zlib default level, unix OS byte
                           writer=dict(encoding="gzip", zlib_level=-1, gz_os_code=3))
    with open(os.path.join(ROOT, "tests", "golden", "upstream_baselines.json"), "w") as f:
        json.dump(table, f, indent=1, sort_keys=True)
    print("wrote upstream_baselines.json")


if __name__ == "__main__":
    main()
