"""Freezes the ORACLE's output for the full-size synthetic configurations of BASELINE.json (C3, C4, C5) and for a
half-scale C5, as sha256 of the voxel array plus one CRC-32 per z-plane (so that a rank of a slab-sharded run can
check just its own planes).  bench.py and the GPU tests compare against tests/golden/large.json instead of re-running
the oracle on 4.3 Gvoxel (14 s on 16 threads, more than a minute on a small box).

    python tests/make_golden_large.py [C3 C4 C5 C5_half]      (needs ~30 GB of RAM for C5)
"""
import hashlib
import json
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from itkcontourinterpolation_b200.synthetic import make_config  # noqa: E402
from oracle import oracle_lib as O  # noqa: E402

CASES = {"C3": ("C3", 1.0), "C4": ("C4", 1.0), "C5": ("C5", 1.0), "C5_half": ("C5", 0.5)}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).data).hexdigest()


def plane_crcs(a):
    return [zlib.crc32(np.ascontiguousarray(a[z]).data) for z in range(a.shape[0])]


def main():
    path = os.path.join(ROOT, "tests", "golden", "large.json")
    table = json.load(open(path)) if os.path.exists(path) else {}
    for key in (sys.argv[1:] or list(CASES)):
        name, scale = CASES[key]
        vol, _, cfg = make_config(name, scale, want_dense=False)
        out = O.interpolate(vol, axis=cfg["axis"], use_distance_transform=cfg["use_dt"], use_ball=cfg["ball"],
                            threads=int(os.environ.get("MCI_ORACLE_THREADS", os.cpu_count() or 1)))
        table[key] = dict(config=name, scale=scale, shape=list(vol.shape), dtype=str(vol.dtype), input=sha(vol), output=sha(out),
                          input_plane_crc32=plane_crcs(vol), output_plane_crc32=plane_crcs(out),
                          changed_voxels=int(np.count_nonzero(out != vol)))
        print(key, table[key]["shape"], table[key]["output"], table[key]["changed_voxels"], flush=True)
        del vol, out
        with open(path, "w") as f:
            json.dump(table, f, sort_keys=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
