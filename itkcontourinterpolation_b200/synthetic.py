"""Synthetic sparse-annotated label volumes.

Dense ground truth = solid ellipsoids painted in label order (later labels overwrite earlier ones),
then sparsified like test/dscComparison.cxx:32-71 (`createSparseCopy`): a voxel survives iff its index
along some axis a with nth[a] > 0 is a multiple of nth[a].  Pure numpy, deterministic from `seed`.
"""
import random

import numpy as np


def ellipsoid_volume(dims, n_labels, seed, semi_axes=(0.08, 0.20), dtype=np.uint16, branching=0, z_range=None):
    """dims = (X, Y, Z).  Returns the dense volume indexed [z, y, x]; with z_range = (lo, hi) only the planes
    lo <= z < hi of the same volume (a rank of a slab-sharded run generates its window only)."""
    X, Y, Z = dims
    rng = random.Random(seed)
    wlo, whi = (0, Z) if z_range is None else (max(0, int(z_range[0])), min(Z, int(z_range[1])))
    vol = np.zeros((max(0, whi - wlo), Y, X), dtype=dtype)

    def paint(cx, cy, cz, rx, ry, rz, label):
        x0, x1 = max(0, int(cx - rx)), min(X, int(cx + rx) + 2)
        y0, y1 = max(0, int(cy - ry)), min(Y, int(cy + ry) + 2)
        z0, z1 = max(wlo, int(cz - rz)), min(whi, int(cz + rz) + 2)
        if x0 >= x1 or y0 >= y1 or z0 >= z1:
            return
        zz = ((np.arange(z0, z1) - cz) / rz) ** 2
        yy = ((np.arange(y0, y1) - cy) / ry) ** 2
        xx = ((np.arange(x0, x1) - cx) / rx) ** 2
        inside = (zz[:, None, None] + yy[None, :, None] + xx[None, None, :]) <= 1.0
        sub = vol[z0 - wlo:z1 - wlo, y0:y1, x0:x1]
        sub[inside] = label

    for label in range(1, n_labels + 1):
        cx, cy, cz = rng.uniform(0, X), rng.uniform(0, Y), rng.uniform(0, Z)
        rx, ry, rz = (rng.uniform(*semi_axes) * d for d in (X, Y, Z))
        paint(cx, cy, cz, rx, ry, rz, label)
        if label <= branching:
            # a second lobe that merges with the first along z: exercises 1-to-N and Extrapolate
            paint(cx + 0.9 * rx, cy, cz + 0.6 * rz, 0.6 * rx, 0.6 * ry, 0.8 * rz, label)
    return vol


def sparsify(vol, nth, z_first=0):
    """nth = (nx, ny, nz); 0 means 'do not keep slices along this axis'.  z_first: z of vol[0] (windows)."""
    Z, Y, X = vol.shape
    if nth[0] <= 0 and nth[1] <= 0 and nth[2] > 0:
        out = np.zeros_like(vol)  # z-slices only: no full-size mask needed
        k0 = (-z_first) % nth[2]
        out[k0::nth[2]] = vol[k0::nth[2]]
        return out
    keep = np.zeros(vol.shape, dtype=bool)
    if nth[0] > 0:
        keep[:, :, ::nth[0]] = True
    if nth[1] > 0:
        keep[:, ::nth[1], :] = True
    if nth[2] > 0:
        keep[(-z_first) % nth[2]::nth[2], :, :] = True
    out = np.where(keep, vol, 0).astype(vol.dtype)
    return out


CONFIGS = {
    "C3": dict(dims=(512, 512, 256), n_labels=20, nth=(0, 0, 8), seed=20240607 + 3, axis=2, use_dt=True, ball=False,
                   semi_axes=(0.08, 0.20)),
    "C4": dict(dims=(1024, 1024, 512), n_labels=20, nth=(16, 16, 8), seed=20240607 + 4, axis=-1, use_dt=False, ball=True,
                   semi_axes=(0.08, 0.20)),
    "C5": dict(dims=(2048, 2048, 1024), n_labels=200, nth=(0, 0, 8), seed=20240607 + 5, axis=2, use_dt=True, ball=False,
                   semi_axes=(0.03, 0.08)),
}


def make_config(name, scale=1.0, z_range=None, want_dense=True):
    """The synthetic configs of BASELINE.json (C3, C4, C5), optionally scaled down for tests.  z_range = (lo, hi):
    only those planes of the volume (same voxels as the corresponding planes of the whole)."""
    cfg = CONFIGS[name]
    dims = tuple(max(8, int(round(d * scale))) for d in cfg["dims"])
    dense = ellipsoid_volume(dims, cfg["n_labels"], cfg["seed"], cfg["semi_axes"], z_range=z_range)
    sparse = sparsify(dense, cfg["nth"], z_first=0 if z_range is None else max(0, int(z_range[0])))
    return sparse, (dense if want_dense else None), dict(cfg, dims=dims)
