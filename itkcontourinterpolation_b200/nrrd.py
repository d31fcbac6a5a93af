"""Minimal NRRD reader for the two example label volumes (gzip / raw encodings).

Fixture I/O only; returns an array indexed [z, y, x].
"""
import gzip

import numpy as np

_TYPES = {
    "unsigned char": np.uint8, "uchar": np.uint8, "uint8": np.uint8,
    "unsigned short": np.uint16, "ushort": np.uint16, "uint16": np.uint16,
    "short": np.int16, "signed short": np.int16, "int16": np.int16,
}


def read_nrrd(path):
    raw = open(path, "rb").read()
    cut = raw.find(b"\n\n")
    if not raw.startswith(b"NRRD") or cut < 0:
        raise ValueError("%s: not a NRRD file" % path)
    fields = {}
    for line in raw[:cut].decode("ascii", "replace").splitlines()[1:]:
        if line.startswith("#") or ":" not in line:
            continue
        k, v = line.split(":", 1)
        fields[k.strip()] = v.strip().lstrip("=").strip()
    dtype = np.dtype(_TYPES[fields["type"]])
    if fields.get("endian", "little") != "little":
        dtype = dtype.newbyteorder(">")
    sizes = [int(s) for s in fields["sizes"].split()]
    data = raw[cut + 2:]
    enc = fields.get("encoding", "raw")
    if enc in ("gzip", "gz"):
        data = gzip.decompress(data)
    elif enc != "raw":
        raise ValueError("unsupported NRRD encoding " + enc)
    arr = np.frombuffer(data, dtype=dtype, count=int(np.prod(sizes)))
    return np.ascontiguousarray(arr.reshape(sizes[::-1]).astype(dtype.newbyteorder("=")))
