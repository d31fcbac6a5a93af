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


_NAMES = {np.dtype(np.uint8): "unsigned char", np.dtype(np.uint16): "unsigned short", np.dtype(np.int16): "short"}


def _g17(v):
    return "%.17g" % float(v)


def nrrd_bytes(arr, space_directions, space_origin, encoding="gzip", zlib_level=-1, gz_os_code=3):
    """The bytes itk::ImageFileWriter (NrrdImageIO, i.e. teem's nrrdSave) writes for a 3-D scalar volume
    `arr` indexed [z, y, x]: teem's header layout and number formatting ("%.17g"), and for encoding "gzip" teem's
    own gzip framing (10-byte header without name or time stamp, OS byte `gz_os_code` = 3 on unix / 11 on Windows,
    raw deflate at `zlib_level` with memLevel 8, CRC-32 and length trailer).  Written so that the content hashes
    of the reference's regression baselines (test/Baseline/*.nrrd.sha512, produced by its test driver with
    UseCompression on, test/itkMorphologicalContourInterpolationTest.cxx:44-48) can be reproduced from an output
    volume; see tests/test_upstream_baselines.py."""
    import zlib
    arr = np.ascontiguousarray(arr)
    assert arr.ndim == 3 and arr.dtype in _NAMES
    raw = arr.astype(arr.dtype.newbyteorder("<")).tobytes()
    vec = lambda v: "(" + ",".join(_g17(c) for c in v) + ")"
    lines = ["NRRD0004",
             "# Complete NRRD file format specification at:",
             "# http://teem.sourceforge.net/nrrd/format.html",
             "type: " + _NAMES[arr.dtype],
             "dimension: 3",
             "space: left-posterior-superior",
             "sizes: %d %d %d" % (arr.shape[2], arr.shape[1], arr.shape[0]),
             "space directions: " + " ".join(vec(d) for d in space_directions),
             "kinds: domain domain domain"]
    if arr.dtype.itemsize > 1:
        lines.append("endian: little")
    lines += ["encoding: " + encoding, "space origin: " + vec(space_origin)]
    head = ("\n".join(lines) + "\n\n").encode("ascii")
    if encoding == "raw":
        return head + raw
    if encoding != "gzip":
        raise ValueError("unsupported NRRD encoding " + encoding)
    c = zlib.compressobj(zlib_level, zlib.DEFLATED, -15, 8, 0)
    body = c.compress(raw) + c.flush()
    return (head + bytes([0x1F, 0x8B, 8, 0, 0, 0, 0, 0, 0, gz_os_code]) + body +
            zlib.crc32(raw).to_bytes(4, "little") + (len(raw) & 0xFFFFFFFF).to_bytes(4, "little"))


def write_nrrd(path, arr, space_directions=((1, 0, 0), (0, 1, 0), (0, 0, 1)), space_origin=(0, 0, 0), **kw):
    with open(path, "wb") as f:
        f.write(nrrd_bytes(arr, space_directions, space_origin, **kw))
