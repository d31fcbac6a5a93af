"""Minimal NIfTI-1 reader (.nii / .nii.gz, single file) for the reference's third fixture,
wasm/test/data/input/64816L_amygdala_int.nii.gz -- also an input of its regression suite
(test/CMakeLists.txt:143-160).  Fixture I/O only; returns an array indexed [z, y, x].
"""
import gzip
import struct

import numpy as np

_TYPES = {2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 256: np.int8, 512: np.uint16, 768: np.uint32}


def read_nifti(path):
    raw = open(path, "rb").read()
    if raw[:2] == b"\x1f\x8b":
        raw = gzip.decompress(raw)
    if len(raw) < 348:
        raise ValueError("%s: not a NIfTI-1 file" % path)
    end = "<" if struct.unpack("<i", raw[:4])[0] == 348 else ">"
    if struct.unpack(end + "i", raw[:4])[0] != 348 or raw[344:347] not in (b"n+1", b"ni1"):
        raise ValueError("%s: not a NIfTI-1 file" % path)
    if raw[344:347] == b"ni1":
        raise ValueError("%s: header/image pairs (.hdr/.img) are not supported" % path)
    dim = struct.unpack(end + "8h", raw[40:56])
    datatype = struct.unpack(end + "h", raw[70:72])[0]
    vox_offset = int(struct.unpack(end + "f", raw[108:112])[0])
    if datatype not in _TYPES:
        raise ValueError("%s: unsupported NIfTI datatype %d" % (path, datatype))
    ndim = dim[0]
    sizes = [int(d) for d in dim[1:1 + ndim]]
    while len(sizes) > 3 and sizes[-1] == 1:
        sizes.pop()
    if len(sizes) != 3:
        raise ValueError("%s: expected a 3-D volume, got sizes %r" % (path, sizes))
    dtype = np.dtype(_TYPES[datatype]).newbyteorder(end)
    arr = np.frombuffer(raw, dtype=dtype, count=int(np.prod(sizes)), offset=vox_offset)
    return np.ascontiguousarray(arr.reshape(sizes[::-1]).astype(dtype.newbyteorder("=")))
