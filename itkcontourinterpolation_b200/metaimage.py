"""MetaImage (.mha / .mhd) reader and writer for label volumes.

Fixture I/O only: the volumes the reference's Dice evaluation reads
(`doc/case_*_labels.csv` name `.mha` inputs, `test/dscComparison.cxx`) are MetaImage files.  The subset ITK's
MetaImageIO writes for a 3-D scalar label image is covered: ElementType MET_UCHAR / MET_USHORT / MET_SHORT, local
(`ElementDataFile = LOCAL`) or external raw data, optional zlib compression (`CompressedData = True`), either byte
order.  Arrays are indexed [z, y, x] like everywhere else in this package.
"""
import os
import zlib

import numpy as np

_TYPES = {"MET_UCHAR": np.uint8, "MET_USHORT": np.uint16, "MET_SHORT": np.int16}
_NAMES = {np.dtype(np.uint8): "MET_UCHAR", np.dtype(np.uint16): "MET_USHORT", np.dtype(np.int16): "MET_SHORT"}


def _true(v):
    return v.strip().lower() in ("true", "1", "yes")


def read_metaimage(path, with_header=False):
    raw = open(path, "rb").read()
    fields = {}
    pos = 0
    while True:
        end = raw.find(b"\n", pos)
        if end < 0:
            raise ValueError("%s: MetaImage header without ElementDataFile" % path)
        line = raw[pos:end].decode("ascii", "replace").rstrip("\r")
        pos = end + 1
        if "=" not in line:
            continue
        k, v = line.split("=", 1)
        fields[k.strip()] = v.strip()
        if k.strip() == "ElementDataFile":  # always the last header line
            break
    if fields.get("ObjectType", "Image") != "Image":
        raise ValueError("%s: ObjectType %s is not an image" % (path, fields.get("ObjectType")))
    ndims = int(fields["NDims"])
    sizes = [int(s) for s in fields["DimSize"].split()]
    if ndims != 3 or len(sizes) != 3:
        raise ValueError("%s: only 3-D volumes are supported (NDims = %d)" % (path, ndims))
    if int(fields.get("ElementNumberOfChannels", "1")) != 1:
        raise ValueError("%s: only scalar pixels are supported" % path)
    et = fields["ElementType"]
    if et not in _TYPES:
        raise ValueError("%s: unsupported ElementType %s (label volumes are MET_UCHAR / MET_USHORT / MET_SHORT)" % (path, et))
    dtype = np.dtype(_TYPES[et])
    msb = _true(fields.get("BinaryDataByteOrderMSB", fields.get("ElementByteOrderMSB", "False")))
    if msb:
        dtype = dtype.newbyteorder(">")
    if not _true(fields.get("BinaryData", "True")):
        raise ValueError("%s: ASCII element data is not supported" % path)
    src = fields["ElementDataFile"]
    if src == "LOCAL":
        data = raw[pos:]
    else:
        data = open(os.path.join(os.path.dirname(os.path.abspath(path)), src), "rb").read()
        skip = int(fields.get("HeaderSize", "0"))
        if skip > 0:
            data = data[skip:]
    if _true(fields.get("CompressedData", "False")):
        n = int(fields.get("CompressedDataSize", len(data)))
        data = zlib.decompress(data[:n])
    count = int(np.prod(sizes))
    if len(data) < count * dtype.itemsize:
        raise ValueError("%s: element data is %d bytes, %d expected" % (path, len(data), count * dtype.itemsize))
    arr = np.frombuffer(data, dtype=dtype, count=count).reshape(sizes[::-1])
    arr = np.ascontiguousarray(arr.astype(dtype.newbyteorder("=")))
    return (arr, fields) if with_header else arr


def metaimage_bytes(vol, spacing=(1.0, 1.0, 1.0), origin=(0.0, 0.0, 0.0), compressed=False, data_file="LOCAL"):
    """header (+ data when data_file == "LOCAL") in the field order ITK's MetaImageIO uses; returns (header, data)"""
    vol = np.ascontiguousarray(vol)
    if vol.ndim != 3 or vol.dtype not in _NAMES:
        raise ValueError("3-D uint8 / uint16 / int16 volumes only")
    data = vol.astype(vol.dtype.newbyteorder("<")).tobytes()

    def fmt(v):
        return " ".join(("%.17g" % float(x)) if float(x) != int(float(x)) else str(int(float(x))) for x in v)

    lines = ["ObjectType = Image", "NDims = 3", "BinaryData = True", "BinaryDataByteOrderMSB = False",
             "CompressedData = %s" % ("True" if compressed else "False")]
    if compressed:
        data = zlib.compress(data, 2)  # MetaIO's default compression level
        lines.append("CompressedDataSize = %d" % len(data))
    lines += ["TransformMatrix = 1 0 0 0 1 0 0 0 1", "Offset = " + fmt(origin), "CenterOfRotation = 0 0 0",
              "AnatomicalOrientation = RAI", "ElementSpacing = " + fmt(spacing),
              "DimSize = %d %d %d" % (vol.shape[2], vol.shape[1], vol.shape[0]), "ElementType = " + _NAMES[vol.dtype],
              "ElementDataFile = " + data_file]
    return ("\n".join(lines) + "\n").encode("ascii"), data


def write_metaimage(path, vol, spacing=(1.0, 1.0, 1.0), origin=(0.0, 0.0, 0.0), compressed=False):
    """.mha: header and data in one file; .mhd: header, with the data next to it in .raw / .zraw"""
    if path.lower().endswith(".mhd"):
        stem = os.path.splitext(os.path.basename(path))[0]
        data_file = stem + (".zraw" if compressed else ".raw")
        header, data = metaimage_bytes(vol, spacing, origin, compressed, data_file)
        with open(path, "wb") as f:
            f.write(header)
        with open(os.path.join(os.path.dirname(os.path.abspath(path)), data_file), "wb") as f:
            f.write(data)
    else:
        header, data = metaimage_bytes(vol, spacing, origin, compressed)
        with open(path, "wb") as f:
            f.write(header + data)
