"""MetaImage (.mha / .mhd) fixture I/O: the container of the reference's Dice-evaluation inputs
(doc/case_*_labels.csv, test/dscComparison.cxx).  CPU only."""
import os

import numpy as np
import pytest

from itkcontourinterpolation_b200.metaimage import metaimage_bytes, read_metaimage, write_metaimage


def _vol(dtype):
    rng = np.random.default_rng(7)
    v = np.zeros((6, 9, 11), dtype)
    v[1:5, 2:7, 3:9] = rng.integers(1, 7, (4, 5, 6))
    if np.dtype(dtype).kind == "i":
        v[0, 0, 0] = -3
    return v


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.int16])
@pytest.mark.parametrize("ext,compressed", [("mha", False), ("mha", True), ("mhd", False), ("mhd", True)])
def test_round_trip(tmp_path, dtype, ext, compressed):
    v = _vol(dtype)
    p = str(tmp_path / ("vol." + ext))
    write_metaimage(p, v, spacing=(0.5, 0.25, 2), origin=(1, -2, 3.5), compressed=compressed)
    back, hdr = read_metaimage(p, with_header=True)
    assert back.dtype == v.dtype and np.array_equal(back, v)
    assert hdr["DimSize"] == "11 9 6" and hdr["ElementSpacing"] == "0.5 0.25 2" and hdr["Offset"] == "1 -2 3.5"
    if ext == "mhd":
        assert os.path.exists(str(tmp_path / ("vol.zraw" if compressed else "vol.raw")))


def test_reads_a_hand_written_big_endian_header(tmp_path):
    v = _vol(np.uint16)
    hdr = ("ObjectType = Image\nNDims = 3\nBinaryData = True\nBinaryDataByteOrderMSB = True\nCompressedData = False\n"
           "TransformMatrix = 1 0 0 0 1 0 0 0 1\nOffset = 0 0 0\nElementSpacing = 1 1 1\nDimSize = 11 9 6\n"
           "ElementType = MET_USHORT\nElementDataFile = LOCAL\n").encode()
    p = tmp_path / "be.mha"
    p.write_bytes(hdr + v.astype(">u2").tobytes())
    assert np.array_equal(read_metaimage(str(p)), v)


def test_header_layout_and_rejections(tmp_path):
    v = _vol(np.int16)
    header, data = metaimage_bytes(v)
    lines = header.decode().splitlines()
    assert lines[0] == "ObjectType = Image" and lines[-1] == "ElementDataFile = LOCAL" and "ElementType = MET_SHORT" in lines
    assert len(data) == v.size * 2
    with pytest.raises(ValueError):
        metaimage_bytes(v.astype(np.float32))
    p = tmp_path / "f.mha"
    p.write_bytes(header.replace(b"MET_SHORT", b"MET_FLOAT") + data)
    with pytest.raises(ValueError):
        read_metaimage(str(p))
    p.write_bytes(header + data[:-4])
    with pytest.raises(ValueError):
        read_metaimage(str(p))
