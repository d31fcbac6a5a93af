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


@pytest.fixture(scope="module")
def driver():
    from itkcontourinterpolation_b200 import build as B
    return B.build_cpp_driver()


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.int16])
@pytest.mark.parametrize("src_ext,src_compressed", [("mha", False), ("mha", True), ("mhd", True)])
@pytest.mark.parametrize("dst_ext", ["mha", "mhd", "nrrd"])
def test_cpp_reader_and_writer_agree_with_the_python_ones(driver, tmp_path, dtype, src_ext, src_compressed, dst_ext):
    """mci_filter_test --copy (no GPU): the C++ ImageFileReader / ImageFileWriter of include/mci_b200_io.hpp read what
    the Python writer wrote and write what the Python reader reads -- Image<short,3> in between, as the reference's
    driver instantiates it, so the values must fit a short"""
    import subprocess
    from itkcontourinterpolation_b200.nrrd import read_nrrd
    v = _vol(dtype)
    src, dst = str(tmp_path / ("in." + src_ext)), str(tmp_path / ("out." + dst_ext))
    write_metaimage(src, v, spacing=(0.5, 2, 3), origin=(1, -2, 3.5), compressed=src_compressed)
    r = subprocess.run([driver, "--copy", src, dst], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    if dst_ext == "nrrd":
        back = read_nrrd(dst)
    else:
        back, hdr = read_metaimage(dst, with_header=True)
        assert hdr["ElementType"] == "MET_SHORT" and hdr["CompressedData"] == "True"
        assert hdr["ElementSpacing"] == "0.5 2 3" and hdr["Offset"] == "1 -2 3.5" and hdr["DimSize"] == "11 9 6"
        assert hdr["TransformMatrix"] == "1 0 0 0 1 0 0 0 1"
        # byte for byte what the Python writer produces for the same volume
        want_header, want_data = metaimage_bytes(v.astype(np.int16), (0.5, 2, 3), (1, -2, 3.5), True,
                                                 "LOCAL" if dst_ext == "mha" else "out.zraw")
        got = open(dst, "rb").read()
        if dst_ext == "mha":
            assert got == want_header + want_data
        else:
            assert got == want_header and open(str(tmp_path / "out.zraw"), "rb").read() == want_data
    assert back.dtype == np.int16 and np.array_equal(back, v.astype(np.int16))


def test_cpp_reader_rejects_what_it_does_not_cover(driver, tmp_path):
    import subprocess
    v = _vol(np.uint16)
    header, data = metaimage_bytes(v)
    bad = tmp_path / "f.mha"
    bad.write_bytes(header.replace(b"MET_USHORT", b"MET_FLOAT") + data)
    r = subprocess.run([driver, "--copy", str(bad), str(tmp_path / "o.mha")], capture_output=True, text=True, timeout=120)
    assert r.returncode != 0 and "ElementType" in r.stderr
    bad.write_bytes(header + data[:-6])
    r = subprocess.run([driver, "--copy", str(bad), str(tmp_path / "o.mha")], capture_output=True, text=True, timeout=120)
    assert r.returncode != 0
