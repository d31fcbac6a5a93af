"""Pins the oracle's restated ITK primitives against scipy / brute force.
CPU only."""
import numpy as np
import pytest
from scipy import ndimage

from oracle import oracle_lib as O

RNG = np.random.default_rng(1234)


def random_masks():
    for k in range(12):
        h, w = RNG.integers(1, 70, size=2)
        p = RNG.choice([0.1, 0.35, 0.5, 0.65, 0.9])
        yield RNG.random((h, w)) < p
    yield np.zeros((5, 7), bool)
    yield np.ones((6, 33), bool)
    m = np.zeros((40, 70), bool)
    m[5:30, 10:60] = True
    m[10:25, 20:50] = False  # ring
    yield m


@pytest.mark.parametrize("fully", [False, True])
def test_connected_components_match_scipy_raster_numbering(fully):
    # A.1: ids 1..n in raster order of each component's first pixel == scipy.ndimage.label numbering
    structure = np.ones((3, 3), int) if fully else None
    for m in random_masks():
        ids, n = O.cc2d(m, fully)
        ref, nref = ndimage.label(m, structure=structure)
        assert n == nref
        assert np.array_equal(ids, ref.astype(np.uint16))
        firsts = [np.flatnonzero(ids.ravel() == k)[0] for k in range(1, n + 1)]
        assert firsts == sorted(firsts)


def test_squared_edt_exact():
    # A.2: exact squared Euclidean distance to the nearest object pixel
    for m in random_masks():
        fast = O.edt2(m, brute=False)
        slow = O.edt2(m, brute=True)
        assert np.array_equal(fast, slow)
        if m.any():
            sp = ndimage.distance_transform_edt(~m) ** 2
            assert np.array_equal(fast, np.rint(sp).astype(np.int64))


def test_squared_edt_large_sparse():
    m = np.zeros((200, 300), bool)
    m[17, 250] = True
    m[180, 3] = True
    assert np.array_equal(O.edt2(m), O.edt2(m, brute=True))


@pytest.mark.parametrize("ball", [False, True])
def test_dilation_structuring_elements(ball):
    # A.3: radius-1 cross = 4-neighbourhood; radius-1 ITK ball = full 3x3 box; outside = background
    st = np.ones((3, 3), bool) if ball else ndimage.generate_binary_structure(2, 1)
    for m in random_masks():
        assert np.array_equal(O.dilate(m, ball).astype(bool), ndimage.binary_dilation(m, structure=st))


def _median_dilations_literal(i, j, ball):
    """FindMedianImageDilations restated independently with scipy (X:340-388)."""
    st = np.ones((3, 3), bool) if ball else ndimage.generate_binary_structure(2, 1)

    def seq(begin, end):
        s = [ndimage.binary_dilation(begin, structure=st) & end]
        while True:
            s.append(ndimage.binary_dilation(s[-1], structure=st) & end)
            if np.array_equal(s[-1], s[-2]):
                break
        s.pop()
        return s

    x = i & j
    iseq, jseq = seq(x, i), seq(x, j)
    iseq.reverse()
    if len(iseq) < len(jseq):
        iseq, jseq = jseq, iseq
    ratio = np.float32(len(jseq)) / np.float32(len(iseq))
    best, bi = i.size, 0
    for k in range(len(iseq)):
        u = iseq[k] | jseq[int(np.float32(ratio * np.float32(k)))]
        sc = abs(int((u != i).sum()) - int((u != j).sum()))
        if sc < best:
            best, bi = sc, k
    return iseq[bi] | jseq[int(np.float32(ratio * np.float32(bi)))]


@pytest.mark.parametrize("ball", [False, True])
def test_median_dilations_against_independent_restatement(ball):
    yy, xx = np.mgrid[:48, :56]
    shapes = [((yy - 24) ** 2 + (xx - 28) ** 2 <= 18 ** 2, (yy - 20) ** 2 + (xx - 24) ** 2 <= 7 ** 2),
              ((abs(yy - 24) < 15) & (abs(xx - 28) < 20), (yy - 24) ** 2 + (xx - 30) ** 2 <= 64),
              ((yy - 24) ** 2 / 4 + (xx - 28) ** 2 <= 100, (yy - 24) ** 2 + (xx - 28) ** 2 / 4 <= 100)]
    for i, j in shapes:
        for a, b in ((i, j), (j, i)):
            got = O.median(a, b, use_distance_transform=False, use_ball=ball).astype(bool)
            assert np.array_equal(got, _median_dilations_literal(a, b, ball))


def test_median_distances_properties():
    yy, xx = np.mgrid[:64, :64]
    i = (yy - 32) ** 2 + (xx - 32) ** 2 <= 24 ** 2
    j = (yy - 32) ** 2 + (xx - 32) ** 2 <= 10 ** 2
    m = O.median(i, j).astype(bool)
    assert (m >= (i & j)).all() and (m <= (i | j)).all()
    area = m.sum()
    assert j.sum() < area < i.sum()
    # identical shapes: the median is that shape (X:463-466)
    assert np.array_equal(O.median(i, i).astype(bool), i)


def test_distance_bin_wrap_depends_on_pixel_type():
    # A.5: `PixelType dist = 10 * sdf` wraps modulo 256 for uchar and modulo 65536 for ushort
    yy, xx = np.mgrid[:40, :220]
    i = (abs(yy - 20) < 18) & (abs(xx - 110) < 105)
    j = (abs(yy - 20) < 3) & (abs(xx - 110) < 4)
    m8 = O.median(i, j, dtype=np.uint8)
    m16 = O.median(i, j, dtype=np.uint16)
    assert not np.array_equal(m8, m16)
    assert np.array_equal(O.median(i, j, dtype=np.int16), m16)  # no wrap below 3276 pixels of distance
    xx = np.arange(3400)[None, :].repeat(5, 0)
    i = xx >= 0
    j = xx < 6
    with pytest.raises(O.OracleError):
        O.median(i, j, dtype=np.int16)  # distance >= 3276.8: negative bin, the reference indexes out of bounds
    assert O.median(i, j, dtype=np.uint16).any()


def _median_distances_literal(i, j, dtype):
    """FindMedianImageDistances restated independently with scipy (X:405-516): the signed Maurer map with
    SquaredDistance off is the Euclidean distance to the intersection for pixels outside it."""
    x = i & j
    sdf = ndimage.distance_transform_edt(~x).astype(np.float32)  # float(sqrt(double(d2)))
    bins = (np.float32(10) * sdf).astype(np.int32).astype(dtype).astype(np.int64)
    ih = np.bincount(bins[i & ~j], minlength=1) if (i & ~j).any() else np.zeros(0, np.int64)
    jh = np.bincount(bins[j & ~i], minlength=1) if (j & ~i).any() else np.zeros(0, np.int64)
    n = max(len(ih), len(jh))
    if n == 0:
        return x
    ih = np.pad(ih, (0, n - len(ih)))
    jh = np.pad(jh, (0, n - len(jh)))
    isum, jsum = np.cumsum(ih), np.cumsum(jh)
    diff = np.abs(np.abs(isum[-1] - isum + jsum) - np.abs(jsum[-1] - jsum + isum))
    best = int(np.argmin(diff))  # first minimum
    upper = np.float32(best) / np.float32(10)
    return ((sdf <= upper) | x) & (i | j)


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.int16])
def test_median_distances_against_independent_restatement(dtype):
    yy, xx = np.mgrid[:48, :90]
    shapes = [((yy - 24) ** 2 + (xx - 28) ** 2 <= 18 ** 2, (yy - 20) ** 2 + (xx - 24) ** 2 <= 7 ** 2),
              ((abs(yy - 24) < 15) & (abs(xx - 45) < 40), (yy - 24) ** 2 + (xx - 30) ** 2 <= 64),
              ((yy - 24) ** 2 / 4 + (xx - 28) ** 2 <= 100, (yy - 24) ** 2 + (xx - 28) ** 2 / 4 <= 100)]
    for i, j in shapes:
        for a, b in ((i, j), (j, i)):
            got = O.median(a, b, dtype=dtype).astype(bool)
            assert np.array_equal(got, _median_distances_literal(a, b, dtype))


def test_float_sqrt_equals_double_sqrt_rounded_to_float():
    # The CUDA path bins with __fsqrt_rn(float(d2)); ITK (and the oracle) compute float(std::sqrt(double(float(d2)))).
    # The two agree for every d2 an 8191-pixel-wide region can produce, and so do the bins 10 * sdf.
    for lo in range(0, 1 << 25, 1 << 22):
        d = np.arange(lo, lo + (1 << 22), dtype=np.int64).astype(np.float32)
        a = np.sqrt(d)
        b = np.sqrt(d.astype(np.float64)).astype(np.float32)
        assert np.array_equal(a, b)
        assert np.array_equal((np.float32(10) * a).astype(np.int32), (np.float32(10) * b).astype(np.int32))
