"""Parity of the CUDA path against the oracle, through the C ABI (ctypes -> libmci_b200.so).
Bit-exact: integer / label work, no tolerance.  Needs a B200: pytest -m gpu."""
import ctypes
import hashlib
import json
import os

import numpy as np
import pytest

import cases
from itkcontourinterpolation_b200 import MorphologicalContourInterpolator, _lib as L
from itkcontourinterpolation_b200.synthetic import make_config
from oracle import oracle_lib as O

pytestmark = pytest.mark.gpu

GOLD = json.load(open(os.path.join(cases.GOLDEN, "golden.json")))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def flt():
    f = MorphologicalContourInterpolator(0)
    yield f
    f.close()


def run_gpu(f, vol, label=0, axis=-1, use_distance_transform=True, use_ball=False, heuristic=True, extrapolation=True):
    f.SetLabel(label)
    f.SetAxis(axis)
    f.SetUseDistanceTransform(use_distance_transform)
    f.SetUseBallStructuringElement(use_ball)
    f.SetHeuristicAlignment(heuristic)
    f.SetUseExtrapolation(extrapolation)
    f.SetUseCustomSlicePositions(False)
    f.SetInput(vol)
    return f.Update()


def assert_same(gpu, ref, what):
    if not np.array_equal(gpu, ref):
        bad = np.argwhere(gpu != ref)
        z, y, x = bad[0]
        raise AssertionError("%s: %d voxels differ, first at (x=%d,y=%d,z=%d): gpu=%d oracle=%d" %
                             (what, len(bad), x, y, z, gpu[z, y, x], ref[z, y, x]))


def test_loaded_native_library_is_the_in_tree_one(flt):
    assert os.path.abspath(L.library_path()).startswith(os.path.dirname(os.path.abspath(L.__file__)))
    maps = open("/proc/self/maps").read()
    assert "libmci_b200.so" in maps


@pytest.mark.parametrize("algo", sorted(cases.ALGOS))
def test_fixtures_and_handcrafted_all_algorithms(flt, seven_labels, many_to_many, algo):
    kw = cases.ALGOS[algo]
    inputs = {"SevenLabels": seven_labels, "ManyToMany": many_to_many, "SevenLabels_u8": seven_labels.astype(np.uint8)}
    inputs.update(cases.handcrafted())
    for name, vol in inputs.items():
        out = run_gpu(flt, vol, **kw)
        assert sha(out) == GOLD["%s_%s" % (name, algo)]["output"], (name, algo)
        assert_same(out, O.interpolate(vol, **kw), "%s_%s" % (name, algo))


def test_synthetic_small_configs(flt):
    for name, (vol, cfg) in cases.synthetic_small().items():
        out = run_gpu(flt, vol, axis=cfg.get("axis", -1), use_distance_transform=cfg["use_dt"], use_ball=cfg["ball"])
        assert sha(out) == GOLD["synthetic_%s" % name]["output"], name


@pytest.mark.parametrize("axis", [0, 1, 2])
def test_axis_restriction(flt, axis):
    vol = cases.handcrafted()["ThreeAxis"]
    assert_same(run_gpu(flt, vol, axis=axis), O.interpolate(vol, axis=axis), "ThreeAxis axis %d" % axis)
    vol, _ = cases.synthetic_small()["C4"]
    assert_same(run_gpu(flt, vol, axis=axis, use_distance_transform=False, use_ball=True),
                O.interpolate(vol, axis=axis, use_distance_transform=False, use_ball=True), "C4 axis %d" % axis)


def test_label_restriction(flt, seven_labels):
    for label in (3, 5, 200):
        assert_same(run_gpu(flt, seven_labels, label=label), O.interpolate(seven_labels, label=label), "label %d" % label)


def test_no_extrapolation_and_exhaustive_alignment(flt):
    h = cases.handcrafted()
    for name in ("ExtrapolationAppearing", "RingExtrapolate", "OneToThree", "DriftOddGaps"):
        v = h[name]
        assert_same(run_gpu(flt, v, extrapolation=False), O.interpolate(v, use_extrapolation=False), name + " noextrap")
        assert_same(run_gpu(flt, v, heuristic=False), O.interpolate(v, heuristic_alignment=False), name + " exhaustive")


def test_int16_and_uint8_pixel_types(flt):
    h = cases.handcrafted()
    for name in ("SimplestOneToOne", "OneToThree", "TwoAxisTwoLabel"):
        for dt in (np.int16, np.uint8):
            v = h[name].astype(dt)
            for kw in cases.ALGOS.values():
                assert_same(run_gpu(flt, v, **kw), O.interpolate(v, **kw), "%s %s" % (name, dt.__name__))


def test_negative_int16_labels_follow_the_write_rule(flt):
    # Image<short,3> may hold negative labels: they are detected and interpolated like any other (X:146, X:1459-1521),
    # but `out < label` (X:757) never holds on the zero background, so only positive labels reach the output; where a
    # negative and a positive label meet, the positive one wins.  Slices along z go through the compose pass.
    h = cases.handcrafted()
    v = h["DoubleTwoLabelBranching"].astype(np.int16)
    v[v == 2] = -3
    w = h["OneToThree"].astype(np.int16)
    w[w == 3] = -7
    for vol in (v, w):
        for algo in ("D", "C"):
            kw = cases.ALGOS[algo]
            out = run_gpu(flt, vol, **kw)
            assert_same(out, O.interpolate(vol, **kw), "negative labels " + algo)
    assert np.array_equal(run_gpu(flt, w), w)  # nothing but a negative label: the output is the input


def test_many_shapes_on_one_row_of_a_mid_slice(flt):
    # 150 components of one label side by side: every mid slice holds 150 shapes crossing the same rows, more than the
    # compose pass lists at once (COMPOSE_LIST = 64: several passes over the row), and with two labels interleaved the
    # larger label must win where their interpolated shapes touch
    n = 150
    v = np.zeros((7, 24, 8 * n + 8), np.uint16)
    for k in range(n):
        x = 4 + 8 * k
        lab = 3 if k % 2 else 9
        v[0, 8:13, x:x + 5] = lab
        v[6, 9:15, x + 1:x + 7] = lab
    for algo in ("D", "C"):
        kw = cases.ALGOS[algo]
        out = run_gpu(flt, v, axis=2, **kw)
        assert_same(out, O.interpolate(v, axis=2, **kw), "many shapes per row " + algo)
        assert (out[3] != 0).sum() > 20 * n


def test_int16_distance_bin_overflow_fails_cleanly(flt):
    # a distance of 3276.8 pixels or more makes `short dist = 10 * sdf` negative: the reference indexes out of
    # bounds (X:431-437); both the oracle and the CUDA path report an error instead
    v = np.zeros((9, 24, 3400), np.int16)
    v[0, 2:22, 2:22] = 1       # block + long thin bar: the alignment keeps the blocks on top of each other,
    v[0, 11:13, 2:3398] = 1    # so the far end of the bar is ~3380 pixels from the intersection
    v[8, 2:22, 2:22] = 1
    with pytest.raises(O.OracleError):
        O.interpolate(v)
    with pytest.raises(L.MciError):
        run_gpu(flt, v)
    # the context stays usable
    w = cases.handcrafted()["SimplestOneToOne"]
    assert_same(run_gpu(flt, w), O.interpolate(w), "after error")


@pytest.mark.parametrize("dt", [np.uint8, np.uint16, np.int16])
def test_far_distances_use_the_wrapped_bin_tables(flt, dt):
    # bins beyond a node's own table (16-bit types: distances of 656 / 330 pixels and more) and, for uchar, bins
    # wrapped modulo 256 many times over
    w = np.zeros((9, 24, 1500), dt)
    w[0, 2:22, 2:22] = 3
    w[0, 11:13, 2:1498] = 3
    w[8, 2:22, 2:22] = 3
    w[0, 4:20, 700:1400] = 3
    assert_same(run_gpu(flt, w), O.interpolate(w), "far distances %s" % np.dtype(dt).name)


def test_empty_and_no_slices_identity(flt):
    h = cases.handcrafted()
    for name in ("Empty", "NoSlices"):
        assert np.array_equal(run_gpu(flt, h[name]), h[name])
    one = np.zeros((1, 1, 1), np.uint16)
    assert np.array_equal(run_gpu(flt, one), one)
    odd = np.zeros((3, 5, 7), np.uint16)
    odd[:, 2, 3] = 4
    assert_same(run_gpu(flt, odd), O.interpolate(odd), "odd sizes")


def test_custom_slice_positions(flt, many_to_many):
    flt.SetLabel(0); flt.SetAxis(-1); flt.SetUseDistanceTransform(True); flt.SetUseBallStructuringElement(False)
    flt.SetHeuristicAlignment(True); flt.SetUseExtrapolation(True)
    flt.ClearLabeledSliceIndices()
    flt.SetLabeledSliceIndices(2, 1, [0, 7])
    flt.SetUseCustomSlicePositions(True)
    flt.SetInput(many_to_many)
    out = flt.Update()
    flt.SetUseCustomSlicePositions(False)
    assert_same(out, O.interpolate(many_to_many, custom_slices=[(2, 1, 0), (2, 1, 7)]), "custom positions")


def test_slice_detection_phase(flt):
    vol, _ = cases.synthetic_small()["C4"]
    flt.SetAxis(-1); flt.SetLabel(0); flt.SetUseCustomSlicePositions(False)
    flt.SetInput(vol)
    table = flt.DetermineSliceOrientations()
    got = set((a, l, s) for a, d in table.items() for l, ss in d.items() for s in ss)
    trip, _ = O.detect(vol)
    assert got == set(map(tuple, trip.tolist()))


def test_streaming_pass_through_the_tma_unit_with_a_ragged_tail(flt):
    # volumes of half a gigabyte and more are copied by cp.async.bulk in 16 KB tiles (k_copy_flag_tma), the remainder by
    # the register path: a size that is not a multiple of the tile, labels sitting in the last tile and in the tail
    X, Y, Z = 1031, 1021, 260                      # 547 MB of uint16, 273 691 660 voxels: 33 409 tiles + 8 184 bytes
    vol = np.zeros((Z, Y, X), np.uint16)
    vol[3, 100:140, 200:260] = 5
    vol[11, 104:150, 190:250] = 5                  # one segment far from the end
    vol[Z - 9, Y - 60:Y - 8, X - 70:X - 6] = 9
    vol[Z - 1, Y - 50:Y - 1, X - 64:X - 1] = 9     # annotated slice in the very last plane: last tile + tail
    vol[Z - 1, Y - 1, X - 5:X] = 9
    flt.SetAxis(2); flt.SetLabel(0); flt.SetUseCustomSlicePositions(False)
    flt.SetInput(vol)
    table = flt.DetermineSliceOrientations()
    got = set((a, l, s) for a, d in table.items() for l, ss in d.items() for s in ss)
    assert got == {(2, 5, 3), (2, 5, 11), (2, 9, Z - 9), (2, 9, Z - 1)}
    out = run_gpu(flt, vol, axis=2)
    assert np.array_equal(out[vol != 0], vol[vol != 0])            # the copy, tail included
    ref = O.interpolate(vol, axis=2, threads=os.cpu_count() or 1)
    assert_same(out, ref, "ragged 547 MB volume")
    assert (out != vol).any()


def test_connected_components_phase_raster_numbering(flt):
    # phase-level ABI: component images must equal the oracle's (ITK numbering), 4- and 8-connected
    rng = np.random.default_rng(5)
    vol = np.zeros((3, 70, 90), np.uint16)
    vol[1] = (rng.random((70, 90)) < 0.55) * 7
    lib = L.lib()
    dims = (ctypes.c_int64 * 3)(90, 70, 3)
    ctx = flt._ctx
    assert lib.mci_upload_volume(ctx, vol.ctypes.data, dims, L.U16) == 0
    for fully in (0, 1):
        desc = (L.SliceDesc * 1)(L.SliceDesc(2, 1, 7, 0, 0, 90, 70))
        n = (ctypes.c_int32 * 1)()
        assert lib.mci_label_slices(ctx, 1, desc, fully, n) == 0
        ids = np.zeros((70, 90), np.uint16)
        assert lib.mci_get_component_image(ctx, 0, ids.ctypes.data) == 0
        ref, nref = O.cc2d(vol[1] == 7, bool(fully))
        assert n[0] == nref
        assert np.array_equal(ids, ref)
        info = (L.ComponentInfo * n[0])()
        assert lib.mci_get_components(ctx, info) == 0
        for k in range(n[0]):
            yy, xx = np.nonzero(ref == k + 1)
            assert (info[k].count, info[k].sum_u, info[k].sum_v) == (len(xx), xx.sum(), yy.sum())
            assert (info[k].min_u, info[k].max_u, info[k].min_v, info[k].max_v) == (xx.min(), xx.max(), yy.min(), yy.max())


def test_full_size_c3_against_oracle_and_invariants(flt):
    # BASELINE.json configs[2]: 512x512x256 uint16, 20 labels, every 8th z-slice, DT, axis 2
    vol, dense, cfg = make_config("C3")
    out = run_gpu(flt, vol, axis=2)
    assert np.array_equal(out[vol != 0], vol[vol != 0])          # output contains input (X:1623-1636)
    assert set(np.unique(out)) <= set(np.unique(vol))            # no new labels
    ann = np.arange(0, vol.shape[0], 8)
    assert np.array_equal(out[ann], vol[ann])                    # annotated slices untouched
    assert (out != 0).sum() > 4 * (vol != 0).sum()               # the gaps were filled
    assert_same(out, O.interpolate(vol, axis=2, threads=os.cpu_count() or 1), "C3 full size")
    again = run_gpu(flt, out, axis=2)                            # a dense input: few or no isolated slices left
    assert_same(again, O.interpolate(out, axis=2, threads=os.cpu_count() or 1), "C3 second pass")


def test_ball_dilation_medium(flt):
    vol, _, cfg = make_config("C4", 0.25)
    out = run_gpu(flt, vol, use_distance_transform=False, use_ball=True)
    assert_same(out, O.interpolate(vol, use_distance_transform=False, use_ball=True, threads=os.cpu_count() or 1), "C4/4 ball")


def test_full_size_c4_ball_dilations_three_axes(flt):
    # BASELINE.json configs[3]: 1024x1024x512, sparse along all three axes, UseDistanceTransform=off, ball
    vol, dense, cfg = make_config("C4")
    del dense
    out = run_gpu(flt, vol, use_distance_transform=False, use_ball=True)
    assert np.array_equal(out[vol != 0], vol[vol != 0])
    assert_same(out, O.interpolate(vol, use_distance_transform=False, use_ball=True, threads=os.cpu_count() or 1), "C4 full size")


def _frozen(key):
    import json
    return json.load(open(os.path.join(cases.GOLDEN, "large.json")))[key]


def _plane_crcs(a):
    import zlib
    return [zlib.crc32(np.ascontiguousarray(a[z]).data) for z in range(a.shape[0])]


def _host_ram_gb():
    try:
        return os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_PHYS_PAGES") / 2.0 ** 30
    except (ValueError, OSError):
        return 0.0


def _check_against_frozen(flt, key, name, scale):
    """tests/golden/large.json freezes the ORACLE's output (tests/make_golden_large.py): sha256 of the volume and one
    CRC-32 per z-plane.  The generated input is checked too, so a drifted generator cannot hide behind a stale hash."""
    import hashlib
    gold = _frozen(key)
    vol, _, cfg = make_config(name, scale, want_dense=False)
    assert hashlib.sha256(vol.data).hexdigest() == gold["input"]
    out = run_gpu(flt, vol, axis=cfg["axis"], use_distance_transform=cfg["use_dt"], use_ball=cfg["ball"])
    crcs = _plane_crcs(out)
    bad = [z for z in range(out.shape[0]) if crcs[z] != gold["output_plane_crc32"][z]]
    assert not bad, "%s: %d planes differ from the frozen oracle output, first z = %d" % (key, len(bad), bad[0])
    assert hashlib.sha256(np.ascontiguousarray(out).data).hexdigest() == gold["output"]
    assert int(np.count_nonzero(out != vol)) == gold["changed_voxels"]


def test_half_scale_c5_against_the_frozen_oracle_output(flt):
    # 1024x1024x512, 200 labels: always runs (2 GB of host memory)
    _check_against_frozen(flt, "C5_half", "C5", 0.5)


@pytest.mark.skipif(_host_ram_gb() < 48 and os.environ.get("MCI_TEST_C5", "0") != "1",
                    reason="full-size C5 needs ~30 GB of host RAM (MCI_TEST_C5=1 forces it)")
def test_full_size_c5_on_one_gpu(flt):
    # BASELINE.json configs[4]: 2048x2048x1024 uint16, 200 labels (here on ONE GPU; the slab-sharded split is
    # tests/test_gpu_sharded.py and bench.py --gpus N)
    _check_against_frozen(flt, "C5", "C5", 1.0)


def _big_roi_volume(n=1200, dtype=np.uint16):
    """Shapes far larger than any shared-memory budget, with ragged borders and holes (many runs per row):
    exercises the global-scratch fall-backs of Align, the 1-to-N split, the node kernels and the run tables."""
    rng = np.random.default_rng(11)
    v = np.zeros((5, n, n), dtype)
    yy, xx = np.mgrid[:n, :n]
    big = (yy - n // 2) ** 2 + (xx - n // 2) ** 2 <= (n // 2 - 20) ** 2
    noise = rng.random((n, n)) < 0.35
    v[0][big & ~(noise & ((yy + xx) % 7 == 0))] = 3           # pepper holes: thousands of runs, one giant component
    two = ((yy - n // 2) ** 2 + (xx - n // 3) ** 2 <= (n // 7) ** 2) | ((yy - n // 2) ** 2 + (xx - 2 * n // 3) ** 2 <= (n // 8) ** 2)
    v[4][two] = 3                                             # 1-to-2 branching against the giant shape
    v[0][:14, :14][noise[:14, :14]] = 3                       # salt: a few dozen tiny components to extrapolate
    return v


def _one_to_n_volume(n_blobs, size, dtype=np.uint16):
    """one big disc on plane 0 against n_blobs small discs on plane 4, all inside the big disc's footprint"""
    v = np.zeros((5, size, size), dtype)
    yy, xx = np.mgrid[0:size, 0:size]
    c = size // 2
    v[0][(yy - c) ** 2 + (xx - c) ** 2 <= (size * 0.45) ** 2] = 2
    for k in range(n_blobs):
        ang = 2 * np.pi * k / n_blobs + 0.3
        cy, cx = c + size * 0.27 * np.sin(ang), c + size * 0.27 * np.cos(ang)
        r = size * (0.055 + 0.01 * (k % 3))
        v[4][(yy - cy) ** 2 + (xx - cx) ** 2 <= r ** 2] = 2
    return v


@pytest.mark.parametrize("algo", ["D", "C", "B"])
def test_one_to_n_split_for_every_blob_count_and_shape_size(flt, algo):
    """k_split: the register-resident path is instantiated for up to 2 / 3 / 4 blobs, one or up to four words per
    thread, cross and ball; more blobs or shapes beyond shared memory take the generic path"""
    kw = cases.ALGOS[algo]
    for n_blobs, size in [(2, 64), (3, 64), (4, 64), (5, 64), (6, 96), (2, 300), (3, 340), (4, 300), (2, 40), (1, 64)]:
        v = _one_to_n_volume(n_blobs, size)
        out = run_gpu(flt, v, **kw)
        assert_same(out, O.interpolate(v, threads=os.cpu_count() or 1, **kw), "1-to-%d at %d px, %s" % (n_blobs, size, algo))


@pytest.mark.parametrize("algo", ["D", "C"])
def test_big_regions_take_the_global_scratch_paths(flt, algo):
    kw = cases.ALGOS[algo]
    v = _big_roi_volume()
    out = run_gpu(flt, v, **kw)
    assert_same(out, O.interpolate(v, threads=os.cpu_count() or 1, **kw), "big ROI " + algo)


def test_volume_series_throughput_mode_matches_single_runs():
    # mci_filter_run_batch: several volumes in flight on separate contexts / streams / host threads
    from itkcontourinterpolation_b200.filter import VolumeSeriesInterpolator
    h = cases.handcrafted()
    vols = [v for v in h.values() if v.shape == (9, 40, 48)]
    assert len(vols) >= 6
    vols = vols * 3                                              # 3 rounds through every in-flight slot
    for depth, kw_gpu, kw_cpu in ((3, dict(), dict()),
                                  (2, dict(use_distance_transform=0, use_ball_structuring_element=1),
                                   dict(use_distance_transform=False, use_ball=True))):
        s = VolumeSeriesInterpolator((9, 40, 48), np.uint16, depth=depth, **kw_gpu)
        try:
            outs = s.run(vols)
        finally:
            s.close()
        assert len(outs) == len(vols)
        for k, (v, o) in enumerate(zip(vols, outs)):
            assert_same(o, O.interpolate(v, **kw_cpu), "series volume %d depth %d" % (k, depth))
    vol, _, cfg = make_config("C3", 0.5)
    ref = O.interpolate(vol, axis=2, threads=os.cpu_count() or 1)
    s = VolumeSeriesInterpolator(vol.shape, np.uint16, depth=3, axis=2)
    try:
        for o in s.run([vol] * 7):
            assert_same(o, ref, "C3/2 series")
    finally:
        s.close()
