"""The batched persistent kernel (xm_frame_batch -> batch_kernel): every frame of a batch must be
bit-identical to the oracle, whatever the mix of frame sizes, and unsorted frames inside a batch must be
re-rendered exactly on the device."""
import numpy as np
import pytest

from oracle import xmaps_oracle as orc
from test_gpu_parity import E, make_engine
from xm_helpers import golden_frame, load_golden_tables

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


# batch_kernel (the persistent batch kernel) with both projector epilogues: the two-pass strip epilogue (default) and
# the shared-memory tile epilogue
BATCH_KERNELS = [(1, 1), (1, 0)]
BATCH_IDS = ["strips", "tiles"]


@pytest.fixture(scope="module", params=BATCH_KERNELS, ids=BATCH_IDS)
def small(request):
    tables, z = load_golden_tables("small")
    eng = make_engine(tables, z)
    eng.set_option("batch", request.param[0])
    eng.set_option("batch_strips", request.param[1])
    assert eng.get_option("batch") == request.param[0] and eng.get_option("batch_occ") >= 1
    assert eng.get_option("batch_strips") == request.param[1]
    yield tables, z, eng
    eng.close()


@pytest.fixture(scope="module", params=BATCH_KERNELS, ids=BATCH_IDS)
def default(request):
    tables, z = load_golden_tables("default")
    eng = make_engine(tables)
    eng.set_option("batch", request.param[0])
    eng.set_option("batch_strips", request.param[1])
    yield tables, z, eng
    eng.close()


def want_frames(tables, frames, view):
    return [orc.frame_depth(tables, f, view) for f in frames]


SIZES = [20_000, 0, 1, 1023, 1024, 1025, 70_001, 5, 0, 33_000, 4097, 64]


@pytest.mark.parametrize("view", [0, 1])
def test_mixed_sizes_match_oracle(small, view):
    tables, _, eng = small
    frames = [orc.synth_events(300 + i, n, 160, 120) for i, n in enumerate(SIZES)]
    launches0 = eng.launch_count()
    out = eng.frame_batch(frames, view=view).cpu().numpy()
    assert eng.launch_count() - launches0 == 3  # bounds + batch + redo scan: the batch kernel really ran
    for i, w in enumerate(want_frames(tables, frames, view)):
        assert np.array_equal(out[i], w), f"frame {i} ({SIZES[i]} events)"
    st = eng.status()  # the last frame of the batch
    ev = frames[-1]
    assert st["n_valid"] == int((ev["p"] == 1).sum())
    assert not st["fixup_ran"] and not st["tbounds_violated"]


def test_outputs_disparity_and_bgr(small):
    tables, _, eng = small
    e = E()
    frames = [orc.synth_events(2, 20_000, 160, 120), orc.synth_events(9, 3_000, 160, 120)]
    g = golden_frame("small_20k_proj")
    bgr = eng.frame_batch(frames, view=0, output=e.OUT_BGR).cpu().numpy()
    assert np.array_equal(bgr[0], g["bgr"])
    disp = eng.frame_batch(frames, view=0, output=e.OUT_DISPARITY).cpu().numpy()
    for i, f in enumerate(frames):
        assert np.array_equal(disp[i], orc.frame_disparity_map(tables, f, 0))
    single = eng.frame(frames[1], view=0, output=e.OUT_BGR).cpu().numpy()
    assert np.array_equal(bgr[1], single)


@pytest.mark.parametrize("n_frames", [33, 40, 65])
def test_more_frames_than_one_launch(small, n_frames):
    tables, _, eng = small
    base = [orc.synth_events(500 + i, 2_000 + 37 * i, 160, 120) for i in range(8)]
    want = want_frames(tables, base, 0)
    frames = [base[i % 8] for i in range(n_frames)]
    out = eng.frame_batch(frames, view=0).cpu().numpy()
    for i in range(n_frames):
        assert np.array_equal(out[i], want[i % 8]), f"frame {i}"


def test_unsorted_frames_inside_a_batch(small):
    tables, _, eng = small
    e = E()
    frames = [orc.synth_events(600 + i, 15_000, 160, 120) for i in range(6)]
    for i in (1, 4, 5):
        np.random.default_rng(i).shuffle(frames[i])
    for view in (0, 1):
        out = eng.frame_batch(frames, view=view, time_bounds=e.TBOUNDS_SORTED).cpu().numpy()
        st = eng.status()
        assert st["fixup_ran"] and st["tbounds_violated"]  # frame 5 is the last one
        for i, w in enumerate(want_frames(tables, frames, view)):
            assert np.array_equal(out[i], w), f"view {view} frame {i}"
        assert st["n_valid"] == int((frames[5]["p"] == 1).sum())
    # a clean batch right afterwards is unaffected, and reports no fix-up
    clean = [orc.synth_events(700 + i, 9_000, 160, 120) for i in range(3)]
    out = eng.frame_batch(clean, view=0).cpu().numpy()
    assert not eng.status()["fixup_ran"]
    for i, w in enumerate(want_frames(tables, clean, 0)):
        assert np.array_equal(out[i], w)


def test_given_bounds(small):
    tables, _, eng = small
    e = E()
    frames = [orc.synth_events(800 + i, 12_000, 160, 120) for i in range(4)]
    good = []
    for f in frames:
        t = f["t"][f["p"] == 1]
        good.append((int(t.min()), int(t.max())))
    want = want_frames(tables, frames, 1)
    out = eng.frame_batch(frames, view=1, time_bounds=e.TBOUNDS_GIVEN, t_bounds=good).cpu().numpy()
    assert not eng.status()["fixup_ran"]
    for i in range(4):
        assert np.array_equal(out[i], want[i])
    bad = list(good)
    bad[2] = (100, 5000)
    bad[3] = (good[3][0] + 50, good[3][1])
    out = eng.frame_batch(frames, view=1, time_bounds=e.TBOUNDS_GIVEN, t_bounds=bad).cpu().numpy()
    assert eng.status()["fixup_ran"]
    for i in range(4):
        assert np.array_equal(out[i], want[i])


def test_batch_single_interleaving_and_epoch_wrap(small):
    tables, _, eng = small
    a = [orc.synth_events(900 + i, 6_000, 160, 120) for i in range(5)]
    wa = want_frames(tables, a, 0)
    eng.set_option("epoch", 0xFFFF - 12)
    for rep in range(4):  # each batch takes 10 epochs: crosses the 16-bit wrap
        out = eng.frame_batch(a, view=0).cpu().numpy()
        for i in range(5):
            assert np.array_equal(out[i], wa[i]), f"rep {rep} frame {i}"
        assert np.array_equal(eng.frame(a[rep], view=0).cpu().numpy(), wa[rep])
    assert eng.get_option("epoch") < 100


def test_batch_option_off_is_identical(small):
    tables, _, eng = small
    frames = [orc.synth_events(40 + i, 8_000 + 100 * i, 160, 120) for i in range(5)]
    on = eng.frame_batch(frames, view=0).cpu().numpy()
    mode = eng.get_option("batch")
    eng.set_option("batch", 0)
    try:
        off = eng.frame_batch(frames, view=0).cpu().numpy()
    finally:
        eng.set_option("batch", mode)
    assert np.array_equal(on, off)


def test_pixel_oob_is_flagged_per_frame(small):
    tables, _, eng = small
    frames = [orc.synth_events(2, 2_000, 160, 120) for _ in range(3)]
    frames[2] = frames[2].copy()
    frames[2]["x"][7] = 160
    eng.frame_batch(frames, view=1)
    assert eng.status()["pixel_oob"]


def test_heavy_collisions_in_a_batch(small):
    tables, _, eng = small
    rng = np.random.default_rng(3)
    frames = []
    for i in range(3):
        n = 150_000
        ev = orc.synth_events(30 + i, n, 160, 120, p_on=1.0)
        ev["x"] = rng.integers(60, 64, n)
        ev["y"] = rng.integers(50, 54, n)
        frames.append(ev)
    for view in (0, 1):
        out = eng.frame_batch(frames, view=view).cpu().numpy()
        for i, w in enumerate(want_frames(tables, frames, view)):
            assert np.array_equal(out[i], w)


def test_default_geometry_1m_events(default):
    """BASELINE geometry, 6 x 1 M events: the tile items of one frame overlap the chunks of the next."""
    tables, _, eng = default
    frames = [orc.synth_events(1000 + i, 1_000_000, 640, 480) for i in range(6)]
    for view in (0, 1):
        out = eng.frame_batch(frames, view=view).cpu().numpy()
        for i in (0, 3, 5):
            assert np.array_equal(out[i], orc.frame_depth(tables, frames[i], view)), f"view {view} frame {i}"
        # the other frames against the single-frame kernels
        for i in (1, 2, 4):
            assert np.array_equal(out[i], eng.frame(frames[i], view=view).cpu().numpy())


# ------------------------------------------------------------------------------------------------
# the "alive" table (events whose pixel block cannot yield an inlier at their time column are only counted / bounds-checked)
# ------------------------------------------------------------------------------------------------
ALIVE_BLOCK = 8  # xm_capi.cu:build_alive picks 8 x 8 pixels for the golden geometries (<= 6144 blocks)


def alive_blocks(tables):
    """Host restatement of the first half of xm_capi.cu:build_alive: per camera pixel, can ANY time column make it an
    inlier?  (The table also keeps, per block, the hull of those columns; the parity tests cover that part.)"""
    xm = tables.x_map.astype(np.int32)
    ly, lx = tables.lut_y.astype(np.int32), tables.lut_x.astype(np.int32)
    h, w = ly.shape
    alive = np.zeros((h, w), bool)
    y_ok = (ly >= 0) & (ly < xm.shape[0] - 1)
    for y in range(h):
        for x in range(w):
            if not y_ok[y, x]:
                continue
            d = (xm[ly[y, x]] - lx[y, x] - tables.x_offset).astype(np.int16)
            alive[y, x] = bool((d >= 0).any())
    b = ALIVE_BLOCK
    bh, bw = (h + b - 1) // b, (w + b - 1) // b
    pad = np.zeros((bh * b, bw * b), bool)
    pad[:h, :w] = alive
    blocks = pad.reshape(bh, b, bw, b).any(axis=(1, 3))
    return alive, blocks


def test_alive_bitmap_is_exact_and_invisible(small):
    tables, _, eng = small
    alive, blocks = alive_blocks(tables)
    px_in_alive_blocks = int(np.kron(blocks, np.ones((ALIVE_BLOCK, ALIVE_BLOCK), bool))[: tables.cam_h, : tables.cam_w].sum())
    assert eng.get_option("alive_px") == px_in_alive_blocks
    assert 0 < px_in_alive_blocks < tables.cam_h * tables.cam_w  # the small geometry has dead blocks
    frames = [orc.synth_events(700 + i, 30_000 + 1000 * i, 160, 120) for i in range(5)]
    outs = {}
    for flag in (1, 0):
        eng.set_option("alive", flag)
        outs[flag] = [eng.frame_batch(frames, view=v).cpu().numpy() for v in (0, 1)]
        st = eng.status()
        outs[flag].append((st["n_valid"], st["n_inliers"], st["flags"]))
    eng.set_option("alive", 1)
    for v in (0, 1):
        assert np.array_equal(outs[1][v], outs[0][v])
        for i, f in enumerate(frames):
            assert np.array_equal(outs[1][v][i], orc.frame_depth(tables, f, v))
    assert outs[1][2] == outs[0][2]


def test_dead_pixel_event_outside_bounds_still_triggers_the_redo(small):
    """t.min() / t.max() run over ALL kept events (x_maps_disparity.py:12-13), dead pixels included: an
    out-of-order timestamp on a dead pixel must invalidate the frame's assumed bounds like any other."""
    tables, _, eng = small
    _, blocks = alive_blocks(tables)
    by, bx = np.argwhere(~blocks)[0]
    ev = orc.synth_events(41, 40_000, 160, 120)
    ev["p"] = 1
    k = 20_000
    ev["x"][k], ev["y"][k] = bx * ALIVE_BLOCK, by * ALIVE_BLOCK
    ev["t"][k] = ev["t"].min() - 500  # earlier than the first event: the sorted-bounds assumption is wrong
    other = orc.synth_events(42, 10_000, 160, 120)
    out = eng.frame_batch([other, ev, other], view=0).cpu().numpy()
    assert np.array_equal(out[1], orc.frame_depth(tables, ev, 0))
    assert np.array_equal(out[0], orc.frame_depth(tables, other, 0)) and np.array_equal(out[0], out[2])
    # and alone, so that the status block is this frame's
    got = eng.frame_batch([ev, ev], view=0).cpu().numpy()
    assert np.array_equal(got[1], orc.frame_depth(tables, ev, 0))
    assert eng.status()["fixup_ran"]


@pytest.mark.parametrize("strips", [1, 0], ids=BATCH_IDS)
def test_ragged_projector_image(strips):
    """A projector image whose pixel count is not a multiple of the strip epilogue's 256-pixel blocks (nor of the tile
    size), with remap targets that leave the rectified image and a window that touches its border: every pixel of every
    frame must still equal the oracle's."""
    import dataclasses

    tables, z = load_golden_tables("small")
    remap = np.ascontiguousarray(tables.remap_xy[3:310, 5:178]).copy()  # 307 rows x 173 columns
    remap[:9, :, 0] = -5  # pixels whose targets lie outside the rectified image (left / below)
    remap[-7:, :, 1] = tables.rect_h
    remap[20:24, :, 0] = 0  # ... and on its first / last column and last row: the window's halo leaves the image
    remap[40:44, :, 0] = tables.rect_w - 1
    remap[30:34, :, 1] = tables.rect_h - 1
    remap[50:52, :, 1] = 0
    t2 = dataclasses.replace(tables, remap_xy=remap)
    eng = make_engine(t2, z)
    try:
        eng.set_option("batch_strips", strips)
        frames = [orc.synth_events(700 + i, n, 160, 120) for i, n in enumerate([30_000, 1, 0, 12_345, 64, 50_000, 2_047])]
        out = eng.frame_batch(frames, view=0).cpu().numpy()
        assert out.shape[1:] == (307, 173)
        for i, f in enumerate(frames):
            assert np.array_equal(out[i], orc.frame_depth(t2, f, 0)), f"frame {i}"
        # item sizes the launch would not pick by itself
        eng.set_option("strip_rows", 5)
        eng.set_option("strip_blocks", 3)
        eng.set_option("tile_warps", 2)
        eng.set_option("batch_maps", 2)
        out = eng.frame_batch(frames, view=0).cpu().numpy()
        for i, f in enumerate(frames):
            assert np.array_equal(out[i], orc.frame_depth(t2, f, 0)), f"frame {i} (odd item sizes)"
    finally:
        eng.close()


def test_exact_rounding_ties_in_every_chunk(small):
    """t range = 2 * T_PX_SCALE: every odd timestamp sits exactly half-way between two time columns (round half to even
    in the reference's float64 expression).  The integer fast pass only detects those; every chunk is redone by the
    general front half, whose live list (which depends on the time column) is laid out anew."""
    tables, _, eng = small
    frames = []
    for i in range(4):
        ev = orc.synth_events(900 + i, 25_000 + 3_000 * i, 160, 120)
        rng = np.random.default_rng(i)
        t = np.sort(rng.integers(0, 2 * tables.t_px_scale + 1, len(ev))).astype(np.int64)
        t[0], t[-1] = 0, 2 * tables.t_px_scale
        ev["t"] = t + 1000 * i
        ev["p"][0] = ev["p"][-1] = 1  # the bounds are the first / last KEPT event
        frames.append(ev)
    for view in (0, 1):
        out = eng.frame_batch(frames, view=view).cpu().numpy()
        for i, f in enumerate(frames):
            assert np.array_equal(out[i], orc.frame_depth(tables, f, view)), f"view {view} frame {i}"
    assert not eng.status()["fixup_ran"]


def test_alive_table_is_conservative_for_a_non_monotonic_x_map():
    """The alive table keeps, per pixel block, the HULL of the time columns that can yield inliers.  For the reference's
    X-maps a pixel's inlier columns are one interval; for an arbitrary table they are not -- shuffle the columns of
    every X-map row differently and the hull must still never drop an inlier."""
    import dataclasses

    tables, z = load_golden_tables("small")
    rng = np.random.default_rng(7)
    xm = tables.x_map.copy()
    for y in range(xm.shape[0]):
        xm[y] = xm[y, rng.permutation(xm.shape[1])]
    t2 = dataclasses.replace(tables, x_map=xm)
    eng = make_engine(t2, z)
    try:
        frames = [orc.synth_events(1300 + i, 40_000, 160, 120) for i in range(3)]
        stats = {}
        for flag in (1, 0):
            eng.set_option("alive", flag)
            for view in (0, 1):
                out = eng.frame_batch(frames, view=view).cpu().numpy()
                for i, f in enumerate(frames):
                    assert np.array_equal(out[i], orc.frame_depth(t2, f, view)), f"alive {flag} view {view} frame {i}"
            st = eng.status()
            stats[flag] = (st["n_valid"], st["n_inliers"])
        assert stats[0] == stats[1] and stats[1][1] > 0
    finally:
        eng.close()
