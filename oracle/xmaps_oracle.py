"""CPU oracle for the X-maps per-event depth path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``x-maps_b200/`` (the product) may import this
module; it is used by ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` as the checker and as the timed CPU baseline.

It restates, in NumPy (and OpenCV for the two dense image ops the reference itself
delegates to OpenCV), the algorithm of fraunhoferhhi/X-maps for SURVEY.md §8 rows A0-A5
plus the "next" rows N1 (colourise) and N3 (X-map build).  Every function cites the
reference lines it follows (paths relative to ``/root/reference``).

Parity status: PINNED.  The reference is pure Python and imports in the authoring
container, so ``tests/golden/generate_golden.py`` runs the *real* reference on seeded
inputs and commits its outputs under ``tests/golden/``; ``tests/test_oracle_golden.py``
checks this restatement against those vectors bit-for-bit.

One exception, marked where it stands: ``ActivityNoiseFilterOracle`` restates the closed Metavision
``ActivityNoiseFilterAlgorithm`` from its published semantics only -- PARITY UNPINNED (no binary, no
golden vector to check it against).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

# Metavision EventCD record: 16-byte AoS, t in microseconds.
EVENT_DTYPE = np.dtype(
    {
        "names": ["x", "y", "p", "t"],
        "formats": ["<u2", "<u2", "<i2", "<i8"],
        "offsets": [0, 2, 4, 8],
        "itemsize": 16,
    }
)

VIEW_PROJECTOR = 0
VIEW_CAMERA = 1


@dataclass
class OracleTables:
    """Read-only tables of one calibration (what the reference keeps on CamProjMaps /
    XMapsDisparity, python/cam_proj_calibration.py:143-172, python/x_maps_disparity.py:35-67)."""

    lut_x: np.ndarray  # [cam_h, cam_w] int16   disp_cam_mapx_i16
    lut_y: np.ndarray  # [cam_h, cam_w] int16   disp_cam_mapy_i16
    x_map: np.ndarray  # [rect_h, xmap_w] int16 proj_x_map (value = x_rect + X_OFFSET, 0 = undefined)
    remap_xy: np.ndarray  # [proj_h, proj_w, 2] int16 disp_proj_mapxy_i16 (ch0 = x_rect, ch1 = y_rect)
    rect_w: int
    rect_h: int
    t_px_scale: int  # X_MAP_WIDTH - 1
    x_offset: int  # 4242
    depth_scale: float  # P2[0, 3] (float64)
    dilate: int = 7

    @property
    def cam_h(self):
        return self.lut_x.shape[0]

    @property
    def cam_w(self):
        return self.lut_x.shape[1]

    @property
    def proj_h(self):
        return self.remap_xy.shape[0]

    @property
    def proj_w(self):
        return self.remap_xy.shape[1]


# --------------------------------------------------------------------------------------
# A0  polarity mask
# --------------------------------------------------------------------------------------
def polarity_mask(events: np.ndarray) -> np.ndarray:
    """Keep positive events, order preserved.

    Reference: ``PolarityFilterAlgorithm(1).process_events`` (closed Metavision binary; call
    site python/depth_reprojection_pipe.py:43,114).  Restated as ``p == 1`` exactly as the
    reference restates it itself in python/frame_event_filter.py:21,47,72,104.
    """
    return events[events["p"] == 1]


# --------------------------------------------------------------------------------------
# A1  rectify event pixels through the int16 LUT
# --------------------------------------------------------------------------------------
def rectify_i16(tables: OracleTables, events) -> Tuple[np.ndarray, np.ndarray]:
    """python/cam_proj_calibration.py:277-281 (rectify_cam_coords_i16)."""
    ys = np.asarray(events["y"])
    xs = np.asarray(events["x"])
    return tables.lut_x[ys, xs], tables.lut_y[ys, xs]


# --------------------------------------------------------------------------------------
# A2  X-map lookup -> disparity
# --------------------------------------------------------------------------------------
def time_to_xmap_column(t: np.ndarray, t_px_scale: int, t_min=None, t_max=None) -> np.ndarray:
    """python/x_maps_disparity.py:12-19.

    float64 arithmetic, two roundings (divide, multiply) then round-half-even, then a cast
    to int16.  ``t`` may be int64 microseconds (live path) or float (eval path,
    python/eval/compute_depth_x_maps.py:91).  A frame whose timestamps are all equal gives
    0/0 = NaN for every event; NumPy's NaN -> int16 cast yields 0 on x86-64, which is the
    behaviour restated (and tested) here.
    """
    t = np.asarray(t)
    lo = t.min() if t_min is None else t_min
    hi = t.max() if t_max is None else t_max
    if hi == lo:
        # every quotient is 0/0 = NaN (NumPy warns and casts NaN -> int16 as 0 on x86-64)
        return np.zeros(t.shape, dtype=np.int16)
    norm = (t - lo) / (hi - lo)
    return np.rint(norm * t_px_scale).astype(np.int16)


def event_disparity(tables: OracleTables, xcr: np.ndarray, ycr: np.ndarray, t: np.ndarray):
    """python/x_maps_disparity.py:9-32 (compute_disparity).

    Returns ``(disp[M] int16, inlier_mask[N] bool)``; ``disp`` is compacted to the inliers.
    Note the last X-map row is excluded (``ycr < H - 1``), and that an undefined X-map cell
    (value 0) drops out because ``0 - xcr - X_OFFSET < 0`` whenever ``xcr > -X_OFFSET``.
    """
    if len(t) == 0:
        return np.zeros(0, np.int16), np.zeros(0, bool)
    col = time_to_xmap_column(t, tables.t_px_scale)
    y_ok = (ycr >= 0) & (ycr < tables.x_map.shape[0] - 1)
    x_proj = tables.x_map[ycr[y_ok], col[y_ok]]
    # int16 arithmetic with wrap-around, as NumPy does for int16 - int16 - small python int
    disp = (x_proj.astype(np.int32) - xcr[y_ok].astype(np.int32) - tables.x_offset).astype(np.int16)
    keep = disp >= 0
    mask = y_ok.copy()
    mask[y_ok] = keep
    return disp[keep], mask


# --------------------------------------------------------------------------------------
# A3  scatter into a disparity map (last write wins)
# --------------------------------------------------------------------------------------
def _last_write_wins_scatter(shape, rows, cols, values) -> np.ndarray:
    """``m[rows, cols] = values`` exactly as the reference writes it (python/cam_proj_calibration.py:301-302,
    315-316): one fancy assignment into a zeroed float32 map.  For duplicate targets NumPy keeps the LAST
    occurrence (documented for integer-array assignment; ``_last_write_wins_scatter_explicit`` spells the rule
    out without relying on it and tests/test_oracle_properties.py holds the two together)."""
    out = np.zeros(shape, dtype=np.float32)
    out[rows, cols] = values
    return out


def _last_write_wins_scatter_explicit(shape, rows, cols, values) -> np.ndarray:
    """The same scatter with the duplicate rule made explicit: the highest event index wins."""
    out = np.zeros(shape, dtype=np.float32)
    if len(values) == 0:
        return out
    flat = rows.astype(np.int64) * shape[1] + cols.astype(np.int64)
    # first occurrence in the reversed sequence == last occurrence in the original
    _, first_rev = np.unique(flat[::-1], return_index=True)
    winners = len(flat) - 1 - first_rev
    out.reshape(-1)[flat[winners]] = values[winners].astype(np.float32)
    return out


def scatter_projector_view(tables: OracleTables, xcr, ycr, mask, disp) -> np.ndarray:
    """python/cam_proj_calibration.py:299-303 (compute_disp_map_projector_view).

    ``xpr = int16(rint(xcr + disp))`` is int16 + int16 (wraps), i.e. ``x_proj - X_OFFSET``.
    Negative indices follow NumPy's wrap-around; out-of-range ones raise IndexError (NumPy's own
    check, as in the reference).
    """
    xpr = np.rint(xcr[mask] + disp).astype(np.int16)  # the reference's own expression (:300)
    return _last_write_wins_scatter((tables.rect_h, tables.rect_w), ycr[mask], xpr, disp)


def scatter_camera_view(tables: OracleTables, events, mask, disp) -> np.ndarray:
    """python/cam_proj_calibration.py:312-317 (compute_disp_map_camera_view)."""
    xs = np.asarray(events["x"])[mask]
    ys = np.asarray(events["y"])[mask]
    return _last_write_wins_scatter((tables.cam_h, tables.cam_w), ys, xs, disp)


# --------------------------------------------------------------------------------------
# A4  7x7 dilate + nearest remap rect -> projector
# --------------------------------------------------------------------------------------
def dilate_remap(tables: OracleTables, rect_disp_map: np.ndarray, use_cv2: bool = True) -> np.ndarray:
    """python/disp_to_depth.py:76-97 (remap_rectified_disp_map_to_proj).

    ``cv2.dilate`` with a k x k all-ones kernel = windowed max, anchor at the centre, taps that
    fall outside the image ignored; ``cv2.remap`` INTER_NEAREST with a CV_16SC2 map and
    BORDER_CONSTANT(0) = plain gather, 0 where the source coordinate is outside the image.
    ``use_cv2=False`` is a dependency-free restatement used to cross-check the OpenCV one.
    """
    k = tables.dilate
    if use_cv2:
        import cv2

        dil = cv2.dilate(rect_disp_map, np.ones((k, k), dtype=np.uint8))
        return cv2.remap(
            dil,
            map1=tables.remap_xy,
            map2=None,
            interpolation=cv2.INTER_NEAREST,
            borderMode=cv2.BORDER_CONSTANT,
        )
    r = k // 2
    h, w = rect_disp_map.shape
    pad = np.full((h + 2 * r, w + 2 * r), -np.inf, dtype=np.float32)
    pad[r : r + h, r : r + w] = rect_disp_map
    rows = pad[:, 0:w].copy()
    for dx in range(1, k):
        np.maximum(rows, pad[:, dx : dx + w], out=rows)
    dil = rows[0:h].copy()
    for dy in range(1, k):
        np.maximum(dil, rows[dy : dy + h], out=dil)
    mx = tables.remap_xy[..., 0].astype(np.int64)
    my = tables.remap_xy[..., 1].astype(np.int64)
    inside = (mx >= 0) & (mx < w) & (my >= 0) & (my < h)
    out = np.zeros(mx.shape, dtype=np.float32)
    out[inside] = dil[my[inside], mx[inside]]
    return out


# --------------------------------------------------------------------------------------
# A5  disparity -> metric depth
# --------------------------------------------------------------------------------------
def disparity_to_depth(disp_map: np.ndarray, depth_scale: float) -> np.ndarray:
    """python/disp_to_depth.py:46-63 (disparity_to_depth_rectified).

    ``depth = 0 if d == 0 else max(P[0,3] / d, 1e-9)``: float64 divide, float64 max, stored as
    float32 (one rounding, round-to-nearest-even).
    """
    d = disp_map.astype(np.float64)
    with np.errstate(divide="ignore"):
        z = np.maximum(np.float64(depth_scale) / d, 1e-9)
    z[disp_map == 0] = 0.0
    return z.astype(np.float32)


# --------------------------------------------------------------------------------------
# whole path (driver = python/depth_reprojection_pipe.py:121-167 up to the depth frame)
# --------------------------------------------------------------------------------------
def frame_disparity_map(tables: OracleTables, events, view: int, apply_polarity: bool = True, use_cv2: bool = True):
    evs = polarity_mask(events) if apply_polarity else events
    xcr, ycr = rectify_i16(tables, evs)
    disp, mask = event_disparity(tables, xcr, ycr, np.asarray(evs["t"]))
    if view == VIEW_CAMERA:
        return scatter_camera_view(tables, evs, mask, disp)
    rect = scatter_projector_view(tables, xcr, ycr, mask, disp)
    return dilate_remap(tables, rect, use_cv2=use_cv2)


def frame_depth(tables: OracleTables, events, view: int, apply_polarity: bool = True, use_cv2: bool = True):
    """Polarity mask -> rectify -> X-map disparity -> scatter -> (dilate + remap) -> depth."""
    return disparity_to_depth(
        frame_disparity_map(tables, events, view, apply_polarity, use_cv2), tables.depth_scale
    )


# --------------------------------------------------------------------------------------
# N1  colourise (tail of process_ev_frame)
# --------------------------------------------------------------------------------------
def frame_disparity_map_bilinear(tables: OracleTables, lut_x_f32: np.ndarray, lut_y_f32: np.ndarray, events, view: int,
                                 apply_polarity: bool = True, use_cv2: bool = True):
    """Opt-in extension (BASELINE config 3, "bilinear X-map lookup"), NOT reference behaviour: the reference rounds the
    rectified row and the time column and reads ONE X-map cell (python/x_maps_disparity.py:19,25).  Here the X-map is
    sampled bilinearly at the un-rounded position:

      (x_r, y_r) = float32 LUT (rectify_cam_coords_f32, cam_proj_calibration.py:272-275);  c = (t - min) / (max - min) * T
      taps X[y0 + i, c0 + j] (y0 = floor(y_r), c0 = floor(c), c0 + 1 clipped to the last column), rows restricted to
      0 <= y0, y0 + 1 <= H - 1 like the reference's `0 <= y < H - 1`; cells equal to 0 are undefined (x_map.py) and
      are left out of the blend, the remaining weights renormalised; disparity = float32(x_p - x_r - X_OFFSET) >= 0;
      last event per cell wins; cell = (rint(y_r), rint(x_p - X_OFFSET)) (projector view) or the camera pixel.

    float64 throughout, in the same operation order as the CUDA kernel (bilinear_scatter_kernel).  Returns the float32
    disparity map of the view (projector view: after the 7x7 dilate + nearest remap)."""
    ev = polarity_mask(events) if apply_polarity else events
    h_rows, w_cols = tables.x_map.shape
    if len(ev) == 0:
        rect = np.zeros((tables.rect_h, tables.rect_w) if view == 0 else (tables.cam_h, tables.cam_w), dtype=np.float32)
        return dilate_remap(tables, rect, use_cv2) if view == 0 else rect
    x, y = ev["x"].astype(np.int64), ev["y"].astype(np.int64)
    t = ev["t"]
    xr = lut_x_f32[y, x].astype(np.float64)
    yr = lut_y_f32[y, x].astype(np.float64)
    with np.errstate(invalid="ignore", divide="ignore"):
        c = (t - t.min()) / (t.max() - t.min()) * tables.t_px_scale  # x_maps_disparity.py:12-19 without the rint
    c = np.where(np.isnan(c), 0.0, c)
    c0, y0 = np.floor(c), np.floor(yr)
    ok = (y0 >= 0) & (y0 + 1 <= h_rows - 1) & (c0 >= 0) & (c0 <= tables.t_px_scale)
    fc, fy = c - c0, yr - y0
    ic0 = np.where(ok, c0, 0).astype(np.int64)
    iy0 = np.where(ok, y0, 0).astype(np.int64)
    ic1 = np.minimum(ic0 + 1, w_cols - 1)
    xm = tables.x_map.astype(np.float64)
    v = [xm[iy0, ic0], xm[iy0, ic1], xm[iy0 + 1, ic0], xm[iy0 + 1, ic1]]
    gy, gc = 1.0 - fy, 1.0 - fc
    w = [gy * gc, gy * fc, fy * gc, fy * fc]
    num = np.zeros(len(ev))
    den = np.zeros(len(ev))
    for vi, wi in zip(v, w):
        m = vi != 0
        num = num + np.where(m, wi * vi, 0.0)
        den = den + np.where(m, wi, 0.0)
    good = ok & (den > 0)
    with np.errstate(invalid="ignore", divide="ignore"):
        xp = num / den
    disp = ((xp - xr) - tables.x_offset).astype(np.float32)
    inl = good & (disp >= 0)
    if view == 1:
        out = np.zeros((tables.cam_h, tables.cam_w), dtype=np.float32)
        out[y[inl], x[inl]] = disp[inl]  # NumPy keeps the last duplicate
        return out
    rows = np.rint(yr)
    cols = np.rint(xp - tables.x_offset)
    inside = inl & (rows >= 0) & (rows < tables.rect_h) & (cols >= 0) & (cols < tables.rect_w)
    rect = np.zeros((tables.rect_h, tables.rect_w), dtype=np.float32)
    rect[rows[inside].astype(np.int64), cols[inside].astype(np.int64)] = disp[inside]
    return dilate_remap(tables, rect, use_cv2)


def clip_normalize_u8(depth: np.ndarray, z_near: float, z_far: float) -> np.ndarray:
    """python/disp_to_depth.py:7-21.  Numba types ``(val - min) / range`` in float32 and the
    following ``* 255`` (an int64 literal) in float64; the result is truncated to uint8."""
    lo, hi = np.float32(z_near), np.float32(z_far)
    rng = np.float32(hi - lo)
    clipped = np.maximum(np.minimum(depth.astype(np.float32), hi), lo)
    frac = ((clipped - lo) / rng).astype(np.float32)
    val = frac.astype(np.float64) * 255.0
    out = val.astype(np.int64).astype(np.uint8)
    out[depth == 0] = 0
    return out


def turbo_lut_bgr() -> np.ndarray:
    """256 x 3 uint8 BGR table of ``cv2.COLORMAP_TURBO`` (python/disp_to_depth.py:36)."""
    import cv2

    return cv2.applyColorMap(np.arange(256, dtype=np.uint8).reshape(1, 256), cv2.COLORMAP_TURBO).reshape(256, 3)


def colorize(disp_map: np.ndarray, depth_scale: float, z_near: float, z_far: float) -> np.ndarray:
    """python/disp_to_depth.py:99-115 (colorize_depth_from_disp) incl. apply_white_mask :24-31."""
    u8 = clip_normalize_u8(disparity_to_depth(disp_map, depth_scale), z_near, z_far)
    bgr = turbo_lut_bgr()[u8]
    bgr[u8 == 0] = 255
    return bgr


# --------------------------------------------------------------------------------------
# N3  setup-time tables the reference builds itself (no OpenCV involved)
# --------------------------------------------------------------------------------------
def linear_projector_time_map(proj_w: int, proj_h: int, scan_upwards: bool) -> np.ndarray:
    """python/proj_time_map.py:6-19."""
    ys, xs = np.mgrid[0:proj_h, 0:proj_w]
    if scan_upwards:
        ys = ys[::-1]
    return ((xs * proj_h + ys) / (proj_w * proj_h)).astype(np.float32)


def build_x_map(time_map: np.ndarray, x_map_width: int, t_px_scale: int, x_offset: int, num_scanlines: int):
    """python/x_map.py:5-55 (compute_x_map_from_time_map).

    For each (y, t_coord): argmin over x of |t - time_map[y, x]| in float64 (first minimum wins,
    cells equal to 0 are undefined and skipped); accept only if the minimum is
    <= 2 / num_scanlines.  ``t_coord == 0`` is skipped entirely.  Returns (x_map int16, t_diffs f32).
    """
    h, w = time_map.shape
    x_map = np.zeros((h, x_map_width), dtype=np.int16)
    t_diffs = np.zeros((h, x_map_width), dtype=np.float32)
    max_t_diff = 2 / num_scanlines
    tm = time_map.astype(np.float64)
    undefined = time_map == 0
    for t_coord in range(1, x_map_width):
        t = t_coord / t_px_scale
        if t == 0:
            continue
        diff = np.abs(t - tm)
        diff[undefined] = np.inf
        best = np.argmin(diff, axis=1)
        best_val = diff[np.arange(h), best]
        ok = np.isfinite(best_val) & (best_val <= max_t_diff)
        x_map[ok, t_coord] = (best[ok] + x_offset).astype(np.int16)
        t_diffs[ok, t_coord] = best_val[ok].astype(np.float32)
    return x_map, t_diffs


# --------------------------------------------------------------------------------------
# seeded synthetic event generators shared by tests and bench
# --------------------------------------------------------------------------------------
def synth_events(seed: int, n: int, cam_w: int, cam_h: int, frame_us: int = 16666, p_on: float = 0.9, t0: int = 0):
    """Uniform-pixel, time-sorted ("homogeneous Poisson conditioned on N") frame.  The draw order
    is part of the contract: x, y, p, t."""
    rng = np.random.default_rng(seed)
    x = rng.integers(0, cam_w, n)
    y = rng.integers(0, cam_h, n)
    p = rng.random(n) < p_on
    t = np.sort(rng.integers(0, frame_us, n))
    ev = np.zeros(n, dtype=EVENT_DTYPE)
    ev["x"], ev["y"], ev["p"], ev["t"] = x, y, p, t + t0
    return ev


def synth_plane_events(tables: OracleTables, time_map_rect: np.ndarray, z: float, frame_us: int = 16666, repeat: int = 1,
                       jitter_us: int = 0, seed: int = 0, t0: int = 0) -> np.ndarray:
    """The ~100 %-inlier workload (SURVEY.md §8c "sanity oracle", §8d input 1): a fronto-parallel plane at
    depth ``z`` seen by the rig.  It has the shape of the reference's real input
    (python/eval/compute_depth_x_maps.py:83-96: one event per lit camera pixel, ``t`` from a time map):
    every camera pixel (x, y) with rectified coordinates (xcr, ycr) fires when the projector column that
    lights it passes, ``d = round(P2[0,3] / z)``, ``t = rint(time_map_rect[ycr, xcr + d] * frame_us)``;
    pixels whose target lies outside the rectified image or in an undefined (0) part of the time map stay
    dark.  Events come out time-sorted (stable: row-major pixel order inside one microsecond), ``p = 1``.

    ``repeat`` > 1 lets every lit pixel fire ``repeat`` times (an event camera emits a burst per edge),
    each with an extra uniform integer jitter in [-jitter_us, jitter_us] (seeded), clipped to the frame."""
    d = int(np.rint(tables.depth_scale / z))
    ys, xs = np.mgrid[0 : tables.cam_h, 0 : tables.cam_w]
    xcr = tables.lut_x.astype(np.int64)
    ycr = tables.lut_y.astype(np.int64)
    xp = xcr + d
    h, w = time_map_rect.shape
    lit = (ycr >= 0) & (ycr < h) & (xp >= 0) & (xp < w)
    tm = np.zeros(xcr.shape, dtype=np.float64)
    tm[lit] = time_map_rect[ycr[lit], xp[lit]]
    lit &= tm > 0
    t = np.rint(tm[lit] * frame_us).astype(np.int64)
    x, y = xs[lit], ys[lit]
    if repeat > 1:
        rng = np.random.default_rng(seed)
        t = np.repeat(t, repeat)
        x, y = np.repeat(x, repeat), np.repeat(y, repeat)
        if jitter_us > 0:
            t = np.clip(t + rng.integers(-jitter_us, jitter_us + 1, len(t)), 0, frame_us - 1)
    order = np.argsort(t, kind="stable")
    ev = np.zeros(len(t), dtype=EVENT_DTYPE)
    ev["x"], ev["y"], ev["p"], ev["t"] = x[order], y[order], 1, t[order] + t0
    return ev


# --------------------------------------------------------------------------------------
# N4 -- per-frame de-duplication filters (python/frame_event_filter.py:19-128)
#
# The reference scatters `t` (and for the YT filter `x`) into zero-initialised int32 images indexed
# by the event's key and reads the images back through a boolean mask, i.e. the survivors come out in
# row-major key order, `t` wrapped to int32, `p` = 1.  Forward assignment keeps the LAST duplicate.
# The "First..." filters assign reversed VIEWS (`img[y[::-1], x[::-1]] = t[::-1]`) and are meant to keep
# the first duplicate, but NumPy's index iterator negates the negative strides of the three views and
# walks them in memory order again, so the reference AS IT RUNS (NumPy 2.3.5 here; golden vectors in
# tests/golden/stream_filters.npz) keeps the LAST duplicate in those filters too, and the "mean" filter
# averages the last timestamp with itself (with the int32 wrap of the sum).  `as_reference=True`
# (default) reproduces that; `as_reference=False` gives the documented intent (true first event).
# Restated with one stable argsort of the keys instead of dense images.
# --------------------------------------------------------------------------------------
FILTER_NONE, FILTER_FIRST_YT, FILTER_FIRST_XY, FILTER_LAST_XY, FILTER_MEAN_XY = 0, 1, 2, 3, 4


class ActivityNoiseFilterOracle:
    """PARITY UNPINNED.  Metavision's ``ActivityNoiseFilterAlgorithm(width, height, threshold)`` is a closed binary
    (``metavision_sdk_cv``) that is absent here and on the GPU box; the reference constructs it at
    python/depth_reprojection_pipe.py:65-67 (``threshold = int(1e6 / projector_fps)``) and applies it to every packet
    at :116-117.  This restates its published behaviour ("accepts an event if a similar event happened in its
    neighbourhood during the last `threshold` microseconds"), the way the SDK's open-source edition implements it:

      * one timestamp per pixel, initially 0, set to the event's timestamp by EVERY incoming event;
      * an event passes iff one of the 8 neighbours of its pixel (the pixel itself excluded, neighbours outside the
        sensor skipped) holds a timestamp  >  t - threshold;
      * events are processed strictly in stream order and the state is carried from packet to packet.

    Sequential by definition, hence a plain loop: use it on small inputs (~3 us per event)."""

    def __init__(self, width: int, height: int, threshold_us: int):
        self.width, self.height, self.threshold = int(width), int(height), int(threshold_us)
        self.last_ts = np.zeros((self.height, self.width), dtype=np.int64)

    def reset(self):
        self.last_ts[:] = 0

    def process_events(self, events: np.ndarray) -> np.ndarray:
        keep = np.zeros(len(events), dtype=bool)
        xs, ys, ts = events["x"].astype(np.int64), events["y"].astype(np.int64), events["t"].astype(np.int64)
        last, w, h, thr = self.last_ts, self.width, self.height, self.threshold
        for i in range(len(events)):
            x, y, t = int(xs[i]), int(ys[i]), int(ts[i])
            if x >= w or y >= h:
                continue  # (outside the sensor: dropped, no state)
            last[y, x] = np.iinfo(np.int64).min  # the pixel itself is not a witness
            keep[i] = bool((last[max(y - 1, 0):y + 2, max(x - 1, 0):x + 2] > t - thr).any())
            last[y, x] = t
        return events[keep]


def _first_last_per_key(key: np.ndarray):
    """Unique keys (ascending) with the index of their first and last occurrence."""
    order = np.argsort(key, kind="stable")
    ks = key[order]
    start = np.flatnonzero(np.concatenate(([True], ks[1:] != ks[:-1]))) if len(ks) else np.zeros(0, np.int64)
    end = np.concatenate((start[1:], [len(ks)])) - 1 if len(ks) else np.zeros(0, np.int64)
    return ks[start], order[start], order[end]


def _t32(t: np.ndarray) -> np.ndarray:
    """`int32_image[...] = events["t"]`: the int64 timestamp wraps to int32 (frame_event_filter.py:26,52)."""
    return t.astype(np.int64).astype(np.int32)


def frame_event_filter(events: np.ndarray, mode: int, xp_i16: Optional[np.ndarray] = None, as_reference: bool = True) -> np.ndarray:
    """`FrameEventFilter.filter_events(events, xp_i16)` of the five filters
    (frame_event_filter.py:10-128).  Errors of the reference are kept: an empty positive set raises
    ValueError (`.max()` of an empty array, :24); for the YT filter `xp_i16` must be as long as the
    positive events (:78 would raise on the shape mismatch) and a negative column wraps like a NumPy
    index (IndexError if it wraps below 0)."""
    if mode == FILTER_NONE:
        return events  # NoFilter :10-16
    ev = events[events["p"] == 1]  # :21,47,72,104
    if len(ev) == 0:
        raise ValueError("zero-size array to reduction operation maximum which has no identity")
    y = ev["y"].astype(np.int64)
    if mode == FILTER_FIRST_YT:
        xp = np.asarray(xp_i16)
        if len(xp) != len(ev):
            raise IndexError("shape mismatch: indexing arrays could not be broadcast together")
        width = int(xp.max()) + 1  # :75
        col = xp.astype(np.int64)
        col = np.where(col < 0, col + width, col)
        if width <= 0 or (col < 0).any():
            raise IndexError("index out of bounds for the (y, x_rect) image")
        keys, first, last = _first_last_per_key(y * width + col)
        if as_reference:
            first = last
        out = np.zeros(len(keys), dtype=events.dtype)
        out["t"] = _t32(ev["t"][first])  # :83 (reversed assignment: first event wins)
        out["x"] = ev["x"][first]        # :79
        out["y"] = keys // width
        out["p"] = 1
        return out
    width = int(ev["x"].max()) + 1
    keys, first, last = _first_last_per_key(y * width + ev["x"].astype(np.int64))
    if as_reference:
        first = last
    out = np.zeros(len(keys), dtype=events.dtype)
    if mode == FILTER_LAST_XY:       # :19-41
        out["t"] = _t32(ev["t"][last])
    elif mode == FILTER_FIRST_XY:    # :44-66
        out["t"] = _t32(ev["t"][first])
    elif mode == FILTER_MEAN_XY:     # :101-128: int32 add (wraps), floor division
        out["t"] = (_t32(ev["t"][last]).astype(np.int64) + _t32(ev["t"][first]).astype(np.int64)).astype(np.int32) // 2
    else:
        raise ValueError(f"unknown filter mode {mode}")
    out["x"] = keys % width
    out["y"] = keys // width
    out["p"] = 1
    return out


# --------------------------------------------------------------------------------------
# N2 -- frame segmentation (python/trigger_finder.py:93-189)
# --------------------------------------------------------------------------------------
MIN_EVENTS_PER_FRAME = 1000  # trigger_finder.py:8


def find_trigger(t: np.ndarray, projector_fps: float, pause_thresh_us: int = 40, min_events: int = MIN_EVENTS_PER_FRAME):
    """Decision of `RobustTriggerFinder.find_trigger` (:146-189) on the timestamps of the popped
    buffer.  Returns (status, prev_idx, next_idx, n_pauses):
      status  1: frame = events[prev_idx + 2 : next_idx - 2], events[next_idx - 2 :] go back to the buffer
      status  0: candidate rejected, events[next_idx :] go back to the buffer
      status -1: no candidate, nothing goes back (the reference has already popped the buffer)."""
    t = np.asarray(t).astype(np.int64)
    pauses = np.flatnonzero(t[1:] - t[:-1] >= pause_thresh_us) if len(t) > 1 else np.zeros(0, np.int64)  # :155
    frame_us = 1e6 / projector_fps
    for k in range(len(pauses) - 1):  # :161 consecutive pauses
        prev_idx, next_idx = int(pauses[k]), int(pauses[k + 1])
        gap = int(t[next_idx] - t[prev_idx])
        if gap > frame_us / 2:  # :168
            if gap <= frame_us and next_idx - prev_idx > min_events:  # :169
                return 1, prev_idx, next_idx, len(pauses)
            return 0, prev_idx, next_idx, len(pauses)  # :184-187
    return -1, -1, -1, len(pauses)


class TriggerFinderOracle:
    """`RobustTriggerFinder` + `EventBufferList` (:11-145) on plain NumPy chunks: same buffering,
    frame-drop and span rules; `frame_callback(events)` fires once per accepted frame."""

    def __init__(self, projector_fps, frame_callback, pause_thresh_us: int = 40):
        self.projector_fps = projector_fps
        self.frame_callback = frame_callback
        self.pause_thresh_us = pause_thresh_us
        self.should_drop = False
        self.last_frame_start_us = -1
        self.chunks = []
        self.log = []  # (status, start_time) per find_trigger call

    # EventBufferList ---------------------------------------------------------------------
    def _first_t(self):
        return int(self.chunks[0]["t"][0]) if self.chunks else -1

    def _last_t(self):
        return int(self.chunks[-1]["t"][-1]) if self.chunks else -1

    def _drop(self, drop_len_ms):  # :63-75
        until = self._first_t() + drop_len_ms * 1000
        dropped = False
        while self.chunks and self._first_t() < until:
            self.chunks.pop(0)
            dropped = True
        return dropped

    # RobustTriggerFinder -----------------------------------------------------------------
    def reset(self):  # :111-114
        self.chunks.clear()
        self.should_drop = False
        self.last_frame_start_us = -1

    def drop_frame(self):
        self.should_drop = True

    def process_events(self, evs):  # :119-144
        if len(evs):
            self.chunks.append(evs)
        if self.should_drop:
            if self._drop(1e3 / self.projector_fps):
                self.should_drop = False
            else:
                return
        if not self.chunks:
            return
        first, last = self._first_t(), self._last_t()
        span = -1 if first < 0 or last < 0 else last - first  # :52-60
        if span < 1e6 / self.projector_fps:
            return
        evs = np.concatenate(self.chunks)
        self.chunks.clear()
        status, prev_idx, next_idx, _ = find_trigger(evs["t"], self.projector_fps, self.pause_thresh_us)
        start_time = -1
        if status == 1:
            self.frame_callback(evs[prev_idx + 2 : next_idx - 2])
            start_time = int(evs["t"][prev_idx + 2])
            self.last_frame_start_us = start_time
            rest = evs[next_idx - 2 :]
        elif status == 0:
            rest = evs[next_idx:]
        else:
            rest = evs[:0]
        if len(rest):
            self.chunks.append(rest)
        self.log.append((status, start_time))


def synth_projector_stream(seed: int, n_frames: int, events_per_frame: int, cam_w: int, cam_h: int, projector_fps: float = 60.2,
                           duty: float = 0.95, glitch_every: int = 0) -> np.ndarray:
    """A continuous stream as the laser projector produces it: per frame period a burst of events over
    `duty` of the period (sorted, gaps well below the 40 us pause threshold), then a pause.  The true
    rate sits slightly above the nominal 60 fps, as it must for the reference's `gap <= 1e6 / fps` test
    (:169) to accept frames.  With `glitch_every` > 0 every such frame is cut short (a rejected
    candidate for the trigger finder)."""
    rng = np.random.default_rng(seed)
    period = 1e6 / projector_fps
    chunks = []
    for f in range(n_frames):
        n = events_per_frame
        active = duty * period
        if glitch_every and f % glitch_every == glitch_every - 1:
            active *= 0.45
            n = max(4, n // 2)
        t0 = int(round(f * period)) + 100
        t = t0 + np.floor(np.arange(n, dtype=np.float64) * (active / n)).astype(np.int64) + rng.integers(0, 3, n)
        ev = np.zeros(n, dtype=EVENT_DTYPE)
        ev["x"] = rng.integers(0, cam_w, n)
        ev["y"] = rng.integers(0, cam_h, n)
        ev["p"] = 1
        ev["t"] = np.sort(t)
        chunks.append(ev)
    out = np.zeros(sum(len(c) for c in chunks), dtype=EVENT_DTYPE)  # (np.concatenate would repack the 16-byte records)
    i = 0
    for c in chunks:
        out[i : i + len(c)] = c
        i += len(c)
    return out
