"""``DepthEngine`` — one device context of the fused X-maps depth path per GPU.

Thin Python layer over the C ABI (``include/xmaps_b200.h``): torch supplies device memory and the
current CUDA stream, every computation happens in the hand-written kernels of ``csrc/``.  There is
no CPU or PyTorch fallback; constructing an engine without a CUDA device raises.

Reference call chain this replaces (per projector frame):
/root/reference/python/depth_reprojection_pipe.py:121-167.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

from . import _native as N
from .events import DeviceEvents

VIEW_PROJECTOR, VIEW_CAMERA = N.VIEW_PROJECTOR, N.VIEW_CAMERA
OUT_DEPTH, OUT_DISPARITY, OUT_BGR = N.OUT_DEPTH, N.OUT_DISPARITY, N.OUT_BGR
TBOUNDS_REDUCE, TBOUNDS_SORTED, TBOUNDS_GIVEN = N.TBOUNDS_REDUCE, N.TBOUNDS_SORTED, N.TBOUNDS_GIVEN


@dataclass
class TableSet:
    """Host-side read-only tables of one calibration (what the reference keeps on CamProjMaps /
    XMapsDisparity / DisparityToDepth)."""

    lut_x: np.ndarray  # [cam_h, cam_w] int16      disp_cam_mapx_i16
    lut_y: np.ndarray  # [cam_h, cam_w] int16      disp_cam_mapy_i16
    x_map: Optional[np.ndarray]  # [rect_h, xmap_w] int16  proj_x_map (None: set later)
    remap_xy: Optional[np.ndarray]  # [proj_h, proj_w, 2] int16  disp_proj_mapxy_i16
    rect_w: int
    rect_h: int
    t_px_scale: int
    x_offset: int
    depth_scale: float  # P2[0, 3]
    dilate: int = 7
    lut_x_f32: Optional[np.ndarray] = None
    lut_y_f32: Optional[np.ndarray] = None


def turbo_bgr_table() -> np.ndarray:
    """256 x 3 uint8 BGR entries of ``cv2.COLORMAP_TURBO`` (reference: disp_to_depth.py:36)."""
    import cv2

    return np.ascontiguousarray(
        cv2.applyColorMap(np.arange(256, dtype=np.uint8).reshape(1, 256), cv2.COLORMAP_TURBO).reshape(256, 3)
    )


def _c16(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


class DepthEngine:
    def __init__(self, tables: TableSet, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("xmaps_b200 needs a CUDA device: the depth path has no CPU fallback")
        dev = torch.device("cuda" if device is None else device)
        if dev.type != "cuda":
            raise ValueError("DepthEngine device must be a CUDA device")
        self.device = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
        self.device_index = self.device.index
        torch.cuda.init()
        with torch.cuda.device(self.device):
            torch.cuda.current_stream()  # makes sure the primary context exists

        self.cam_h, self.cam_w = tables.lut_x.shape
        self.rect_w, self.rect_h = int(tables.rect_w), int(tables.rect_h)
        if tables.remap_xy is not None:
            self.proj_h, self.proj_w = tables.remap_xy.shape[:2]
        else:
            self.proj_h = self.proj_w = 0
        self.depth_scale = float(tables.depth_scale)
        self.dilate = int(tables.dilate)

        keep = []  # host arrays must outlive xm_ctx_create

        def ptr(a, dtype):
            if a is None:
                return None
            arr = _c16(a, dtype)
            keep.append(arr)
            return arr.ctypes.data

        t = N.XmTables()
        t.cam_w, t.cam_h = self.cam_w, self.cam_h
        t.rect_w, t.rect_h = self.rect_w, self.rect_h
        t.proj_w, t.proj_h = self.proj_w, self.proj_h
        t.xmap_w = 0 if tables.x_map is None else tables.x_map.shape[1]
        t.t_px_scale, t.x_offset, t.dilate = int(tables.t_px_scale), int(tables.x_offset), self.dilate
        t.depth_scale = self.depth_scale
        if tables.x_map is not None and tables.x_map.shape[0] != self.rect_h:
            raise ValueError("x_map must have rect_h rows")
        t.lut_x, t.lut_y = ptr(tables.lut_x, np.int16), ptr(tables.lut_y, np.int16)
        t.x_map = ptr(tables.x_map, np.int16)
        t.remap_xy = ptr(tables.remap_xy, np.int16)
        t.lut_x_f32, t.lut_y_f32 = ptr(tables.lut_x_f32, np.float32), ptr(tables.lut_y_f32, np.float32)
        handle = C.c_void_p()
        N.check(N.lib.xm_ctx_create(C.byref(t), self.device_index, C.byref(handle)))
        self._ctx = handle
        turbo = turbo_bgr_table()
        N.check(N.lib.xm_ctx_set_colormap(self._ctx, turbo.ctypes.data))

    # ------------------------------------------------------------------ life cycle / options
    def close(self):
        if getattr(self, "_ctx", None):
            N.lib.xm_ctx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_x_map(self, x_map: np.ndarray, t_px_scale: int, x_offset: int):
        a = _c16(x_map, np.int16)
        N.check(N.lib.xm_ctx_set_xmap(self._ctx, a.ctypes.data, a.shape[0], a.shape[1], int(t_px_scale), int(x_offset)))

    def set_option(self, key: str, value: int):
        N.check(N.lib.xm_ctx_set_option(self._ctx, key.encode(), int(value)))

    @staticmethod
    def launch_count() -> int:
        """Kernels launched by the library in this process so far."""
        return N.launch_count()

    def get_option(self, key: str) -> int:
        v = C.c_int64()
        N.check(N.lib.xm_ctx_get_option(self._ctx, key.encode(), C.byref(v)))
        return int(v.value)

    # ------------------------------------------------------------------ helpers
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def events(self, events, time_f64: Optional[bool] = None) -> DeviceEvents:
        ev = DeviceEvents.from_any(events, device=self.device, time_f64=time_f64)
        if ev.device != self.device:
            raise ValueError(f"events live on {ev.device}, engine on {self.device}")
        return ev

    def out_shape(self, view: int, output: int):
        h, w = (self.cam_h, self.cam_w) if view == VIEW_CAMERA else (self.proj_h, self.proj_w)
        return (h, w, 3) if output == OUT_BGR else (h, w)

    def _args(self, ev: DeviceEvents, view, output, time_bounds, polarity, t_min, t_max, z_near, z_far, out_ptr):
        a = N.XmFrameArgs()
        a.d_events = ev.raw.data_ptr() if len(ev) else None
        a.n_events = len(ev)
        a.flags = (N.FLAG_POLARITY if polarity else 0) | (N.FLAG_TIME_F64 if ev.time_f64 else 0)
        a.view, a.time_bounds, a.output = int(view), int(time_bounds), int(output)
        if time_bounds == TBOUNDS_GIVEN:
            if t_min is None or t_max is None:
                raise ValueError("TBOUNDS_GIVEN needs t_min and t_max")
            if ev.time_f64:
                a.t_min = int(np.float64(t_min).view(np.int64))
                a.t_max = int(np.float64(t_max).view(np.int64))
            else:
                a.t_min, a.t_max = int(t_min), int(t_max)
        a.d_out = out_ptr
        a.z_near, a.z_far = float(z_near), float(z_far)
        return a

    def _alloc_out(self, view, output, out):
        shape = self.out_shape(view, output)
        dtype = torch.uint8 if output == OUT_BGR else torch.float32
        if out is None:
            return torch.empty(shape, dtype=dtype, device=self.device)
        if out.shape != torch.Size(shape) or out.dtype != dtype or not out.is_contiguous() or out.device != self.device:
            raise ValueError(f"out must be a contiguous {dtype} tensor of shape {shape} on {self.device}")
        return out

    # ------------------------------------------------------------------ the hot path
    def frame(
        self,
        events,
        view: int = VIEW_PROJECTOR,
        output: int = OUT_DEPTH,
        time_bounds: int = TBOUNDS_SORTED,
        polarity: bool = True,
        t_min=None,
        t_max=None,
        z_near: float = 0.1,
        z_far: float = 1.0,
        out: Optional[torch.Tensor] = None,
        bilinear: bool = False,
    ) -> torch.Tensor:
        """One projector frame of events -> depth / disparity / BGR frame (asynchronous on the
        current CUDA stream of the engine's device).  ``bilinear``: opt-in bilinear X-map lookup at the un-rounded
        rectified row / time column (``XM_FLAG_BILINEAR``; not reference behaviour, whose lookup is nearest)."""
        ev = self.events(events)
        out = self._alloc_out(view, output, out)
        a = self._args(ev, view, output, time_bounds, polarity, t_min, t_max, z_near, z_far, out.data_ptr())
        if bilinear:
            a.flags |= N.FLAG_BILINEAR
        N.check(N.lib.xm_frame(self._ctx, C.byref(a), self._stream()))
        return out

    def frame_batch(
        self,
        frames: Sequence,
        view: int = VIEW_PROJECTOR,
        output: int = OUT_DEPTH,
        time_bounds: int = TBOUNDS_SORTED,
        polarity: bool = True,
        z_near: float = 0.1,
        z_far: float = 1.0,
        out: Optional[torch.Tensor] = None,
        t_bounds: Optional[Sequence] = None,
        out_ptrs: Optional[Sequence[int]] = None,
    ) -> Optional[torch.Tensor]:
        """Independent frames on one stream -> ``[n_frames, ...]``.  Uniform batches (one view / output,
        integer timestamps) are rendered by ONE persistent kernel per 32 frames (``xm_frame_batch``): the
        epilogue of a frame overlaps the event stream of the next one.  ``t_bounds``: per-frame
        ``(t_min, t_max)`` for ``TBOUNDS_GIVEN``.  ``out_ptrs``: raw device addresses, one per frame, instead of
        ``out`` -- e.g. peer memory mapped with ``xm_peer_open``, so that the kernels write finished frames
        straight into another GPU's buffer (``sharding.FrameSharder``); nothing is returned then."""
        evs = [self.events(f) for f in frames]
        shape = self.out_shape(view, output)
        dtype = torch.uint8 if output == OUT_BGR else torch.float32
        if out_ptrs is not None:
            if len(out_ptrs) != len(evs) or out is not None:
                raise ValueError("out_ptrs needs one address per frame and excludes out")
        elif out is None:
            out = torch.empty((len(evs),) + tuple(shape), dtype=dtype, device=self.device)
        elif tuple(out.shape) != (len(evs),) + tuple(shape) or out.dtype != dtype or not out.is_contiguous():
            raise ValueError("out has the wrong shape / dtype")
        arr = (N.XmFrameArgs * max(1, len(evs)))()
        for i, ev in enumerate(evs):
            lo, hi = t_bounds[i] if t_bounds is not None else (None, None)
            dst = int(out_ptrs[i]) if out_ptrs is not None else out[i].data_ptr()
            arr[i] = self._args(ev, view, output, time_bounds, polarity, lo, hi, z_near, z_far, dst)
        N.check(N.lib.xm_frame_batch(self._ctx, arr, len(evs), self._stream()))
        return out

    def status(self) -> dict:
        """Device-side findings of the most recent frame (synchronises the current stream)."""
        st = N.XmFrameStatus()
        N.check(N.lib.xm_frame_status(self._ctx, C.byref(st), self._stream()))
        return {
            "n_valid": int(st.n_valid),
            "n_inliers": int(st.n_inliers),
            "t_min": int(st.t_min),
            "t_max": int(st.t_max),
            "flags": int(st.flags),
            "fixup_ran": bool(st.fixup_ran),
            "tbounds_violated": bool(st.flags & N.STATUS_TBOUNDS_VIOLATED),
            "pixel_oob": bool(st.flags & N.STATUS_PIXEL_OOB),
            "scatter_oob": bool(st.flags & N.STATUS_SCATTER_OOB),
            "filter_polarity": bool(st.flags & N.STATUS_FILTER_POLARITY),
            "filter_index": bool(st.flags & N.STATUS_FILTER_INDEX),
        }

    # ------------------------------------------------------------------ the rows next to the path
    def polarity_filter(self, events) -> DeviceEvents:
        """``events[events["p"] == 1]`` on the device, order preserved (the reference's Metavision
        ``PolarityFilterAlgorithm(1)``, depth_reprojection_pipe.py:43,114)."""
        ev = self.events(events)
        n = len(ev)
        out = torch.empty((max(n, 1), 4), dtype=torch.int32, device=self.device)
        count = torch.zeros(1, dtype=torch.int64, device=self.device)
        N.check(N.lib.xm_polarity_filter(self._ctx, ev.raw.data_ptr() if n else None, n, out.data_ptr(), count.data_ptr(), self._stream()))
        return DeviceEvents(out[: int(count.item())], ev.time_f64)

    def activity_filter(self, events, threshold_us: int) -> DeviceEvents:
        """The reference's Metavision ``ActivityNoiseFilterAlgorithm(w, h, threshold_us).process_events``
        (depth_reprojection_pipe.py:65-67,116-117) on the device: keeps an event iff one of the 8 neighbours of its
        pixel fired less than ``threshold_us`` earlier; the per-pixel timestamps live in the context and are carried
        from packet to packet (``activity_reset``).  Semantics restated from the SDK's documentation (closed binary)."""
        ev = self.events(events)
        n = len(ev)
        if ev.time_f64:
            raise ValueError("the activity filter works on integer timestamps")
        out = torch.empty((max(n, 1), 4), dtype=torch.int32, device=self.device)
        count = torch.zeros(1, dtype=torch.int64, device=self.device)
        N.check(N.lib.xm_activity_filter(self._ctx, ev.raw.data_ptr() if n else None, n, int(threshold_us), out.data_ptr(), count.data_ptr(), self._stream()))
        return DeviceEvents(out[: int(count.item())], False)

    def activity_reset(self):
        N.check(N.lib.xm_activity_reset(self._ctx, self._stream()))

    def filter_events(self, events, mode: int, x_rect: Optional[torch.Tensor] = None, as_reference: bool = True) -> DeviceEvents:
        """One survivor per key (``frame_event_filter.py``'s filters, ``N.FILTER_*``) as a new device
        event buffer in row-major key order; see ``xm_filter_events`` in the header."""
        ev = self.events(events)
        n = len(ev)
        if ev.time_f64:
            raise ValueError("the frame filters work on integer timestamps")
        if mode == N.FILTER_FIRST_YT:
            if x_rect is None or x_rect.dtype != torch.int16 or not x_rect.is_contiguous() or x_rect.numel() != n:
                raise ValueError("FILTER_FIRST_YT needs x_rect: contiguous int16, one entry per event")
        out = torch.empty((max(n, 1), 4), dtype=torch.int32, device=self.device)
        count = torch.zeros(1, dtype=torch.int64, device=self.device)
        N.check(
            N.lib.xm_filter_events(
                self._ctx, ev.raw.data_ptr() if n else None, n, int(mode), x_rect.data_ptr() if (x_rect is not None and n) else None,
                1 if as_reference else 0, out.data_ptr(), count.data_ptr(), self._stream(),
            )
        )
        k = int(count.item())
        return DeviceEvents(out[:k], False)

    def find_trigger(self, events, projector_fps: float, pause_thresh_us: int = 40, min_events: int = 1000):
        """``RobustTriggerFinder.find_trigger``'s decision on one device buffer ->
        ``(status, prev_idx, next_idx, n_pauses, start_time, end_time)`` (one 64-byte read-back)."""
        ev = self.events(events)
        n = len(ev)
        res = torch.empty(8, dtype=torch.int64, device=self.device)
        N.check(
            N.lib.xm_find_trigger(
                self._ctx, ev.raw.data_ptr() if n else None, n, int(pause_thresh_us), 1e6 / float(projector_fps), int(min_events),
                res.data_ptr(), self._stream(),
            )
        )
        return tuple(int(v) for v in res.cpu().tolist()[:6])

    def frame_host(
        self,
        events: np.ndarray,
        view: int = VIEW_PROJECTOR,
        output: int = OUT_DEPTH,
        time_bounds: int = TBOUNDS_SORTED,
        polarity: bool = True,
        z_near: float = 0.1,
        z_far: float = 1.0,
        out: Optional[np.ndarray] = None,
    ) -> np.ndarray:
        """Host buffers in, host frame out (H2D copy, kernels, D2H copy, synchronise) — the call a
        CPU-side user of the reference makes.  ``events`` / ``out`` may be pinned (torch
        ``pin_memory`` / ``xm_host_alloc``) for full PCIe speed."""
        if isinstance(events, torch.Tensor):
            if events.is_cuda:
                raise ValueError("frame_host takes host buffers; use frame() for CUDA tensors")
            n = events.numel() * events.element_size() // 16
            ev_ptr, time_f64 = events.data_ptr(), False
        else:
            events = np.ascontiguousarray(events)
            if events.dtype.itemsize != 16:
                raise ValueError("host events must be 16-byte EventCD records")
            n = events.shape[0]
            ev_ptr = events.ctypes.data
            time_f64 = events.dtype.names is not None and np.issubdtype(events.dtype["t"], np.floating)
        shape = self.out_shape(view, output)
        dtype = np.uint8 if output == OUT_BGR else np.float32
        if out is None:
            out = np.empty(shape, dtype=dtype)
        if isinstance(out, torch.Tensor):
            out_ptr = out.data_ptr()
        else:
            if out.shape != tuple(shape) or out.dtype != dtype or not out.flags.c_contiguous:
                raise ValueError("out has the wrong shape / dtype")
            out_ptr = out.ctypes.data
        a = N.XmFrameArgs()
        a.n_events = n
        a.flags = (N.FLAG_POLARITY if polarity else 0) | (N.FLAG_TIME_F64 if time_f64 else 0)
        a.view, a.time_bounds, a.output = int(view), int(time_bounds), int(output)
        a.z_near, a.z_far = float(z_near), float(z_far)
        N.check(N.lib.xm_frame_host(self._ctx, C.byref(a), ev_ptr if n else None, out_ptr, None, self._stream()))
        return out

    # ------------------------------------------------------------------ stage by stage
    def rectify_i16(self, events):
        ev = self.events(events)
        n = len(ev)
        x = torch.empty(n, dtype=torch.int16, device=self.device)
        y = torch.empty(n, dtype=torch.int16, device=self.device)
        N.check(N.lib.xm_rectify_i16(self._ctx, ev.raw.data_ptr() if n else None, n, x.data_ptr(), y.data_ptr(), self._stream()))
        return x, y

    def rectify_f32(self, events):
        ev = self.events(events)
        n = len(ev)
        x = torch.empty(n, dtype=torch.float32, device=self.device)
        y = torch.empty(n, dtype=torch.float32, device=self.device)
        N.check(N.lib.xm_rectify_f32(self._ctx, ev.raw.data_ptr() if n else None, n, x.data_ptr(), y.data_ptr(), self._stream()))
        return x, y

    def event_disparity(self, events, x_rect=None, y_rect=None, time_bounds: int = TBOUNDS_REDUCE, polarity: bool = False):
        """Un-compacted per-event disparity (−1 where not an inlier) and the inlier mask."""
        ev = self.events(events)
        n = len(ev)
        disp = torch.empty(n, dtype=torch.int16, device=self.device)
        mask = torch.empty(n, dtype=torch.uint8, device=self.device)
        a = self._args(ev, VIEW_PROJECTOR, OUT_DISPARITY, time_bounds, polarity, None, None, 0.0, 0.0, None)
        xp = x_rect.data_ptr() if x_rect is not None and n else None
        yp = y_rect.data_ptr() if y_rect is not None and n else None
        for tns in (x_rect, y_rect):
            if tns is not None and (tns.dtype != torch.int16 or not tns.is_contiguous() or tns.numel() != n):
                raise ValueError("x_rect / y_rect must be contiguous int16 tensors with one entry per event")
        N.check(N.lib.xm_event_disparity(self._ctx, C.byref(a), xp, yp, disp.data_ptr(), mask.data_ptr(), self._stream()))
        return disp, mask

    def compact_i16(self, vals: torch.Tensor, mask: torch.Tensor):
        """``vals[mask]`` (order preserved) on the device; returns the compacted tensor."""
        n = vals.numel()
        out = torch.empty(n, dtype=torch.int16, device=self.device)
        count = torch.zeros(1, dtype=torch.int64, device=self.device)
        N.check(
            N.lib.xm_compact_i16(
                self._ctx, vals.data_ptr() if n else None, mask.data_ptr() if n else None, n, out.data_ptr() if n else None,
                count.data_ptr(), self._stream()
            )
        )
        return out[: int(count.item())]

    def scatter_last_wins(self, rows: torch.Tensor, cols: torch.Tensor, vals: torch.Tensor, h: int, w: int) -> torch.Tensor:
        n = vals.numel()
        out = torch.empty((h, w), dtype=torch.float32, device=self.device)
        N.check(
            N.lib.xm_scatter_last_wins(
                self._ctx, rows.data_ptr() if n else None, cols.data_ptr() if n else None, vals.data_ptr() if n else None, n,
                h, w, out.data_ptr(), self._stream()
            )
        )
        return out

    def dilate_remap(self, rect_map: torch.Tensor) -> torch.Tensor:
        if tuple(rect_map.shape) != (self.rect_h, self.rect_w) or rect_map.dtype != torch.float32:
            raise ValueError("rect_map must be float32 [rect_h, rect_w]")
        rect_map = rect_map.contiguous()
        out = torch.empty((self.proj_h, self.proj_w), dtype=torch.float32, device=self.device)
        N.check(N.lib.xm_dilate_remap(self._ctx, rect_map.data_ptr(), out.data_ptr(), self._stream()))
        return out

    def disp_to_depth(self, disp_map: torch.Tensor, depth_scale: Optional[float] = None) -> torch.Tensor:
        d = disp_map.contiguous()
        if d.dtype != torch.float32:
            raise ValueError("disparity map must be float32")
        out = torch.empty_like(d)
        scale = self.depth_scale if depth_scale is None else float(depth_scale)
        N.check(N.lib.xm_disp_to_depth(self._ctx, d.data_ptr(), d.numel(), scale, out.data_ptr(), self._stream()))
        return out

    def colorize(self, disp_map: torch.Tensor, z_near: float, z_far: float, depth_scale: Optional[float] = None) -> torch.Tensor:
        d = disp_map.contiguous()
        if d.dtype != torch.float32:
            raise ValueError("disparity map must be float32")
        out = torch.empty(tuple(d.shape) + (3,), dtype=torch.uint8, device=self.device)
        scale = self.depth_scale if depth_scale is None else float(depth_scale)
        N.check(N.lib.xm_colorize(self._ctx, d.data_ptr(), d.numel(), scale, float(z_near), float(z_far), out.data_ptr(), self._stream()))
        return out

    def point_cloud(self, x: torch.Tensor, y: torch.Tensor, disp: torch.Tensor, Q: np.ndarray) -> torch.Tensor:
        n = x.numel()
        x, y, disp = (t.contiguous().to(torch.float32) for t in (x, y, disp))
        q = np.ascontiguousarray(Q, dtype=np.float64)
        out = torch.empty((n, 3), dtype=torch.float32, device=self.device)
        N.check(
            N.lib.xm_point_cloud(
                self._ctx, x.data_ptr() if n else None, y.data_ptr() if n else None, disp.data_ptr() if n else None, n,
                q.ctypes.data, out.data_ptr() if n else None, self._stream()
            )
        )
        return out


def build_x_map(time_map_rect, x_map_width: int, t_px_scale: int, x_offset: int, num_scanlines: int, device=None):
    """compute_x_map_from_time_map (/root/reference/python/x_map.py:5-55) on the GPU.
    Returns ``(x_map int16 [H, x_map_width], t_diffs float32)`` as CUDA tensors."""
    if not torch.cuda.is_available():
        raise RuntimeError("xmaps_b200 needs a CUDA device")
    dev = torch.device("cuda" if device is None else device)
    dev = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
    tm = torch.as_tensor(np.ascontiguousarray(time_map_rect, dtype=np.float32) if not isinstance(time_map_rect, torch.Tensor) else time_map_rect)
    tm = tm.to(dev, torch.float32).contiguous()
    h, w = tm.shape
    x_map = torch.empty((h, x_map_width), dtype=torch.int16, device=dev)
    t_diffs = torch.empty((h, x_map_width), dtype=torch.float32, device=dev)
    N.check(
        N.lib.xm_build_xmap(
            dev.index, tm.data_ptr(), h, w, int(x_map_width), int(t_px_scale), int(x_offset), int(num_scanlines),
            x_map.data_ptr(), t_diffs.data_ptr(), torch.cuda.current_stream(dev).cuda_stream
        )
    )
    return x_map, t_diffs


def build_inverse_lut(K, D, R, P, size, device=None, with_i16: bool = False):
    """initUndistortRectifyMapInverse (/root/reference/python/cam_proj_calibration.py:31-41) on the GPU: the rectified
    coordinates of every pixel of an unrectified ``size = (W, H)`` image, i.e. ``cv2.undistortPoints`` over the whole
    grid, bit-identical to OpenCV's float32 result.  Returns ``(map_x, map_y)`` float32 CUDA tensors [H, W] (and the
    interleaved int16 table [H, W, 2] = ``mapf_to_i16`` of both with ``with_i16``)."""
    import cv2

    if not torch.cuda.is_available():
        raise RuntimeError("xmaps_b200 needs a CUDA device")
    dev = torch.device("cuda" if device is None else device)
    dev = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
    w, h = int(size[0]), int(size[1])
    k = np.ascontiguousarray(np.asarray(K, dtype=np.float64).reshape(3, 3))
    d = np.ascontiguousarray(np.asarray(D, dtype=np.float64).reshape(-1)) if D is not None else np.zeros(0)
    # RR = P[:, :3] * R exactly as cvUndistortPointsInternal forms it (cvMatMul = cv::gemm)
    rr = np.ascontiguousarray(cv2.gemm(np.ascontiguousarray(np.asarray(P, dtype=np.float64)[:3, :3]), np.ascontiguousarray(np.asarray(R, dtype=np.float64)), 1, None, 0))
    mx = torch.empty((h, w), dtype=torch.float32, device=dev)
    my = torch.empty((h, w), dtype=torch.float32, device=dev)
    xy = torch.empty((h, w, 2), dtype=torch.int16, device=dev) if with_i16 else None
    N.check(
        N.lib.xm_build_inverse_lut(
            dev.index, k.ctypes.data, d.ctypes.data if d.size else None, int(d.size), rr.ctypes.data, w, h,
            mx.data_ptr(), my.data_ptr(), xy.data_ptr() if xy is not None else None, torch.cuda.current_stream(dev).cuda_stream,
        )
    )
    return (mx, my, xy) if with_i16 else (mx, my)
