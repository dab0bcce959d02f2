"""Disparity map -> projector view -> metric depth -> colour.

Mirrors the reference's ``disp_to_depth`` module (/root/reference/python/disp_to_depth.py):
``DisparityToDepth.remap_rectified_disp_map_to_proj`` (:76-97), ``colorize_depth_from_disp``
(:99-115) and the free function ``disparity_to_depth_rectified`` (:46-63).  When handed the lazy
handles of the earlier stages these calls launch the fused kernels (one pass over the event
buffer + one per-pixel epilogue); handed materialised maps they run the per-stage kernels.
"""
from __future__ import annotations

from contextlib import nullcontext
from dataclasses import dataclass
from typing import Any

import numpy as np
import torch

from .lazy import DeviceArray, LazyDispMap, engine_for, to_tensor


def _measure(stats, key):
    return stats.measure_time(key) if stats is not None and hasattr(stats, "measure_time") else nullcontext()


def disparity_to_depth_rectified(disparity, P1):
    """Reference :46-63: ``depth = 0 if d == 0 else max(P1[0, 3] / d, 1e-9)`` (float64 divide,
    float32 result).  Returns a device array."""
    from .engine import OUT_DEPTH

    scale = float(np.asarray(P1)[0, 3])
    if isinstance(disparity, LazyDispMap) and disparity._value is None and disparity.stage != "rect":
        eng = disparity.ticket.engine
        if scale == eng.depth_scale:
            return DeviceArray.of(disparity.fused(OUT_DEPTH))
    if isinstance(disparity, DeviceArray):
        t = disparity.tensor
    elif isinstance(disparity, torch.Tensor):
        t = disparity
    else:
        t = torch.from_numpy(np.ascontiguousarray(disparity, dtype=np.float32)).cuda()
    eng = engine_for(t.device)
    return DeviceArray.of(eng.disp_to_depth(t.to(torch.float32), scale))


@dataclass
class DisparityToDepth:
    stats: Any
    calib_params: Any
    calib_maps: Any
    z_near: float
    z_far: float
    # frames handed to ``frame_callback`` are host uint8 arrays, as in the reference
    return_host: bool = True

    def __post_init__(self):
        self.dilate_kernel = np.ones((7, 7), dtype=np.uint8)  # reference :74 (size fixed in the engine tables)

    def remap_rectified_disp_map_to_proj(self, rectified_disp_map):
        m = rectified_disp_map
        with _measure(self.stats, "dilate"), _measure(self.stats, "remap disp"):
            if isinstance(m, LazyDispMap) and m._value is None and m.stage == "rect":
                return LazyDispMap(m.ticket, "proj")
            eng = self.calib_maps.engine()
            return DeviceArray.of(eng.dilate_remap(to_tensor(m, eng.device, torch.float32)))

    def colorize_depth_from_disp(self, disp_map):
        from .engine import OUT_BGR

        with _measure(self.stats, "d2d_rect"), _measure(self.stats, "clip_norm"), _measure(self.stats, "color_map"):
            if isinstance(disp_map, LazyDispMap) and disp_map._value is None and disp_map.stage != "rect":
                bgr = disp_map.fused(OUT_BGR, z_near=self.z_near, z_far=self.z_far)
            else:
                eng = self.calib_maps.engine()
                bgr = eng.colorize(to_tensor(disp_map, eng.device, torch.float32), self.z_near, self.z_far, float(self.calib_maps.P2[0, 3]))
            return bgr.cpu().numpy() if self.return_host else bgr
