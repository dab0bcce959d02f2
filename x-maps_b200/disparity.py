"""X-map ownership and the per-event disparity lookup.

Mirrors the reference's ``x_maps_disparity`` module (/root/reference/python/x_maps_disparity.py):
``XMapsDisparity.__post_init__`` (:44-67) builds the X-map from the rectified projector time map
(here with the GPU builder kernel, csrc ``build_xmap_kernel``, bit-identical to the reference's
Numba loop python/x_map.py:5-55) and ``compute_event_disparity`` (:69-82) looks events up in it.
"""
from __future__ import annotations

from dataclasses import InitVar, dataclass, field

import numpy as np

from .calibration import CamProjCalibrationParams, CamProjMaps

X_OFFSET = 4242  # x' = x + X_OFFSET so that 0 can mean "undefined" (reference :49)


@dataclass
class XMapsDisparity:
    calib_params: CamProjCalibrationParams
    cam_proj_maps: CamProjMaps
    proj_time_map_rect: InitVar[np.ndarray]

    proj_x_map: np.ndarray = field(init=False)

    def __post_init__(self, proj_time_map_rect):
        from .engine import build_x_map

        self.X_OFFSET = X_OFFSET
        # int16 capacity checks of the reference (:52-53)
        assert proj_time_map_rect.shape[0] <= 2**15 - 1
        assert proj_time_map_rect.shape[1] + self.X_OFFSET <= 2**15 - 1
        self.X_MAP_WIDTH = self.calib_params.projector_width
        self.T_PX_SCALE = self.X_MAP_WIDTH - 1
        x_map, t_diffs = build_x_map(
            proj_time_map_rect,
            x_map_width=self.X_MAP_WIDTH,
            t_px_scale=self.T_PX_SCALE,
            x_offset=self.X_OFFSET,
            num_scanlines=self.calib_params.projector_width,
        )
        self.proj_x_map = x_map.cpu().numpy()
        self.t_diffs = t_diffs
        self.cam_proj_maps.register_x_map(self.proj_x_map, self.T_PX_SCALE, self.X_OFFSET)

    def compute_event_disparity(self, events, ev_x_rect_i16, ev_y_rect_i16):
        """Reference :69-82 -> ``(disparity[M] int16, inlier_mask[N] bool)`` as lazy device handles:
        nothing runs until a later stage renders the frame (fused kernels) or someone looks at the
        values (stage-by-stage kernels)."""
        from .lazy import DeviceArray, FrameTicket, as_device_events, to_tensor

        ticket = getattr(ev_x_rect_i16, "ticket", None)
        if ticket is None:
            # plain arrays supplied by the caller: honour them in the staged path
            import torch

            maps = self.cam_proj_maps
            dev = as_device_events(events)
            ticket = FrameTicket(maps.engine(dev.device), dev)
            ticket._rect = (
                to_tensor(ev_x_rect_i16, dev.device, torch.int16),
                to_tensor(ev_y_rect_i16, dev.device, torch.int16),
            )
        n = len(ticket.events)
        return (
            DeviceArray(lambda: ticket.disparity()[0], ticket=ticket, role="disparity"),
            DeviceArray(lambda: ticket.disparity()[1], length=n, ticket=ticket, role="inlier_mask"),
        )
