"""Per-frame driver: the body of the reference's ``DepthReprojectionPipe.process_ev_frame``
(/root/reference/python/depth_reprojection_pipe.py:121-167) and its set-up (:64-108), without the
camera SDK around it.

``DepthFramePipeline`` owns the calibration objects exactly as the reference pipe does
(``calib_maps``, ``x_maps_disp``, ``disp_to_depth``) and renders one frame of events per call.
The stages are invoked in the reference's order through the reference's call surface, so this is
also the executable specification of how the unchanged reference pipe drives the drop-in modules.
"""
from __future__ import annotations

from contextlib import nullcontext
from dataclasses import dataclass
from typing import Callable, Optional

from .calibration import CamProjCalibrationParams, CamProjMaps
from .depth import DisparityToDepth
from .disparity import XMapsDisparity
from .frame_event_filter import FrameEventFilterProcessor
from .time_map import ProjectorTimeMap
from .trigger_finder import RobustTriggerFinder


@dataclass
class RuntimeParams:
    """Same fields as the reference's ``RuntimeParams`` (python/depth_reprojection_processor.py:13-36)."""

    camera_width: int
    camera_height: int
    projector_width: int
    projector_height: int
    projector_fps: int
    z_near: float
    z_far: float
    calib: str
    projector_time_map: Optional[str]
    no_frame_dropping: bool
    camera_perspective: bool

    @property
    def should_drop_frames(self):
        return not self.no_frame_dropping


class _NullStats:
    def measure_time(self, key):
        return nullcontext()

    def add_metric(self, *a, **k):
        pass

    def count(self, *a, **k):
        pass


class DepthFramePipeline:
    def __init__(self, params: RuntimeParams, stats_printer=None, frame_callback: Optional[Callable] = None, return_host: bool = True,
                 activity_filter: bool = True):
        self.params = params
        self.stats_printer = stats_printer if stats_printer is not None else _NullStats()
        self.frame_callback = frame_callback
        calib_params = CamProjCalibrationParams.from_yaml(
            params.calib, params.camera_width, params.camera_height, params.projector_width, params.projector_height
        )
        self.calib_maps = CamProjMaps(calib_params)
        if params.projector_time_map is not None:
            proj_time_map = ProjectorTimeMap.from_file(params.projector_time_map)
        else:
            proj_time_map = ProjectorTimeMap.from_calib(calib_params, self.calib_maps)
        self.x_maps_disp = XMapsDisparity(
            calib_params=calib_params,
            cam_proj_maps=self.calib_maps,
            proj_time_map_rect=proj_time_map.projector_time_map_rectified,
        )
        self.disp_to_depth = DisparityToDepth(
            stats=self.stats_printer, calib_params=calib_params, calib_maps=self.calib_maps,
            z_near=params.z_near, z_far=params.z_far, return_host=return_host,
        )
        # the rows either side of the path (depth_reprojection_pipe.py:97-105): per-frame de-duplication
        # filter (NoFilter until select_next_frame_event_filter) and the frame segmentation of the stream
        eng = self.calib_maps.engine()
        # ActivityNoiseFilterAlgorithm(width, height, int(1e6 / projector_fps)) (depth_reprojection_pipe.py:65-67)
        self.activity_filter = activity_filter
        self.activity_threshold_us = int(1e6 / params.projector_fps)
        eng.activity_reset()
        self.ev_filter_proc = FrameEventFilterProcessor(engine=eng)
        self.trigger_finder = RobustTriggerFinder(
            projector_fps=params.projector_fps, stats=self.stats_printer, pool=None, frame_callback=self.process_ev_frame, engine=eng,
        )

    def process_events(self, evs):
        """A slice of the continuous stream (depth_reprojection_pipe.py:108-119): polarity filter,
        activity-noise filter (:116-117; the Metavision binary's semantics restated, see
        ``DepthEngine.activity_filter``), then the trigger finder, which calls process_ev_frame once per
        projector frame it finds."""
        eng = self.calib_maps.engine()
        pos = eng.polarity_filter(evs)
        if self.activity_filter:
            pos = eng.activity_filter(pos, self.activity_threshold_us)
        self.trigger_finder.process_events(pos)

    def select_next_frame_event_filter(self):
        return self.ev_filter_proc.select_next_filter()

    def reset(self):
        self.trigger_finder.reset()
        self.calib_maps.engine().activity_reset()

    def process_ev_frame(self, evs):
        """One frame of (already polarity-filtered) events -> colourised depth frame; the call
        sequence of depth_reprojection_pipe.py:121-167."""
        sp = self.stats_printer
        with sp.measure_time("ev rect"):
            ev_x_rect_i16, ev_y_rect_i16 = self.calib_maps.rectify_cam_coords_i16(evs)
        with sp.measure_time("frame ev filter"):
            filtered_evs = self.ev_filter_proc.filter_events(evs, ev_x_rect_i16)
            if len(evs):
                sp.add_metric("frame evs filtered out [%]", 100 - len(filtered_evs) / len(evs) * 100)
            if len(filtered_evs) < len(evs):  # :137-139: the survivors are rectified again
                ev_x_rect_i16, ev_y_rect_i16 = self.calib_maps.rectify_cam_coords_i16(filtered_evs)
            evs = filtered_evs
        with sp.measure_time("x-maps disp"):
            ev_disparity_f32, inlier_mask = self.x_maps_disp.compute_event_disparity(
                events=evs, ev_x_rect_i16=ev_x_rect_i16, ev_y_rect_i16=ev_y_rect_i16
            )
        if self.params.camera_perspective:
            with sp.measure_time("disp map"):
                disp_map = self.calib_maps.compute_disp_map_camera_view(
                    events=evs, inlier_mask=inlier_mask, ev_disparity_f32=ev_disparity_f32
                )
        else:
            with sp.measure_time("disp map"):
                disp_map = self.calib_maps.compute_disp_map_projector_view(
                    ev_x_rect_i16=ev_x_rect_i16, ev_y_rect_i16=ev_y_rect_i16,
                    inlier_mask=inlier_mask, ev_disparity_f32=ev_disparity_f32,
                )
            with sp.measure_time("remap disp"):
                disp_map = self.disp_to_depth.remap_rectified_disp_map_to_proj(disp_map)
        with sp.measure_time("disp2rgb"):
            depth_map = self.disp_to_depth.colorize_depth_from_disp(disp_map)
        if self.frame_callback is not None:
            self.frame_callback(depth_map)
        return depth_map

    def depth_frame(self, evs, polarity: bool = True):
        """Metric depth frame (float32 CUDA tensor) of one raw frame of events, polarity mask fused."""
        from .engine import OUT_DEPTH, VIEW_CAMERA, VIEW_PROJECTOR

        eng = self.calib_maps.engine()
        view = VIEW_CAMERA if self.params.camera_perspective else VIEW_PROJECTOR
        return eng.frame(evs, view=view, output=OUT_DEPTH, polarity=polarity)
