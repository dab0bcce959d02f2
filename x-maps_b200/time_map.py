"""Projector time map (set-up time, host side).

Mirrors the reference's ``proj_time_map`` module (/root/reference/python/proj_time_map.py):
a laser projector that scans column by column lights pixel (x, y) at normalised time
``(x * H + y) / (W * H)`` (:6-19); the map is then warped into the rectified frame with a
nearest-neighbour ``cv2.remap`` (:22-29).  A calibrated map can be loaded from ``.npy`` (:46-49).
"""
from __future__ import annotations

from dataclasses import dataclass

import cv2
import numpy as np


def generate_linear_projector_time_map(proj_width: int, proj_height: int, scan_upwards: bool) -> np.ndarray:
    col = np.arange(proj_width, dtype=np.int64)[None, :]
    row = np.arange(proj_height, dtype=np.int64)[:, None]
    if scan_upwards:
        row = row[::-1]  # bottom-to-top scan inside a column
    order = col * proj_height + row
    return (order / (proj_width * proj_height)).astype(np.float32)


def remap_proj_time_map(cam_proj_maps, proj_time_map, border_mode) -> np.ndarray:
    return cv2.remap(
        proj_time_map, cam_proj_maps.projector_mapx, cam_proj_maps.projector_mapy, cv2.INTER_NEAREST, border_mode
    )


@dataclass
class ProjectorTimeMap:
    projector_time_map_rectified: np.ndarray

    @staticmethod
    def from_calib(calib_params, cam_proj_maps, scan_upwards=True, remap_border_mode=cv2.BORDER_REPLICATE):
        linear = generate_linear_projector_time_map(
            calib_params.projector_width, calib_params.projector_height, scan_upwards
        )
        return ProjectorTimeMap(remap_proj_time_map(cam_proj_maps, linear, border_mode=remap_border_mode))

    @staticmethod
    def from_file(proj_time_map_path):
        return ProjectorTimeMap(np.load(proj_time_map_path))
