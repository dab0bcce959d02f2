"""The reference's own per-frame chain, run on the unmodified modules under ``oracle/_ref/``.

TEST INFRASTRUCTURE ONLY: used by ``tests/`` (second checker), by ``bench.py``'s ``cpu_baseline`` leg and by
``bench.py --impl reference`` (the thing timed there IS the reference).  Nothing under ``x-maps_b200/``
imports this.  ``oracle/_ref/`` is produced by ``oracle/build_ref.py`` (copies of the five path modules and
the calibration YAML, git-ignored); where it is absent ``available()`` is False and callers fall back to the
NumPy restatement (``xmaps_oracle``), saying so.

``RefPath.frame_depth`` is the call sequence of ``DepthReprojectionPipe.process_ev_frame``
(python/depth_reprojection_pipe.py:121-162) up to the depth frame, every call going into the reference's
own functions: ``rectify_cam_coords_i16`` -> ``compute_event_disparity`` -> ``compute_disp_map_projector_view``
| ``compute_disp_map_camera_view`` -> ``remap_rectified_disp_map_to_proj`` -> ``disparity_to_depth_rectified``;
the polarity filter (closed Metavision binary) is restated as ``p == 1`` the way the reference restates it
(python/frame_event_filter.py:21).
"""
from __future__ import annotations

import importlib
import os
import sys
from types import SimpleNamespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
MODULES = ["x_map", "cam_proj_calibration", "proj_time_map", "x_maps_disparity", "disp_to_depth"]
CALIB_YAML = os.path.join(REF_DIR, "ESL_calib_hhi.yaml")

_loaded = None


def available() -> bool:
    return all(os.path.exists(os.path.join(REF_DIR, m + ".py")) for m in MODULES) and os.path.exists(CALIB_YAML)


def load():
    """Import the reference modules from oracle/_ref WITHOUT leaving them in ``sys.modules`` / ``sys.path``
    (the product's drop-in shims use the same module names)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise ImportError("oracle/_ref is not built (python oracle/build_ref.py needs /root/reference)")
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/xmaps_numba_cache")
    saved = {n: sys.modules.pop(n, None) for n in MODULES}
    sys.path.insert(0, REF_DIR)
    try:
        mods = {n: importlib.import_module(n) for n in MODULES}
    finally:
        sys.path.remove(REF_DIR)
        for n in MODULES:
            sys.modules.pop(n, None)
            if saved[n] is not None:
                sys.modules[n] = saved[n]
    # The modules are not importable by name afterwards, which Numba's on-disk cache (cache=True in the
    # reference) needs when it reloads an entry: compile in-process instead (1-2 s, once per process).
    try:
        from numba.core.caching import NullCache
        from numba.core.dispatcher import Dispatcher

        for mod in mods.values():
            for obj in vars(mod).values():
                if isinstance(obj, Dispatcher):
                    obj._cache = NullCache()
    except ImportError:  # a Numba without these internals: keep its default behaviour
        pass
    _loaded = SimpleNamespace(**mods)
    return _loaded


class _NullStats:
    """``stats.measure_time(key)`` context manager the reference's DisparityToDepth expects
    (python/disp_to_depth.py:85,88)."""

    class _T:
        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

    def measure_time(self, key):
        return self._T()


class RefPath:
    """Reference objects of one geometry, built by the reference's own constructors from its own YAML
    (python/depth_reprojection_pipe.py:64-99).  ``x_map``: skip the X-map construction (1-6 s of Numba,
    set-up only) and adopt a table that the same constructor produced earlier (the golden fixture)."""

    def __init__(self, cam_w=640, cam_h=480, proj_w=720, proj_h=1280, x_map=None, camera_K_scale=None, cy_shift=0.0, proj_K_scale=None):
        m = load()
        self.m = m
        p = m.cam_proj_calibration.CamProjCalibrationParams.from_yaml(CALIB_YAML, cam_w, cam_h, proj_w, proj_h)
        if camera_K_scale is not None:  # intrinsics scaled
            k = p.camera_K.copy()
            k[:2, :] *= camera_K_scale
            k[1, 2] += cy_shift
            p.camera_K = k
        if proj_K_scale is not None:
            kp = p.projector_K.copy()
            kp[:2, :] *= proj_K_scale
            p.projector_K = kp
        self.params = p
        self.maps = m.cam_proj_calibration.CamProjMaps(p)
        self.time_map = m.proj_time_map.ProjectorTimeMap.from_calib(p, self.maps)
        if x_map is None:
            self.xd = m.x_maps_disparity.XMapsDisparity(
                calib_params=p, cam_proj_maps=self.maps, proj_time_map_rect=self.time_map.projector_time_map_rectified
            )
        else:
            xd = object.__new__(m.x_maps_disparity.XMapsDisparity)  # constructor bypassed: its only job is the table
            xd.calib_params, xd.cam_proj_maps = p, self.maps
            xd.X_OFFSET = 4242
            xd.X_MAP_WIDTH = p.projector_width
            xd.T_PX_SCALE = xd.X_MAP_WIDTH - 1
            xd.proj_x_map = np.ascontiguousarray(x_map, dtype=np.int16)
            self.xd = xd
        self.d2d = m.disp_to_depth.DisparityToDepth(stats=_NullStats(), calib_params=p, calib_maps=self.maps, z_near=0.1, z_far=1.0)

    def frame_disparity_map(self, events, view: int, apply_polarity: bool = True):
        evs = events[events["p"] == 1] if apply_polarity else events
        xr, yr = self.maps.rectify_cam_coords_i16(evs)
        disp, mask = self.xd.compute_event_disparity(events=evs, ev_x_rect_i16=xr, ev_y_rect_i16=yr)
        if view == 1:
            return self.maps.compute_disp_map_camera_view(events=evs, inlier_mask=mask, ev_disparity_f32=disp)
        rect = self.maps.compute_disp_map_projector_view(ev_x_rect_i16=xr, ev_y_rect_i16=yr, inlier_mask=mask, ev_disparity_f32=disp)
        return self.d2d.remap_rectified_disp_map_to_proj(rect)

    def frame_depth(self, events, view: int = 0, apply_polarity: bool = True):
        return self.m.disp_to_depth.disparity_to_depth_rectified(self.frame_disparity_map(events, view, apply_polarity), self.maps.P2)

    def tables_match(self, tables) -> bool:
        """The reference objects built here carry the same tables as an ``OracleTables`` fixture."""
        return bool(
            np.array_equal(self.maps.disp_cam_mapx_i16, tables.lut_x)
            and np.array_equal(self.maps.disp_cam_mapy_i16, tables.lut_y)
            and np.array_equal(self.maps.disp_proj_mapxy_i16, tables.remap_xy)
            and np.array_equal(self.xd.proj_x_map, tables.x_map)
            and float(self.maps.P2[0, 3]) == float(tables.depth_scale)
        )
