"""The drop-in modules (x-maps_b200/dropin) must offer the reference's call surface.  Runs only where
the reference checkout exists (the authoring container); compares names and signatures, and makes
the reference's UNCHANGED depth_reprojection_pipe.py / _processor.py import against them."""
import importlib
import inspect
import os
import subprocess
import sys

import pytest

from xm_helpers import ROOT

REF = "/root/reference/python"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")

DROPIN = os.path.join(ROOT, "x-maps_b200", "dropin")


MODULES = ("cam_proj_calibration", "x_maps_disparity", "disp_to_depth", "proj_time_map", "x_map", "frame_event_filter", "trigger_finder",
           "metavision_sdk_base", "stats_printer")


def _load(path_first, name):
    """Import `name` with `path_first` in front of sys.path, isolated from earlier imports."""
    saved_path, saved_mods = list(sys.path), dict(sys.modules)
    try:
        for m in MODULES:
            sys.modules.pop(m, None)
        sys.path[:0] = path_first
        return importlib.import_module(name)
    finally:
        sys.path[:] = saved_path
        for m in MODULES:
            sys.modules.pop(m, None)
        sys.modules.update({k: v for k, v in saved_mods.items() if k not in sys.modules})


def params(fn):
    return [p for p in inspect.signature(fn).parameters if p != "self"]


SURFACE = {
    "cam_proj_calibration": {
        "CamProjCalibrationParams": ["from_yaml", "from_ESL_yaml"],
        "CamProjMaps": [
            "rectify_cam_coords_i16", "rectify_cam_coords_f32", "compute_disp_map_projector_view",
            "compute_disp_map_camera_view", "construct_point_cloud",
        ],
    },
    "x_maps_disparity": {"XMapsDisparity": ["compute_event_disparity"]},
    "disp_to_depth": {"DisparityToDepth": ["remap_rectified_disp_map_to_proj", "colorize_depth_from_disp"]},
    "proj_time_map": {"ProjectorTimeMap": ["from_calib", "from_file"]},
    "frame_event_filter": {
        "LastEventPerXYFilter": ["filter_events"], "FirstEventPerXYFilter": ["filter_events"],
        "FirstEventPerYTFilter": ["filter_events"], "MeanFirstLastEventPerXYFilter": ["filter_events"],
        "NoFilter": ["filter_events"], "FrameEventFilterProcessor": ["selected_filter", "filter_events", "select_next_filter"],
    },
    "trigger_finder": {"RobustTriggerFinder": ["reset", "drop_frame", "process_events", "find_trigger"]},
}
STUBS = os.path.join(ROOT, "tests", "stubs")  # Metavision / stats_printer stand-ins the reference's trigger_finder imports


@pytest.mark.parametrize("module", sorted(SURFACE))
def test_methods_have_reference_signatures(module):
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/xmaps_numba_cache")
    ref = _load([STUBS, REF], module)
    ours = _load([DROPIN, ROOT], module)
    assert ours.__file__.startswith(DROPIN)
    for cls, methods in SURFACE[module].items():
        rc, oc = getattr(ref, cls), getattr(ours, cls)
        # dataclass constructor fields (the way the reference's pipe builds the objects)
        ref_init = [p for p in params(rc.__init__)]
        our_init = [p for p in params(oc.__init__)]
        if ref_init != ["args", "kwargs"]:  # (classes without their own __init__ report object's)
            assert our_init[: len(ref_init)] == ref_init, f"{cls} constructor: {our_init} vs {ref_init}"
        for m in methods:
            assert params(getattr(oc, m)) == params(getattr(rc, m)), f"{cls}.{m}"


def test_free_functions():
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/xmaps_numba_cache")
    ours = _load([DROPIN, ROOT], "disp_to_depth")
    assert params(ours.disparity_to_depth_rectified) == ["disparity", "P1"]
    xm = _load([DROPIN, ROOT], "x_map")
    assert params(xm.compute_x_map_from_time_map) == ["time_map", "x_map_width", "t_px_scale", "X_OFFSET", "num_scanlines"]


def test_unchanged_reference_pipe_imports_against_dropin():
    """`import depth_reprojection_processor` (which imports depth_reprojection_pipe) with the
    drop-in directory ahead of the reference's: the pipe must bind OUR classes."""
    code = (
        "import depth_reprojection_processor as P, depth_reprojection_pipe as D, xmaps_b200.calibration as C, "
        "xmaps_b200.disparity as X, xmaps_b200.depth as Z\n"
        "assert D.__file__.startswith('/root/reference'), D.__file__\n"
        "assert D.CamProjMaps is C.CamProjMaps and D.XMapsDisparity is X.XMapsDisparity and D.DisparityToDepth is Z.DisparityToDepth\n"
        "import xmaps_b200.trigger_finder as T, xmaps_b200.frame_event_filter as F\n"
        "assert D.RobustTriggerFinder is T.RobustTriggerFinder and D.FrameEventFilterProcessor is F.FrameEventFilterProcessor\n"
        "assert P.DepthReprojectionPipe is D.DepthReprojectionPipe\n"
        "print('bound')\n"
    )
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([ROOT, DROPIN, os.path.join(ROOT, "tests", "stubs"), REF])
    env["NUMBA_CACHE_DIR"] = "/tmp/xmaps_numba_cache"
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "bound" in out.stdout, out.stderr[-2000:]
