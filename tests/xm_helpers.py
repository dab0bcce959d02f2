"""Shared helpers for the test-suite (golden loading)."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_golden_tables(name):
    """OracleTables from tests/golden/tables_<name>.npz (written by generate_golden.py)."""
    from oracle.xmaps_oracle import OracleTables

    z = np.load(os.path.join(GOLDEN, f"tables_{name}.npz"))
    tables = OracleTables(
        lut_x=z["lut_x"],
        lut_y=z["lut_y"],
        x_map=z["x_map"],
        remap_xy=np.ascontiguousarray(np.stack((z["remap_x"], z["remap_y"]), axis=-1)),
        rect_w=int(z["rect_wh"][0]),
        rect_h=int(z["rect_wh"][1]),
        t_px_scale=int(z["consts"][0]),
        x_offset=int(z["consts"][1]),
        depth_scale=float(z["depth_scale"][0]),
    )
    return tables, z


def golden_frame(name):
    return np.load(os.path.join(GOLDEN, f"frame_{name}.npz"))


def host_time_map_rect(cam_w=640, cam_h=480, proj_w=720, proj_h=1280, camera_K_scale=None, cy_shift=0.0):
    """Rectified projector time map of a geometry from the product's HOST table builder (same OpenCV calls as
    the reference; tests/test_host_tables.py pins it to the reference's hashes).  Input of the plane workload."""
    from xmaps_b200.calibration import CamProjCalibrationParams, CamProjMaps
    from xmaps_b200.time_map import ProjectorTimeMap

    p = CamProjCalibrationParams.from_yaml(os.path.join(ROOT, "data", "esl_calib_hhi.json"), cam_w, cam_h, proj_w, proj_h)
    if camera_K_scale is not None:
        k = p.camera_K.copy()
        k[:2, :] *= camera_K_scale
        k[1, 2] += cy_shift
        p.camera_K = k
    maps = CamProjMaps(p)
    return ProjectorTimeMap.from_calib(p, maps).projector_time_map_rectified
