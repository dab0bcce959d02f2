"""The host-side table builder (x-maps_b200/calibration.py, time_map.py) must reproduce the
reference's tables bit-for-bit: same OpenCV calls, same arguments.  CPU only."""
import hashlib
import os

import numpy as np
import pytest

from xm_helpers import ROOT, load_golden_tables

CALIB = os.path.join(ROOT, "data", "esl_calib_hhi.json")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def scaled(params, cam_scale, proj_scale, cy_shift=0.0):
    k = params.camera_K.copy()
    k[:2, :] *= cam_scale
    k[1, 2] += cy_shift
    params.camera_K = k
    kp = params.projector_K.copy()
    kp[:2, :] *= proj_scale
    params.projector_K = kp
    return params


def test_default_tables_match_reference(manifest):
    from xmaps_b200.calibration import CamProjCalibrationParams, CamProjMaps
    from xmaps_b200.time_map import ProjectorTimeMap

    h = manifest["configs"]["default"]["hash"]
    p = CamProjCalibrationParams.from_yaml(CALIB, 640, 480, 720, 1280)
    assert (p.rect_image_width, p.rect_image_height) == (1760, 1320)
    maps = CamProjMaps(p)
    assert sha(maps.disp_cam_mapx_i16) == h["lut_x"]
    assert sha(maps.disp_cam_mapy_i16) == h["lut_y"]
    assert sha(maps.disp_cam_mapx_f32) == h["lut_x_f32"]
    assert sha(maps.disp_cam_mapy_f32) == h["lut_y_f32"]
    assert sha(maps.disp_proj_mapxy_i16) == h["remap_xy"]
    assert float(maps.P2[0, 3]) == manifest["configs"]["default"]["P2_03"]
    tm = ProjectorTimeMap.from_calib(p, maps)
    assert sha(tm.projector_time_map_rectified) == h["time_map_rect"]


def test_small_tables_match_reference(manifest):
    from xmaps_b200.calibration import CamProjCalibrationParams, CamProjMaps
    from xmaps_b200.time_map import ProjectorTimeMap

    tables, z = load_golden_tables("small")
    p = scaled(CamProjCalibrationParams.from_yaml(CALIB, 160, 120, 180, 320), 0.25, 0.25)
    maps = CamProjMaps(p)
    assert np.array_equal(maps.disp_cam_mapx_i16, tables.lut_x)
    assert np.array_equal(maps.disp_cam_mapy_i16, tables.lut_y)
    assert np.array_equal(maps.disp_proj_mapxy_i16, tables.remap_xy)
    assert np.array_equal(maps.Q, z["Q"])
    tm = ProjectorTimeMap.from_calib(p, maps)
    assert np.array_equal(tm.projector_time_map_rectified, z["time_map_rect"])


def test_hd_tables_match_reference(manifest):
    from xmaps_b200.calibration import CamProjCalibrationParams, CamProjMaps
    from xmaps_b200.time_map import ProjectorTimeMap

    cfg = manifest["configs"]["hd"]
    p = scaled(CamProjCalibrationParams.from_yaml(CALIB, 1280, 720, 1080, 1920), 2.0, 1.0, -120.0)
    assert [p.rect_image_width, p.rect_image_height] == cfg["rect_wh"]
    maps = CamProjMaps(p)
    assert sha(maps.disp_cam_mapx_i16) == cfg["hash"]["lut_x"]
    assert sha(maps.disp_cam_mapy_i16) == cfg["hash"]["lut_y"]
    assert sha(maps.disp_proj_mapxy_i16) == cfg["hash"]["remap_xy"]
    assert float(maps.P2[0, 3]) == cfg["P2_03"]
    tm = ProjectorTimeMap.from_calib(p, maps)
    assert sha(tm.projector_time_map_rectified) == cfg["hash"]["time_map_rect"]


def test_missing_matrix_raises(tmp_path):
    from xmaps_b200.calibration import CamProjCalibrationParams

    bad = tmp_path / "bad.json"
    bad.write_text('{"matrices": {}}')
    with pytest.raises(ValueError):
        CamProjCalibrationParams.from_yaml(str(bad), 640, 480, 720, 1280)
