"""Drop-in for the reference's ``cam_proj_calibration`` module (python/cam_proj_calibration.py)."""
from xmaps_b200.calibration import (  # noqa: F401
    CamProjCalibrationParams,
    CamProjMaps,
    inverse_rectify_map as initUndistortRectifyMapInverse,
    round_map_to_i16 as mapf_to_i16,
)
