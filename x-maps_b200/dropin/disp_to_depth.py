"""Drop-in for the reference's ``disp_to_depth`` module (python/disp_to_depth.py)."""
from xmaps_b200.depth import DisparityToDepth, disparity_to_depth_rectified  # noqa: F401
