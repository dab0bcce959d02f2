"""Drop-in for the reference's ``x_maps_disparity`` module (python/x_maps_disparity.py)."""
from xmaps_b200.disparity import XMapsDisparity, X_OFFSET  # noqa: F401
