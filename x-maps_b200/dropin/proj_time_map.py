"""Drop-in for the reference's ``proj_time_map`` module (python/proj_time_map.py)."""
from xmaps_b200.time_map import (  # noqa: F401
    ProjectorTimeMap,
    generate_linear_projector_time_map,
    remap_proj_time_map,
)
