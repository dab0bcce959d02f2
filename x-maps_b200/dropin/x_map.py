"""Drop-in for the reference's ``x_map`` module (python/x_map.py)."""


def compute_x_map_from_time_map(time_map, x_map_width, t_px_scale, X_OFFSET, num_scanlines):
    """Same signature and return value (host arrays) as python/x_map.py:5-55, computed on the GPU."""
    from xmaps_b200.engine import build_x_map

    x_map, t_diffs = build_x_map(time_map, x_map_width, t_px_scale, X_OFFSET, num_scanlines)
    return x_map.cpu().numpy(), t_diffs.cpu().numpy()
