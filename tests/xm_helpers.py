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
