"""Helper for generate_golden_stream.py: the small geometry's rectify LUT from the committed tables."""
import os

import numpy as np


def load_small_lut():
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "tables_small.npz"))
    return z["lut_x"]
