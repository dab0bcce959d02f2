#!/usr/bin/env python
"""Tiny reproducer for the device-side fix-up path (unsorted frame with optimistic bounds)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(1, os.path.join(ROOT, "tests"))
import numpy as np, torch
from oracle import xmaps_oracle as orc
from xm_helpers import load_golden_tables, golden_frame
from xmaps_b200.engine import DepthEngine, TableSet, TBOUNDS_SORTED, TBOUNDS_REDUCE
t, _ = load_golden_tables("small")
eng = DepthEngine(TableSet(t.lut_x, t.lut_y, t.x_map, t.remap_xy, t.rect_w, t.rect_h, t.t_px_scale, t.x_offset, t.depth_scale), device="cuda:0")
mode = TBOUNDS_SORTED
for kv in sys.argv[1:]:
    k, v = kv.split("=")
    if k == "mode":
        mode = int(v)
    else:
        eng.set_option(k, int(v))
ev = orc.synth_events(2, 20_000, 160, 120)
print("sorted frame ok:", np.array_equal(eng.frame(ev, view=0, time_bounds=mode).cpu().numpy(), golden_frame("small_20k_proj")["depth"]), flush=True)
np.random.default_rng(7).shuffle(ev)
g = golden_frame("small_20k_shuffled_proj")
for i in range(3):
    d = eng.frame(ev, view=0, time_bounds=mode).cpu().numpy()
    print("unsorted frame", i, "ok:", np.array_equal(d, g["depth"]), eng.status(), flush=True)
