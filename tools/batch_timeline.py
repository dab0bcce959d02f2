#!/usr/bin/env python
"""Per-frame hand-off timeline of one batch_kernel launch (XM_DEBUG_HOOKS build, option debug = 8):
first consumer enters frame / last chunk count published / first tile group sees the frame complete / last tile done.
    XMAPS_B200_LIB=build_variants/libxm_hooks.so python tools/batch_timeline.py --events 1000000 [--frames 16]"""
import argparse, ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(1, os.path.join(ROOT, "tests"))
import numpy as np, torch
import bench
from xmaps_b200.engine import DepthEngine, TableSet, OUT_DEPTH
ap = argparse.ArgumentParser()
ap.add_argument("--events", type=int, default=1_000_000)
ap.add_argument("--frames", type=int, default=16)
ap.add_argument("--opt", action="append", default=[])
a = ap.parse_args()
dev = torch.device("cuda", 0)
t = bench.load_geometry("5m")[0]
eng = DepthEngine(TableSet(t.lut_x, t.lut_y, t.x_map, t.remap_xy, t.rect_w, t.rect_h, t.t_px_scale, t.x_offset, t.depth_scale), device=dev)
for kv in a.opt:
    k, v = kv.split("="); eng.set_option(k, int(v))
dbg = 8
for kv in a.opt:
    if kv.startswith("debug="): dbg |= int(kv.split("=")[1])
eng.set_option("debug", dbg)
frames = [bench.synth_frame_cuda(i, a.events, dev, 640, 480) for i in range(a.frames)]
for _ in range(3):
    eng.frame_batch(frames, output=OUT_DEPTH)
torch.cuda.synchronize()
ptr = eng.get_option("debug_ptr")
buf = torch.empty(32 * 4 + 256, dtype=torch.int64, device=dev)
ctypes.CDLL("libcudart.so").cudaMemcpy(ctypes.c_void_p(buf.data_ptr()), ctypes.c_void_p(ptr), ctypes.c_size_t((32 * 4 + 256) * 8), 3)
raw = buf.cpu().numpy()
x = raw[:128].reshape(32, 4)[: a.frames].astype(np.float64)
t0 = x[:, 0].min()
r = (x - t0) / 1e3
print("frame  enter  events-done  tiles-start  tiles-done   (us)   | events  publish->seen  tiles   frame-to-frame")
for f in range(a.frames):
    d = r[f]
    ff = r[f, 3] - r[f - 1, 3] if f else 0.0
    print(f"{f:5d} {d[0]:7.2f} {d[1]:10.2f} {d[2]:11.2f} {d[3]:11.2f}          | {d[1]-d[0]:6.2f} {d[2]-d[1]:10.2f} {d[3]-d[2]:9.2f} {ff:10.2f}")
print("total %.2f us = %.2f us per frame" % (r[:, 3].max(), r[:, 3].max() / a.frames))
acc = raw[256:266].astype(np.float64)
if acc[8] + acc[9] > 0:  # strip epilogue: lane 0's cycles per phase (launches of the warm-up included: only the ratios and per-item means matter)
    for ps, name in ((0, "pass 1"), (1, "pass 2")):
        n = max(acc[8 + ps], 1.0)
        print("%s: %d items, per item (us at 1.965 GHz): ticket %.2f  wait %.2f  work %.2f  publish %.2f" % (
            name, acc[8 + ps], *(acc[ps * 4 + k] / n / 1965.0 for k in range(4))))
