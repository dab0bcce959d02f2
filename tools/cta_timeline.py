#!/usr/bin/env python
"""Per-CTA phase timeline of K1 (lean variant): where does a CTA spend its time?"""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(1, os.path.join(ROOT, "tests"))
import numpy as np, torch
import bench
from xmaps_b200.engine import DepthEngine, TableSet, OUT_DEPTH
dev = torch.device("cuda", 0)
t = bench.load_geometry("5m")[0]
eng = DepthEngine(TableSet(t.lut_x, t.lut_y, t.x_map, t.remap_xy, t.rect_w, t.rect_h, t.t_px_scale, t.x_offset, t.depth_scale), device=dev)
for kv in sys.argv[1:]:
    k, v = kv.split("="); eng.set_option(k, int(v))
eng.set_option("debug", 8)
frames = [bench.synth_frame_cuda(i, 5_000_000, dev) for i in range(4)]
for _ in range(3):
    eng.frame_batch(frames, output=OUT_DEPTH)
torch.cuda.synchronize()
ptr = eng.get_option("debug_ptr")
buf = torch.empty(4096 * 8, dtype=torch.int64, device=dev)
cudart = ctypes.CDLL("libcudart.so")
cudart.cudaMemcpy(ctypes.c_void_p(buf.data_ptr()), ctypes.c_void_p(ptr), ctypes.c_size_t(4096 * 8 * 8), 3)
a = buf.cpu().numpy().reshape(4096, 8)
a = a[a[:, 0] > 0]
t0 = a[:, 0].min()
rel = (a[:, :7] - t0) / 1e3
names = ["entry", "prologue done", "front(0) done", "loop c=1", "kernel end (fused)", "phase 1 end", "after grid barrier"]
print("CTAs:", len(a), " SMs:", len(np.unique(a[:, 7])))
for i, n in enumerate(names):
    col = rel[:, i]
    print(f"{n:24s} min {col.min():7.2f}  p50 {np.median(col):7.2f}  max {col.max():7.2f} us")
print("phase 2 per CTA (barrier -> end): p50 %.2f  max %.2f us" % (np.median(rel[:, 4] - rel[:, 6]), (rel[:, 4] - rel[:, 6]).max()))
print("raw row 0:", a[0].tolist())
print("raw row 5:", a[5].tolist())
