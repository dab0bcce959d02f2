#!/usr/bin/env python
"""Small driver for ncu: a few 5 M-event frames of the bench workload through the fused path.

    ncu --set full --clock-control none --import-source on -k regex:events_kernel -s 8 -c 3 \
        -o gpurun_out/prof_k1 python tools/profile_frames.py [--frames 4] [--reps 4] [--view 0]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(1, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import bench  # noqa: E402
from xmaps_b200.engine import OUT_DEPTH, DepthEngine, TableSet  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=4)
    ap.add_argument("--reps", type=int, default=4)
    ap.add_argument("--events", type=int, default=bench.WORKLOADS["5m"][5])
    ap.add_argument("--view", type=int, default=0)
    ap.add_argument("--opt", action="append", default=[])
    ap.add_argument("--single", action="store_true", help="one xm_frame call per frame instead of one batch")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    t = bench.load_geometry("5m")[0]
    eng = DepthEngine(TableSet(t.lut_x, t.lut_y, t.x_map, t.remap_xy, t.rect_w, t.rect_h, t.t_px_scale, t.x_offset, t.depth_scale), device=dev)
    for kv in a.opt:
        k, v = kv.split("=")
        eng.set_option(k, int(v))
    frames = [bench.synth_frame_cuda(i, a.events, dev, 640, 480) for i in range(a.frames)]
    torch.cuda.synchronize()
    for _ in range(a.reps):
        if a.single:
            for f in frames:
                eng.frame(f, view=a.view, output=OUT_DEPTH)
        else:
            eng.frame_batch(frames, view=a.view, output=OUT_DEPTH)
    torch.cuda.synchronize()
    print("done", eng.status())


if __name__ == "__main__":
    main()
