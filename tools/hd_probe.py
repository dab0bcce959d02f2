#!/usr/bin/env python
"""BASELINE config 3 (HD: cam 1280x720, proj 1080x1920, 20 M events / frame) through the batch kernel:
frame time for a few tile-region capacities.  python tools/hd_probe.py [--events N] [--frames F]"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(1, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from xmaps_b200.calibration import CamProjCalibrationParams, CamProjMaps  # noqa: E402
from xmaps_b200.disparity import XMapsDisparity  # noqa: E402
from xmaps_b200.time_map import ProjectorTimeMap  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--events", type=int, default=20_000_000)
    ap.add_argument("--frames", type=int, default=8)
    a = ap.parse_args()
    p = CamProjCalibrationParams.from_yaml(os.path.join(ROOT, "data", "esl_calib_hhi.json"), 1280, 720, 1080, 1920)
    k = p.camera_K.copy()
    k[:2, :] *= 2.0
    k[1, 2] += -120.0
    p.camera_K = k
    maps = CamProjMaps(p)
    tm = ProjectorTimeMap.from_calib(p, maps)
    XMapsDisparity(calib_params=p, cam_proj_maps=maps, proj_time_map_rect=tm.projector_time_map_rectified)
    eng = maps.engine("cuda:0")
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    frames = []
    for f in range(a.frames):
        g.manual_seed(f)
        n = a.events
        x = torch.randint(0, 1280, (n,), generator=g, device=dev, dtype=torch.int32)
        y = torch.randint(0, 720, (n,), generator=g, device=dev, dtype=torch.int32)
        t = torch.sort(torch.randint(0, 16666, (n,), generator=g, device=dev, dtype=torch.int64)).values
        raw = torch.empty((n, 4), dtype=torch.int32, device=dev)
        raw[:, 0] = x | (y << 16)
        raw[:, 1] = (torch.rand(n, generator=g, device=dev) < 0.9).to(torch.int32)
        raw.view(torch.int64)[:, 1] = t
        frames.append(raw)
    # the opt-in bilinear X-map lookup (XM_FLAG_BILINEAR; staged kernels, one frame per call) next to the nearest path
    for mode in (True, False):
        out1 = eng.frame(frames[0], view=0, bilinear=mode)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            for fr in frames:
                eng.frame(fr, view=0, bilinear=mode, out=out1)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3 / a.frames
        print(f"single frames, {'bilinear' if mode else 'nearest '} lookup: {dt * 1e6:8.1f} us / frame  {a.events / dt / 1e9:6.1f} G ev/s")
    for cells in (3072, 4608, 6144):
        for batch in (1, 0):
            eng.set_option("region_cells", cells)
            eng.set_option("batch", batch)
            out = eng.frame_batch(frames, view=0)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                eng.frame_batch(frames, view=0, out=out)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / 3 / a.frames
            print(f"region_cells={cells} batch={batch}: {dt * 1e6:8.1f} us / frame  {a.events / dt / 1e9:6.1f} G ev/s  batch_smem={eng.get_option('batch_smem')} occ={eng.get_option('batch_occ')}")


if __name__ == "__main__":
    main()
