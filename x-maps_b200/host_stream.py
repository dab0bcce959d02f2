"""Host-buffer front end: frames of events in pinned host memory -> depth frames in pinned host
memory, with the PCIe copies of neighbouring frames overlapped with the kernels.

This is the call a user of the reference makes (events arrive in host memory from the camera SDK,
/root/reference/python/depth_reprojection.py:10-29): three CUDA streams (H2D, compute, D2H) and a
ring of device staging buffers; nothing is computed on the host.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch

from .engine import OUT_BGR, OUT_DEPTH, TBOUNDS_SORTED, VIEW_PROJECTOR, DepthEngine
from .events import DeviceEvents


def pin_events(events: np.ndarray) -> torch.Tensor:
    """Copy a host EventCD array into pinned memory as an int32 [N, 4] tensor."""
    arr = np.ascontiguousarray(events)
    t = torch.from_numpy(arr.view(np.int32).reshape(-1, 4))
    return t.pin_memory()


class HostFrameStream:
    def __init__(self, engine: DepthEngine, max_events: int, view: int = VIEW_PROJECTOR, output: int = OUT_DEPTH, depth: int = 3):
        self.engine = engine
        self.view, self.output = view, output
        dev = engine.device
        self.h2d = torch.cuda.Stream(dev)
        self.compute = torch.cuda.Stream(dev)
        self.d2h = torch.cuda.Stream(dev)
        self.depth = depth
        self.ev_bufs = [torch.empty((max_events, 4), dtype=torch.int32, device=dev) for _ in range(depth)]
        shape = engine.out_shape(view, output)
        dtype = torch.uint8 if output == OUT_BGR else torch.float32
        self.out_bufs = [torch.empty(shape, dtype=dtype, device=dev) for _ in range(depth)]
        self.copied = [torch.cuda.Event() for _ in range(depth)]
        self.computed = [torch.cuda.Event() for _ in range(depth)]
        self.drained = [torch.cuda.Event() for _ in range(depth)]
        self.out_shape, self.out_dtype = shape, dtype

    def alloc_outputs(self, n_frames: int) -> torch.Tensor:
        return torch.empty((n_frames,) + tuple(self.out_shape), dtype=self.out_dtype).pin_memory()

    def run(self, host_frames: Sequence[torch.Tensor], host_out: torch.Tensor, polarity: bool = True,
            time_bounds: int = TBOUNDS_SORTED, z_near: float = 0.1, z_far: float = 1.0) -> torch.Tensor:
        """``host_frames``: pinned int32 [N_i, 4] tensors; ``host_out``: pinned [n_frames, ...].
        Returns ``host_out`` after everything has landed (synchronises at the end)."""
        eng = self.engine
        for i, hf in enumerate(host_frames):
            slot = i % self.depth
            n = hf.shape[0]
            with torch.cuda.stream(self.h2d):
                if i >= self.depth:
                    self.h2d.wait_event(self.computed[slot])  # staging buffer free again
                dev_ev = self.ev_bufs[slot][:n]
                dev_ev.copy_(hf, non_blocking=True)
                self.copied[slot].record(self.h2d)
            with torch.cuda.stream(self.compute):
                self.compute.wait_event(self.copied[slot])
                if i >= self.depth:
                    self.compute.wait_event(self.drained[slot])  # output buffer free again
                eng.frame(DeviceEvents(dev_ev), view=self.view, output=self.output, polarity=polarity,
                          time_bounds=time_bounds, z_near=z_near, z_far=z_far, out=self.out_bufs[slot])
                self.computed[slot].record(self.compute)
            with torch.cuda.stream(self.d2h):
                self.d2h.wait_event(self.computed[slot])
                host_out[i].copy_(self.out_bufs[slot], non_blocking=True)
                self.drained[slot].record(self.d2h)
        self.d2h.synchronize()
        self.compute.synchronize()
        return host_out
