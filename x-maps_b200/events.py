"""Event buffers: the Metavision ``EventCD`` record and its device-resident form.

The reference passes events around as NumPy structured arrays of ``EventCD``
(/root/reference/python/trigger_finder.py:2,21; fields x:u16 y:u16 p:i16 t:i64, 16 bytes) or, in
the evaluation script, as a plain dict of arrays with a float ``t``
(/root/reference/python/eval/compute_depth_x_maps.py:83-96).  Here a frame of events lives on the
GPU as ONE contiguous buffer of 16-byte records (a torch CUDA tensor); ``DeviceEvents`` wraps it and
still answers ``events["x"]``-style indexing for code written against the reference.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

EVENT_RECORD_BYTES = 16

# Metavision EventCD: 16-byte array-of-structs, t in microseconds
EVENT_DTYPE = np.dtype(
    {
        "names": ["x", "y", "p", "t"],
        "formats": ["<u2", "<u2", "<i2", "<i8"],
        "offsets": [0, 2, 4, 8],
        "itemsize": EVENT_RECORD_BYTES,
    }
)
# same record with a float64 time field (evaluation path)
EVENT_DTYPE_F64 = np.dtype(
    {
        "names": ["x", "y", "p", "t"],
        "formats": ["<u2", "<u2", "<i2", "<f8"],
        "offsets": [0, 2, 4, 8],
        "itemsize": EVENT_RECORD_BYTES,
    }
)


def pack_events(x, y, t, p=None) -> np.ndarray:
    """Build a host EventCD array from columns.  A floating ``t`` yields the float64-time record.

    Note on float32 timestamps: the reference's ``compute_disparity`` normalises time in ``t``'s own dtype
    (x_maps_disparity.py:12-19), so a float32 ``t`` is subtracted / divided / scaled in float32 there; this record
    stores float64 and the kernels normalise in float64 (bit-identical to the reference for int64 and float64
    timestamps).  For float32 input the rounded time column can differ by one in rare ties; convert such data to
    int64 microseconds (what the camera delivers) or float64 before comparing bit for bit."""
    t = np.asarray(t)
    n = len(t)
    dtype = EVENT_DTYPE_F64 if np.issubdtype(t.dtype, np.floating) else EVENT_DTYPE
    ev = np.zeros(n, dtype=dtype)
    ev["x"] = np.asarray(x)
    ev["y"] = np.asarray(y)
    ev["p"] = 1 if p is None else np.asarray(p)
    ev["t"] = t
    return ev


class DeviceEvents:
    """One frame of events resident on a GPU: ``raw`` is an int32 ``[N, 4]`` CUDA tensor."""

    __slots__ = ("raw", "time_f64", "_cache")

    def __init__(self, raw: torch.Tensor, time_f64: bool = False):
        if not raw.is_cuda:
            raise ValueError("DeviceEvents needs a CUDA tensor")
        if raw.dtype != torch.int32 or raw.dim() != 2 or raw.shape[1] != 4 or not raw.is_contiguous():
            raise ValueError("DeviceEvents.raw must be a contiguous int32 [N, 4] tensor")
        if raw.data_ptr() % 16:
            raise ValueError("event buffer must be 16-byte aligned")
        self.raw = raw
        self.time_f64 = bool(time_f64)
        self._cache = {}

    # -- construction ------------------------------------------------------------------------
    @staticmethod
    def from_any(events, device=None, time_f64: Optional[bool] = None) -> "DeviceEvents":
        """Accepts a DeviceEvents, a CUDA tensor holding 16-byte records (uint8 [N,16], int16 [N,8],
        int32 [N,4] or int64 [N,2]), a host EventCD structured array, or a dict of columns."""
        if isinstance(events, DeviceEvents):
            return events
        if isinstance(events, torch.Tensor):
            if not events.is_cuda:
                raise ValueError("event tensors must live on a CUDA device (no CPU path)")
            t = events.contiguous()
            if t.numel() * t.element_size() % EVENT_RECORD_BYTES:
                raise ValueError("event tensor is not a whole number of 16-byte records")
            return DeviceEvents(t.view(torch.uint8).reshape(-1).view(torch.int32).reshape(-1, 4), bool(time_f64))
        if isinstance(events, dict):
            events = pack_events(events["x"], events["y"], events["t"], events.get("p"))
        arr = np.ascontiguousarray(events)
        if arr.dtype.itemsize != EVENT_RECORD_BYTES or arr.dtype.names is None:
            raise ValueError("host events must be an EventCD structured array (16-byte records)")
        if time_f64 is None:
            time_f64 = np.issubdtype(arr.dtype["t"], np.floating)
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        host = torch.from_numpy(arr.view(np.int32).reshape(-1, 4))
        return DeviceEvents(host.to(device, non_blocking=False), bool(time_f64))

    # -- reference-style access --------------------------------------------------------------
    def __len__(self):
        return self.raw.shape[0]

    @property
    def device(self):
        return self.raw.device

    def __getitem__(self, key):
        """``events["x" | "y" | "p" | "t"]`` -> CUDA column (materialised on demand)."""
        if isinstance(key, str):
            if key not in self._cache:
                i16 = self.raw.view(torch.int16)  # [N, 8]
                if key == "x":
                    col = i16[:, 0].to(torch.int32) & 0xFFFF
                elif key == "y":
                    col = i16[:, 1].to(torch.int32) & 0xFFFF
                elif key == "p":
                    col = i16[:, 2].clone()
                elif key == "t":
                    col = self.raw.view(torch.float64 if self.time_f64 else torch.int64)[:, 1].clone()
                else:
                    raise KeyError(key)
                self._cache[key] = col
            return self._cache[key]
        if isinstance(key, slice):
            start, stop, step = key.indices(len(self))
            if step != 1:
                raise IndexError("event buffers only support contiguous slices")
            return DeviceEvents(self.raw[start:stop], self.time_f64)
        raise TypeError("DeviceEvents supports field names and contiguous slices")

    def numpy(self) -> np.ndarray:
        dtype = EVENT_DTYPE_F64 if self.time_f64 else EVENT_DTYPE
        return self.raw.cpu().numpy().view(dtype).reshape(-1)
