"""Host-side mirror of the reference's ``frame_event_filter`` module
(/root/reference/python/frame_event_filter.py): same class names, ``filter_events(events, xp_i16)``
signature, ``__str__`` and rotation order; the work runs in ``xm_filter_events`` on the device and
the result is a device event buffer (``DeviceEvents``) the depth path takes as it is.

The reference's "First..." filters assign reversed NumPy *views*; NumPy's index iterator negates their
strides and walks them in memory order, so as it runs the reference keeps the LAST event per key in
those filters too.  ``as_reference=True`` (default) reproduces the reference bit for bit;
``as_reference=False`` keeps the first event per key, the documented intent.
"""
from __future__ import annotations

from collections import deque

import numpy as np
import torch

from . import _native as N
from .events import DeviceEvents
from .lazy import engine_for


def _device_events(events) -> DeviceEvents:
    return DeviceEvents.from_any(events)


def _x_rect_tensor(xp_i16, device) -> torch.Tensor:
    t = getattr(xp_i16, "tensor", None)
    if t is None:
        t = xp_i16 if isinstance(xp_i16, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(xp_i16), dtype=np.int16))
    return t.to(device=device, dtype=torch.int16).contiguous()


class FrameEventFilter:
    mode = None
    as_reference = True

    def __init__(self, as_reference: bool = True, engine=None):
        self.as_reference = as_reference
        self.engine = engine  # None: the engine CamProjMaps created on the events' device

    def filter_events(self, events, xp_i16):
        if self.mode is None:
            raise NotImplementedError()
        ev = _device_events(events)
        eng = self.engine if self.engine is not None else engine_for(ev.device)
        if len(ev) == 0 or not bool((ev["p"] == 1).any()):
            # the reference takes .max() of the (empty) positive set (:24)
            raise ValueError("zero-size array to reduction operation maximum which has no identity")
        x_rect = _x_rect_tensor(xp_i16, ev.device) if self.mode == N.FILTER_FIRST_YT else None
        out = eng.filter_events(ev, self.mode, x_rect, self.as_reference)
        st = eng.status()
        if st["filter_polarity"]:
            raise IndexError("shape mismatch: xp_i16 has one entry per event but some events have p != 1")
        if st["filter_index"] or st["pixel_oob"]:
            raise IndexError("index out of bounds for the filter's key image")
        return out


class NoFilter(FrameEventFilter):
    def filter_events(self, events, xp_i16):
        return events

    def __str__(self):
        return "NoFilter"


class LastEventPerXYFilter(FrameEventFilter):
    mode = N.FILTER_LAST_XY

    def __str__(self):
        return "LastEventPerXYFilter"


class FirstEventPerXYFilter(FrameEventFilter):
    mode = N.FILTER_FIRST_XY

    def __str__(self):
        return "FirstEventPerXYFilter"


class FirstEventPerYTFilter(FrameEventFilter):
    mode = N.FILTER_FIRST_YT

    def __str__(self):
        return "FirstEventPerYTFilter"


class MeanFirstLastEventPerXYFilter(FrameEventFilter):
    mode = N.FILTER_MEAN_XY

    def __str__(self):
        return "MeanFirstLastEventPerXYFilter"


class FrameEventFilterProcessor:
    """Rotating selection of the five filters (frame_event_filter.py:131-152)."""

    def __init__(self, as_reference: bool = True, engine=None):
        kw = dict(as_reference=as_reference, engine=engine)
        self.filters = deque(
            (
                NoFilter(**kw),
                FirstEventPerYTFilter(**kw),
                FirstEventPerXYFilter(**kw),
                LastEventPerXYFilter(**kw),
                MeanFirstLastEventPerXYFilter(**kw),
            )
        )

    def selected_filter(self):
        return self.filters[0]

    def filter_events(self, evs, xp_i16):
        return self.selected_filter().filter_events(evs, xp_i16)

    def select_next_filter(self):
        self.filters.rotate(-1)
        return self.selected_filter()
