"""Host-side mirror of the reference's ``trigger_finder`` module
(/root/reference/python/trigger_finder.py): ``RobustTriggerFinder`` cuts the continuous event stream
into projector frames.  Buffering, the frame-drop rule and the "one frame's worth of events" gate are
host logic exactly as in the reference (:119-144); the search itself -- pauses of >= 40 us, the first
two consecutive pauses more than half a frame apart (:146-189) -- runs on the device
(``xm_find_trigger``) over the concatenated buffer, and the frame handed to ``frame_callback`` is a
device slice, so a stream that already lives on the GPU never returns to the host.
"""
from __future__ import annotations

from typing import Callable, List, Optional

import torch

from .events import DeviceEvents
from .lazy import engine_for

MIN_EVENTS_PER_FRAME = 1000


class _NullStats:
    def count(self, *a, **k):
        pass

    def add_metric(self, *a, **k):
        pass


class DeviceEventBufferList:
    """`EventBufferList` (:11-91) over device chunks.  First / last timestamps of every chunk are read
    once when it is appended (two 8-byte copies), so the host-side gates need no further traffic."""

    def __init__(self):
        self._chunks: List[DeviceEvents] = []
        self._spans: List[tuple] = []

    def append(self, evs):
        ev = DeviceEvents.from_any(evs)
        if len(ev) == 0:
            return
        t = ev.raw.view(torch.int64)[:, 1]
        first_last = torch.stack((t[0], t[-1])).cpu().tolist()
        self._chunks.append(ev)
        self._spans.append((int(first_last[0]), int(first_last[1])))

    def clear(self):
        self._chunks.clear()
        self._spans.clear()

    def empty(self):
        return not self._chunks

    def first_ev_time(self):
        return self._spans[0][0] if self._chunks else -1

    def last_ev_time(self):
        return self._spans[-1][1] if self._chunks else -1

    def time_span_us(self):
        first, last = self.first_ev_time(), self.last_ev_time()
        if first < 0 or last < 0:
            return -1
        return last - first

    def num_events(self):
        return sum(len(c) for c in self._chunks)

    def drop(self, drop_len_ms):
        until = self.first_ev_time() + drop_len_ms * 1000
        dropped = False
        while not self.empty() and self.first_ev_time() < until:
            self._chunks.pop(0)
            self._spans.pop(0)
            dropped = True
        return dropped

    def pop_all(self) -> Optional[DeviceEvents]:
        if not self._chunks:
            return None
        raw = self._chunks[0].raw if len(self._chunks) == 1 else torch.cat([c.raw for c in self._chunks])
        self.clear()
        return DeviceEvents(raw.contiguous(), False)

    def push(self, evs: DeviceEvents, first_t: int, last_t: int):
        assert self.empty()
        if len(evs):
            self._chunks.append(evs)
            self._spans.append((first_t, last_t))


class RobustTriggerFinder:
    frame_paused_thresh_us = 40

    def __init__(self, projector_fps, stats=None, frame_callback: Callable = None, pool=None, engine=None):
        self.engine = engine  # None: the engine CamProjMaps created on the events' device
        self.projector_fps = projector_fps
        self.stats = stats if stats is not None else _NullStats()
        self.frame_callback = frame_callback
        self.pool = pool
        self.should_drop = False
        self.last_frame_start_us = -1
        self._ev_buf = DeviceEventBufferList()

    @property
    def frame_len_ms(self):
        return 1e3 / self.projector_fps

    def reset(self):
        self._ev_buf.clear()
        self.should_drop = False
        self.last_frame_start_us = -1

    def drop_frame(self):
        self.should_drop = True

    def process_events(self, evs):
        if hasattr(evs, "numpy") and not isinstance(evs, (DeviceEvents, torch.Tensor)):
            evs = evs.numpy()  # Metavision EventCDBuffer
        self._ev_buf.append(evs)
        if self.should_drop:
            if self._ev_buf.drop(self.frame_len_ms):
                self.stats.count("frames dropped")
                self.should_drop = False
            else:
                return
        if self._ev_buf.empty():
            return
        if self._ev_buf.time_span_us() < 1e6 / self.projector_fps:
            return
        self.stats.add_metric("evs in buf", self._ev_buf.num_events())
        ev_time = self.find_trigger() / 1000
        self.stats.count("trig ✅" if ev_time > 0 else "trig ❌")

    def find_trigger(self):
        last_t = self._ev_buf.last_ev_time()
        evs = self._ev_buf.pop_all()
        eng = self.engine if self.engine is not None else engine_for(evs.device)
        status, prev_idx, next_idx, _, start_time, end_time = eng.find_trigger(
            evs, self.projector_fps, self.frame_paused_thresh_us, MIN_EVENTS_PER_FRAME
        )
        if status == 1:
            self.frame_callback(evs[prev_idx + 2 : next_idx - 2])
            self.stats.add_metric("frame len [ms]", (end_time - start_time) / 1000)
            if self.last_frame_start_us != -1:
                self.stats.add_metric("frame interval [ms]", (start_time - self.last_frame_start_us) / 1000)
            self.last_frame_start_us = start_time
            self._ev_buf.push(evs[next_idx - 2 :], end_time, last_t)
            return start_time
        if status == 0:
            rest = evs[next_idx:]
            if len(rest):
                first_t = int(rest.raw.view(torch.int64)[0, 1].item())
                self._ev_buf.push(rest, first_t, last_t)
        return -1
