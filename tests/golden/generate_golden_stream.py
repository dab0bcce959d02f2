#!/usr/bin/env python
"""Golden vectors for the rows either side of the depth path, produced by the
REAL reference: `frame_event_filter.py` (the four de-duplication filters) and `trigger_finder.py`
(`RobustTriggerFinder`, fed through the test stub of the closed Metavision buffer type).

    python tests/golden/generate_golden_stream.py        # authoring container only (/root/reference)

Writes tests/golden/stream_filters.npz and stream_trigger.npz; inputs are regenerated from seeds by
the tests (oracle.synth_events / oracle.synth_projector_stream), only the reference's outputs are stored.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "tests", "stubs"))  # metavision_sdk_base, stats_printer stand-ins
sys.path.insert(1, os.path.join(REF, "python"))
sys.path.insert(2, ROOT)
sys.path.insert(3, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import frame_event_filter as ref_filters  # noqa: E402  (the reference, unmodified)
import trigger_finder as ref_trigger  # noqa: E402  (the reference, unmodified)
from metavision_sdk_base import EventCDBuffer  # noqa: E402  (stub)
from stats_printer import StatsPrinter  # noqa: E402  (stub)

from oracle import xmaps_oracle as orc  # noqa: E402  (seeded generators only)
from xm_helpers_path import load_small_lut  # noqa: E402

from stream_cases import FILTER_CASES, TRIGGER_CASES, chunked, filter_inputs, trigger_stream, yt_subset  # noqa: E402


class _Pool:
    def return_buf(self, buf):
        pass


def main():
    lut_x = load_small_lut()
    out = {}
    classes = {
        "first_yt": ref_filters.FirstEventPerYTFilter,
        "first_xy": ref_filters.FirstEventPerXYFilter,
        "last_xy": ref_filters.LastEventPerXYFilter,
        "mean_xy": ref_filters.MeanFirstLastEventPerXYFilter,
    }
    for name, seed, n, p_on in FILTER_CASES:
        ev = filter_inputs(name, seed, n, p_on)
        pos = yt_subset(ev, lut_x)
        xp = lut_x[pos["y"], pos["x"]]  # rectify_cam_coords_i16 of the (already positive) frame
        for key, cls in classes.items():
            if key == "first_yt":
                res = cls().filter_events(pos, xp)  # in the pipe the polarity filter ran upstream
            else:
                res = cls().filter_events(ev, None)
            out[f"{name}.{key}"] = np.ascontiguousarray(res).view(np.uint8)
    np.savez_compressed(os.path.join(HERE, "stream_filters.npz"), **out)

    trig = {}
    for name, seed, frames, per_frame, glitch, chunks in TRIGGER_CASES:
        stream = trigger_stream(seed, frames, per_frame, glitch)
        got = []
        stats = StatsPrinter()
        tf = ref_trigger.RobustTriggerFinder(projector_fps=60, stats=stats, frame_callback=lambda e: got.append(e.copy()), pool=_Pool())
        for part in chunked(stream, chunks):
            tf.process_events(EventCDBuffer(part))
        trig[f"{name}.frame_first_t"] = np.array([f["t"][0] for f in got], np.int64)
        trig[f"{name}.frame_len"] = np.array([len(f) for f in got], np.int64)
        trig[f"{name}.frame_sum_x"] = np.array([int(f["x"].astype(np.int64).sum()) for f in got], np.int64)
        trig[f"{name}.ok_fail"] = np.array([stats.counts.get("trig ✅", 0), stats.counts.get("trig ❌", 0)], np.int64)
        # single-buffer decision on the whole stream
        tf2 = ref_trigger.RobustTriggerFinder(projector_fps=60, stats=StatsPrinter(), frame_callback=lambda e: got2.append(e), pool=_Pool())
        got2 = []
        tf2._ev_buf.push(stream)
        start = tf2.find_trigger()
        rest = tf2._ev_buf.num_events()
        trig[f"{name}.single"] = np.array([start, len(got2[0]) if got2 else -1, rest], np.int64)
        print(name, "frames", len(got), "ok/fail", trig[f"{name}.ok_fail"], "single", trig[f"{name}.single"])
    np.savez_compressed(os.path.join(HERE, "stream_trigger.npz"), **trig)


if __name__ == "__main__":
    main()
