"""The oracle's restatement of the rows either side of the depth path (N2 frame
segmentation, N4 de-duplication filters) against golden vectors produced by the real reference
(tests/golden/generate_golden_stream.py runs /root/reference's frame_event_filter.py and
trigger_finder.py on the seeded inputs of tests/stream_cases.py)."""
import os

import numpy as np
import pytest

from oracle import xmaps_oracle as orc
from stream_cases import FILTER_CASES, FILTER_KEYS, TRIGGER_CASES, chunked, filter_inputs, trigger_stream, yt_subset
from xm_helpers import GOLDEN, load_golden_tables

MODES = {"first_yt": orc.FILTER_FIRST_YT, "first_xy": orc.FILTER_FIRST_XY, "last_xy": orc.FILTER_LAST_XY, "mean_xy": orc.FILTER_MEAN_XY}


@pytest.fixture(scope="module")
def golden_filters():
    return np.load(os.path.join(GOLDEN, "stream_filters.npz"))


@pytest.fixture(scope="module")
def golden_trigger():
    return np.load(os.path.join(GOLDEN, "stream_trigger.npz"))


@pytest.mark.parametrize("case", FILTER_CASES, ids=[c[0] for c in FILTER_CASES])
@pytest.mark.parametrize("key", FILTER_KEYS)
def test_filters_match_reference(golden_filters, case, key):
    name, seed, n, p_on = case
    ev = filter_inputs(name, seed, n, p_on)
    if key == "first_yt":
        lut_x = load_golden_tables("small")[0].lut_x
        pos = yt_subset(ev, lut_x)
        got = orc.frame_event_filter(pos, MODES[key], lut_x[pos["y"], pos["x"]])
    else:
        got = orc.frame_event_filter(ev, MODES[key])
    want = golden_filters[f"{name}.{key}"].view(orc.EVENT_DTYPE)
    assert got.dtype == want.dtype and np.array_equal(got, want)


def test_filter_errors_of_the_reference():
    ev = orc.synth_events(1, 100, 160, 120)
    neg = ev.copy()
    neg["p"] = 0
    with pytest.raises(ValueError):  # .max() of an empty array (frame_event_filter.py:24)
        orc.frame_event_filter(neg, orc.FILTER_LAST_XY)
    pos = ev[ev["p"] == 1]
    with pytest.raises(IndexError):  # column wraps below zero (:79)
        orc.frame_event_filter(pos, orc.FILTER_FIRST_YT, np.full(len(pos), -5000, np.int16) + np.arange(len(pos), dtype=np.int16) % 2 * 5100)
    assert orc.frame_event_filter(ev, orc.FILTER_NONE) is ev


@pytest.mark.parametrize("case", TRIGGER_CASES, ids=[c[0] for c in TRIGGER_CASES])
def test_trigger_finder_matches_reference(golden_trigger, case):
    name, seed, frames, per_frame, glitch, chunks = case
    stream = trigger_stream(seed, frames, per_frame, glitch)
    got = []
    tf = orc.TriggerFinderOracle(60, lambda e: got.append(e.copy()))
    for part in chunked(stream, chunks):
        tf.process_events(part)
    assert np.array_equal(np.array([f["t"][0] for f in got], np.int64), golden_trigger[f"{name}.frame_first_t"])
    assert np.array_equal(np.array([len(f) for f in got], np.int64), golden_trigger[f"{name}.frame_len"])
    assert np.array_equal(np.array([int(f["x"].astype(np.int64).sum()) for f in got], np.int64), golden_trigger[f"{name}.frame_sum_x"])
    ok = sum(1 for st, _ in tf.log if st == 1)
    assert [ok, len(tf.log) - ok] == list(golden_trigger[f"{name}.ok_fail"])
    # single-buffer decision on the whole stream
    status, prev_idx, next_idx, _ = orc.find_trigger(stream["t"], 60)
    start, flen, rest = (int(v) for v in golden_trigger[f"{name}.single"])
    assert status == 1 and int(stream["t"][prev_idx + 2]) == start
    assert next_idx - 2 - (prev_idx + 2) == flen and len(stream) - (next_idx - 2) == rest


def trigger_case(step, count):
    """2500 events, a pause, `count` events `step` us apart, a pause, 2500 events."""
    head = np.arange(0, 5000, 2, dtype=np.int64)
    mid = 5100 + np.arange(count, dtype=np.int64) * step
    return np.concatenate((head, mid, mid[-1] + 100 + head))


def test_find_trigger_statuses():
    t = np.arange(0, 5000, 2, dtype=np.int64)
    assert orc.find_trigger(t, 60)[0] == -1  # no pause at all
    assert orc.find_trigger(t[:1], 60)[0] == -1 and orc.find_trigger(t[:0], 60)[0] == -1
    assert orc.find_trigger(trigger_case(6, 2500), 60)[:3] == (1, 2499, 4999)   # 15.1 ms between the pauses
    assert orc.find_trigger(trigger_case(7, 2500), 60)[:3] == (0, 2499, 4999)   # 17.6 ms: longer than a frame
    assert orc.find_trigger(trigger_case(17, 900), 60)[:3] == (0, 2499, 3399)   # long enough, too few events
    assert orc.find_trigger(trigger_case(2, 2500), 60)[0] == -1                 # 5.1 ms: shorter than half a frame


def test_activity_filter_oracle_semantics():
    """The restated Metavision activity filter (UNPINNED): witnesses are the 8 neighbours only, strictly younger
    than the threshold, in stream order, with the state carried across packets."""
    def ev(rows):
        a = np.zeros(len(rows), dtype=orc.EVENT_DTYPE)
        for i, (x, y, t) in enumerate(rows):
            a[i] = (x, y, 1, t)
        return a

    f = orc.ActivityNoiseFilterOracle(8, 8, 100)
    out = f.process_events(ev([(3, 3, 1000), (3, 3, 1010), (4, 4, 1020), (7, 7, 1030), (5, 5, 1119), (5, 5, 1120), (0, 0, 1130), (1, 0, 1229), (1, 1, 1330)]))
    # 1: nothing around; 2: same pixel is no witness; 3: diagonal neighbour 20 us ago; 4: too far; 5: (4,4) 99 us ago;
    # 6: (4,4) exactly 100 us ago does not count; 7: corner, nothing; 8: (0,0) 99 us ago; 9: (1,0) 101 us ago
    assert [tuple(int(v) for v in (e["x"], e["y"], e["t"])) for e in out] == [(4, 4, 1020), (5, 5, 1119), (1, 0, 1229)]
    # carried state: a new packet still sees (1,1) at t = 1330
    assert len(f.process_events(ev([(2, 2, 1400)]))) == 1
    f.reset()
    assert len(f.process_events(ev([(2, 2, 1400)]))) == 0
    # timestamps start at 0: during the first `threshold` us everything passes
    assert len(orc.ActivityNoiseFilterOracle(8, 8, 100).process_events(ev([(2, 2, 99), (7, 7, 100)]))) == 1
