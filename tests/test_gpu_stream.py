"""CUDA parity of the rows either side of the depth path: N4 de-duplication filters
(xm_filter_events) and N2 frame segmentation (xm_find_trigger + the RobustTriggerFinder mirror), against
the golden vectors of the real reference and against the oracle.  Bit-exact everywhere."""
import os

import numpy as np
import pytest

from oracle import xmaps_oracle as orc
from stream_cases import FILTER_CASES, FILTER_KEYS, TRIGGER_CASES, chunked, filter_inputs, trigger_stream, yt_subset
from test_gpu_parity import make_engine
from test_oracle_stream import trigger_case
from xm_helpers import GOLDEN, load_golden_tables

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

MODES = {"first_yt": orc.FILTER_FIRST_YT, "first_xy": orc.FILTER_FIRST_XY, "last_xy": orc.FILTER_LAST_XY, "mean_xy": orc.FILTER_MEAN_XY}


@pytest.fixture(scope="module")
def small():
    tables, z = load_golden_tables("small")
    eng = make_engine(tables, z)
    yield tables, eng
    eng.close()


@pytest.fixture(scope="module")
def default():
    tables, z = load_golden_tables("default")
    eng = make_engine(tables)
    yield tables, eng
    eng.close()


def filter_classes():
    import xmaps_b200.frame_event_filter as F

    return {"first_yt": F.FirstEventPerYTFilter, "first_xy": F.FirstEventPerXYFilter, "last_xy": F.LastEventPerXYFilter, "mean_xy": F.MeanFirstLastEventPerXYFilter}


# ------------------------------------------------------------------------------------------------
# N4 filters
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", FILTER_CASES, ids=[c[0] for c in FILTER_CASES])
@pytest.mark.parametrize("key", FILTER_KEYS)
def test_filters_match_reference_golden(small, case, key):
    tables, eng = small
    name, seed, n, p_on = case
    golden = np.load(os.path.join(GOLDEN, "stream_filters.npz"))
    ev = filter_inputs(name, seed, n, p_on)
    if key == "first_yt":
        ev = yt_subset(ev, tables.lut_x)
        xp, _ = eng.rectify_i16(ev)  # what the pipe passes (depth_reprojection_pipe.py:128-131)
        got = filter_classes()[key](engine=eng).filter_events(ev, xp)
    else:
        got = filter_classes()[key](engine=eng).filter_events(ev, None)
    want = golden[f"{name}.{key}"].view(orc.EVENT_DTYPE)
    assert np.array_equal(got.numpy(), want)


@pytest.mark.parametrize("key", FILTER_KEYS)
def test_filters_true_first_event(small, key):
    """as_reference=False: the documented intent (first event per key), against the oracle."""
    tables, eng = small
    ev = filter_inputs("dense_collisions", 3, 50_000, 1.0)
    xp = None
    if key == "first_yt":
        ev = yt_subset(ev, tables.lut_x)
        xp = tables.lut_x[ev["y"], ev["x"]]
    want = orc.frame_event_filter(ev, MODES[key], xp, as_reference=False)
    got = filter_classes()[key](as_reference=False, engine=eng).filter_events(ev, xp)
    assert np.array_equal(got.numpy(), want)
    if key != "last_xy":
        assert not np.array_equal(want, orc.frame_event_filter(ev, MODES[key], xp, as_reference=True))


def test_filter_errors_follow_the_reference(small):
    tables, eng = small
    import xmaps_b200.frame_event_filter as F

    ev = orc.synth_events(1, 500, 160, 120)
    neg = ev.copy()
    neg["p"] = 0
    with pytest.raises(ValueError):
        F.LastEventPerXYFilter(engine=eng).filter_events(neg, None)
    with pytest.raises(ValueError):
        F.LastEventPerXYFilter(engine=eng).filter_events(ev[:0], None)
    xp = tables.lut_x[ev["y"], ev["x"]]
    with pytest.raises(IndexError):  # p != 1 present: the reference's index arrays do not line up
        F.FirstEventPerYTFilter(engine=eng).filter_events(ev, xp)
    pos = ev[ev["p"] == 1]
    bad = np.full(len(pos), 100, np.int16)
    bad[3] = -5000  # wraps below zero
    with pytest.raises(IndexError):
        F.FirstEventPerYTFilter(engine=eng).filter_events(pos, bad)
    assert F.NoFilter().filter_events(ev, None) is ev
    proc = F.FrameEventFilterProcessor()
    names = [str(proc.selected_filter())] + [str(proc.select_next_filter()) for _ in range(5)]
    assert names == ["NoFilter", "FirstEventPerYTFilter", "FirstEventPerXYFilter", "LastEventPerXYFilter", "MeanFirstLastEventPerXYFilter", "NoFilter"]


def test_filters_default_geometry_2m(default):
    tables, eng = default
    ev = orc.synth_events(7, 2_000_000, 640, 480)
    ev["t"] += 5_000_000_000
    for key in ("last_xy", "mean_xy"):
        got = eng.filter_events(ev, MODES[key]).numpy()
        assert np.array_equal(got, orc.frame_event_filter(ev, MODES[key]))
    # filtered events go straight into the depth path (the pipe re-rectifies them, :134-139)
    flt = eng.filter_events(ev, orc.FILTER_LAST_XY)
    want = orc.frame_depth(tables, orc.frame_event_filter(ev, orc.FILTER_LAST_XY), 0)
    e = __import__("xmaps_b200.engine", fromlist=["x"])
    assert np.array_equal(eng.frame(flt, view=0, time_bounds=e.TBOUNDS_REDUCE).cpu().numpy(), want)


# ------------------------------------------------------------------------------------------------
# N2 trigger finder
# ------------------------------------------------------------------------------------------------
def as_events(t):
    ev = np.zeros(len(t), dtype=orc.EVENT_DTYPE)
    ev["t"] = t
    ev["p"] = 1
    return ev


def test_find_trigger_decisions(small):
    tables, eng = small
    cases = [trigger_case(6, 2500), trigger_case(7, 2500), trigger_case(17, 900), trigger_case(2, 2500),
             np.arange(0, 5000, 2, dtype=np.int64), np.arange(0, 400_000, 50, dtype=np.int64),  # no pause / only pauses
             np.arange(1, dtype=np.int64), np.arange(0, dtype=np.int64)]
    for t in cases:
        got = eng.find_trigger(as_events(t), 60)
        want = orc.find_trigger(t, 60)
        assert got[:4] == want, (len(t), got, want)
        if want[0] == 1:
            assert got[4] == t[want[1] + 2] and got[5] == t[want[2] - 2]


@pytest.mark.parametrize("case", TRIGGER_CASES, ids=[c[0] for c in TRIGGER_CASES])
def test_trigger_finder_stream_matches_reference_golden(small, case):
    tables, eng = small
    from xmaps_b200.trigger_finder import RobustTriggerFinder

    name, seed, frames, per_frame, glitch, chunks = case
    golden = np.load(os.path.join(GOLDEN, "stream_trigger.npz"))
    stream = trigger_stream(seed, frames, per_frame, glitch)
    got = []

    class Stats:
        def __init__(self):
            self.counts = {}

        def count(self, k):
            self.counts[k] = self.counts.get(k, 0) + 1

        def add_metric(self, k, v):
            pass

    stats = Stats()
    tf = RobustTriggerFinder(projector_fps=60, stats=stats, frame_callback=lambda e: got.append(e.numpy()), pool=None, engine=eng)
    for part in chunked(stream, chunks):
        tf.process_events(part)
    assert np.array_equal(np.array([f["t"][0] for f in got], np.int64), golden[f"{name}.frame_first_t"])
    assert np.array_equal(np.array([len(f) for f in got], np.int64), golden[f"{name}.frame_len"])
    assert np.array_equal(np.array([int(f["x"].astype(np.int64).sum()) for f in got], np.int64), golden[f"{name}.frame_sum_x"])
    assert [stats.counts.get("trig ✅", 0), stats.counts.get("trig ❌", 0)] == list(golden[f"{name}.ok_fail"])
    status, prev_idx, next_idx, _, start, _ = eng.find_trigger(stream, 60)
    g_start, g_len, g_rest = (int(v) for v in golden[f"{name}.single"])
    assert status == 1 and start == g_start and next_idx - prev_idx - 4 == g_len and len(stream) - (next_idx - 2) == g_rest


def test_stream_to_depth_frames(small):
    """Device-resident stream -> trigger finder -> depth frames == oracle on the oracle's frames."""
    tables, eng = small
    from xmaps_b200.events import DeviceEvents
    from xmaps_b200.trigger_finder import RobustTriggerFinder

    stream = trigger_stream(1, 30, 3000, 5)
    want_frames = []
    otf = orc.TriggerFinderOracle(60, lambda e: want_frames.append(e.copy()))
    depth = []
    tf = RobustTriggerFinder(projector_fps=60, frame_callback=lambda e: depth.append(eng.frame(e, view=0).cpu().numpy()), engine=eng)
    dev = DeviceEvents.from_any(stream)
    i = 0
    for part in chunked(stream, [3500, 1900, 4200]):
        otf.process_events(part)
        tf.process_events(dev[i : i + len(part)])
        i += len(part)
    assert len(depth) == len(want_frames) > 5
    for d, f in zip(depth, want_frames):
        assert np.array_equal(d, orc.frame_depth(tables, f, 0))


def test_drop_frame_rule(small):
    tables, eng = small
    from xmaps_b200.trigger_finder import RobustTriggerFinder

    stream = trigger_stream(1, 30, 3000, 0)
    a, b = [], []
    otf = orc.TriggerFinderOracle(60, lambda e: a.append(len(e)))
    tf = RobustTriggerFinder(projector_fps=60, frame_callback=lambda e: b.append(len(e)), engine=eng)
    for k, part in enumerate(chunked(stream, [3500, 1900, 4200])):
        if k in (4, 9):
            otf.drop_frame()
            tf.drop_frame()
        otf.process_events(part)
        tf.process_events(part)
    assert a == b and len(a) > 3
    assert otf.last_frame_start_us == tf.last_frame_start_us


# ---- N2: activity-noise filter (Metavision semantics restated; the oracle is UNPINNED) ------------------------
def same_events(a, b):
    """field by field (the records have two padding bytes that NumPy does not carry through fancy indexing)"""
    return len(a) == len(b) and all(np.array_equal(a[k], b[k]) for k in ("x", "y", "p", "t"))


def activity_stream(seed, n, w, h, span_us, hot=0):
    """Time-sorted events: uniform background + `hot` bursts on small patches (so that both outcomes occur)."""
    rng = np.random.default_rng(seed)
    ev = np.zeros(n, dtype=orc.EVENT_DTYPE)
    ev["x"] = rng.integers(0, w, n)
    ev["y"] = rng.integers(0, h, n)
    for k in range(hot):
        cx, cy = rng.integers(2, w - 2), rng.integers(2, h - 2)
        idx = rng.choice(n, n // (4 * max(hot, 1)), replace=False)
        ev["x"][idx] = cx + rng.integers(-2, 3, len(idx))
        ev["y"][idx] = cy + rng.integers(-2, 3, len(idx))
    ev["p"] = 1
    ev["t"] = np.sort(rng.integers(0, span_us, n)) + 1_000_000
    return ev


@pytest.mark.parametrize("n,span,thr,hot", [(60_000, 12_000, 16_666, 0), (60_000, 12_000, 800, 3), (40_000, 100_000, 5_000, 2), (30_000, 400, 16_666, 0)])
def test_activity_filter_matches_oracle(small, n, span, thr, hot):
    """One packet: sparse / dense, thresholds shorter and longer than the packet (the long packet of case 3 is cut
    into ~20 sub-packets on the device), borders included."""
    tables, eng = small
    w, h = tables.cam_w, tables.cam_h
    ev = activity_stream(n, n, w, h, span, hot)
    ev["x"][:50] = 0  # image borders
    ev["y"][50:100] = h - 1
    eng.activity_reset()
    want = orc.ActivityNoiseFilterOracle(w, h, thr).process_events(ev)
    got = eng.activity_filter(ev, thr).numpy()
    assert 0 < len(want) < len(ev)
    assert same_events(got, want)


def test_activity_filter_carries_state_across_packets(small):
    """The per-pixel timestamps survive from packet to packet (and activity_reset clears them); ties in time included."""
    tables, eng = small
    w, h = tables.cam_w, tables.cam_h
    ev = activity_stream(11, 90_000, w, h, 30_000, hot=2)
    ev["t"] = (ev["t"] // 7) * 7  # many equal timestamps: only the stream order separates them
    thr = 2_000
    oaf = orc.ActivityNoiseFilterOracle(w, h, thr)
    eng.activity_reset()
    for part in chunked(ev, [1, 7_000, 20_000, 3, 40_000, 0, 22_996]):
        want = oaf.process_events(part)
        got = eng.activity_filter(part, thr).numpy()
        assert same_events(got, want)
    # a fresh state: the very first event of a recording has no witness unless t < threshold (timestamps start at 0)
    eng.activity_reset()
    first = ev[:1].copy()
    assert len(eng.activity_filter(first, thr)) == 0
    first["t"] = thr - 1
    eng.activity_reset()
    assert len(eng.activity_filter(first, thr)) == 1


def test_activity_filter_rejects_unsorted_packets(small):
    tables, eng = small
    ev = activity_stream(3, 5_000, tables.cam_w, tables.cam_h, 10_000)
    ev["t"][2_000] -= 5_000
    eng.activity_reset()
    with pytest.raises(Exception, match="sorted"):
        eng.activity_filter(ev, 1_000)
    eng.activity_reset()
