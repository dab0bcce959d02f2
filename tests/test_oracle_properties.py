"""Size-independent properties of the oracle (CPU): they back the shortcuts the CUDA kernels take.

* the integer time column the kernels use, floor((2 dt s + R) / 2R), equals the reference's float64
  expression int16(rint(((t - min) / (max - min)) * s)) except on exact .5 ties (where the kernels fall back to
  the float64 expression) -- DESIGN.md §2 / xm_device.cuh IntCol;
* the de-duplication filters are idempotent and agree with each other when no key repeats;
* the trigger finder's decision does not depend on the time origin or on the pixel coordinates.
"""
import numpy as np
import pytest

from oracle import xmaps_oracle as orc


@pytest.mark.parametrize("seed", range(6))
def test_integer_time_column_equals_float64_expression(seed):
    rng = np.random.default_rng(seed)
    s = 719
    for R in (1, 2, 3, 7, 1000, 16665, 16666, 33333, 999_983, 1_400_000):
        t0 = int(rng.integers(0, 2**40))
        dt = np.unique(np.concatenate((rng.integers(0, R + 1, 20_000), [0, R], np.arange(min(R + 1, 2000)))))
        t = t0 + dt
        want = orc.time_to_xmap_column(t.astype(np.int64), s, t_min=t0, t_max=t0 + R).astype(np.int64)
        num = 2 * dt.astype(object) * s + R  # exact (Python integers)
        q = np.array([int(n) // (2 * R) for n in num], dtype=np.int64)
        tie = np.array([int(n) % (2 * R) == 0 for n in num])
        assert np.array_equal(q[~tie], want[~tie]), f"R = {R}"
        # on exact ties the reference rounds half to even (after its two float64 roundings): either neighbour
        assert np.all((want[tie] == q[tie]) | (want[tie] == q[tie] - 1)), f"R = {R} (ties)"


@pytest.mark.parametrize("mode", [orc.FILTER_FIRST_XY, orc.FILTER_LAST_XY])
@pytest.mark.parametrize("as_reference", [True, False])
def test_xy_filters_are_idempotent(mode, as_reference):
    ev = orc.synth_events(9, 40_000, 160, 120, p_on=0.8)
    once = orc.frame_event_filter(ev, mode, as_reference=as_reference)
    twice = orc.frame_event_filter(once, mode, as_reference=as_reference)
    assert np.array_equal(once, twice)
    key = once["y"].astype(np.int64) * 65536 + once["x"]
    assert np.all(np.diff(key) > 0)  # row-major key order, every key once
    assert np.all(once["p"] == 1)


def test_filters_agree_without_duplicates():
    rng = np.random.default_rng(1)
    cells = rng.permutation(160 * 120)[:5000]
    ev = np.zeros(5000, dtype=orc.EVENT_DTYPE)
    ev["x"], ev["y"], ev["p"] = cells % 160, cells // 160, 1
    ev["t"] = np.sort(rng.integers(0, 16666, 5000))
    ref = orc.frame_event_filter(ev, orc.FILTER_LAST_XY)
    for mode in (orc.FILTER_FIRST_XY, orc.FILTER_MEAN_XY):
        for as_reference in (True, False):
            assert np.array_equal(orc.frame_event_filter(ev, mode, as_reference=as_reference), ref)
    # the same set of events, in key order
    order = np.argsort(cells, kind="stable")
    assert np.array_equal(ref["t"], ev["t"][order])


@pytest.mark.parametrize("seed", range(3))
def test_trigger_decision_is_invariant(seed):
    stream = orc.synth_projector_stream(seed, 8, 3000, 160, 120, glitch_every=3)
    base = orc.find_trigger(stream["t"], 60)
    assert orc.find_trigger(stream["t"] + 123_456_789_012, 60) == base
    shuffled_xy = stream.copy()
    shuffled_xy["x"] = 0
    assert orc.find_trigger(shuffled_xy["t"], 60) == base
    # a stream cut right after the accepted frame's closing pause decides the same
    status, prev_idx, next_idx, _ = base
    if status == 1:
        assert orc.find_trigger(stream["t"][: next_idx + 2], 60)[:3] == (1, prev_idx, next_idx)


def test_bilinear_lookup_reduces_to_nearest_on_integral_positions(tables_default):
    """The opt-in bilinear X-map lookup (not reference behaviour) agrees with the reference's nearest lookup when
    every event sits on integral rectified coordinates and integral time columns (up to the last-ulp cases where the
    float64 column lands just below an integer next to an undefined cell)."""
    tables = tables_default
    ev = orc.synth_events(31, 200_000, tables.cam_w, tables.cam_h)
    span = tables.t_px_scale * 16
    rng = np.random.default_rng(2)
    ev["t"] = np.sort(rng.integers(0, tables.t_px_scale + 1, len(ev))) * 16 + 1000
    ev["t"][0], ev["t"][-1] = 1000, 1000 + span
    near = orc.frame_disparity_map(tables, ev, 0)
    bil = orc.frame_disparity_map_bilinear(tables, tables.lut_x.astype(np.float32), tables.lut_y.astype(np.float32), ev, 0)
    assert near.shape == bil.shape
    differ = np.abs(near - bil) > 1e-3
    assert differ.mean() < 2e-3, differ.mean()
