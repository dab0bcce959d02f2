"""The ~100 %-inlier "plane" workload (SURVEY.md §8c sanity oracle, §8d input 1) and the reference's own
modules under oracle/_ref as a second checker.  CPU only."""
import hashlib

import numpy as np
import pytest

from oracle import ref_chain
from oracle import xmaps_oracle as orc
from xm_helpers import golden_frame, host_time_map_rect


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


@pytest.fixture(scope="module")
def time_map():
    return host_time_map_rect()


@pytest.mark.parametrize("tag", ["z030", "z050", "z080"])
def test_plane_frames_match_reference(tables_default, manifest, time_map, tag):
    cfg = manifest["configs"]["default"]["plane"][tag]
    ev = orc.synth_plane_events(tables_default, time_map, cfg["z"])
    assert len(ev) == cfg["n_events"] and sha(ev) == cfg["events"]
    assert np.all(np.diff(ev["t"]) >= 0)
    evp = orc.polarity_mask(ev)
    xr, yr = orc.rectify_i16(tables_default, evp)
    disp, mask = orc.event_disparity(tables_default, xr, yr, evp["t"])
    assert int(mask.sum()) == cfg["n_inliers"] and sha(disp) == cfg["disp"]
    assert mask.mean() > 0.999  # the point of this workload: (almost) every event is an inlier
    proj = orc.frame_depth(tables_default, ev, orc.VIEW_PROJECTOR)
    cam = orc.frame_depth(tables_default, ev, orc.VIEW_CAMERA)
    assert sha(proj) == cfg["depth_proj"] and sha(cam) == cfg["depth_cam"]
    # physically meaningful: the plane comes back at its depth
    assert abs(float(np.median(cam[cam > 0])) - cfg["z"]) < 2e-3
    assert float(np.median(cam[cam > 0])) == cfg["median_depth_cam"]
    assert sha(orc.colorize(orc.frame_disparity_map(tables_default, ev, 0), tables_default.depth_scale, 0.1, 1.0)) == cfg["bgr_proj"]
    if tag == "z050":
        g = golden_frame("default_plane_z050")
        assert np.array_equal(proj, g["depth_proj"]) and np.array_equal(cam, g["depth_cam"])


def test_plane_burst_variant_matches_reference(tables_default, manifest, time_map):
    cfg = manifest["configs"]["default"]["plane"]["z050_x16"]
    ev = orc.synth_plane_events(tables_default, time_map, cfg["z"], repeat=cfg["repeat"], jitter_us=cfg["jitter_us"], seed=cfg["seed"])
    assert len(ev) == cfg["n_events"] and sha(ev) == cfg["events"]
    assert sha(orc.frame_depth(tables_default, ev, 0)) == cfg["depth_proj"]
    assert sha(orc.frame_depth(tables_default, ev, 1)) == cfg["depth_cam"]


def test_bench_frame_5m_hash(tables_default, manifest):
    """The 5 M-event frame bench.py's CPU arm renders (seed 1000), pinned by the real reference's hash."""
    ev = orc.synth_events(1000, 5_000_000, 640, 480)
    assert sha(orc.frame_depth(tables_default, ev, 0)) == manifest["configs"]["default"]["hash"]["depth_proj_seed1000_5m"]


def test_scatter_duplicate_rule(tables_default):
    """One fancy assignment (what the reference writes) == explicit highest-index-wins."""
    rng = np.random.default_rng(11)
    n = 200_000
    rows, cols = rng.integers(0, 40, n), rng.integers(0, 50, n)
    vals = rng.integers(1, 3000, n).astype(np.int16)
    a = orc._last_write_wins_scatter((40, 50), rows, cols, vals)
    b = orc._last_write_wins_scatter_explicit((40, 50), rows, cols, vals)
    assert np.array_equal(a, b)


@pytest.mark.skipif(not ref_chain.available(), reason="oracle/_ref not built (needs /root/reference at build time)")
def test_restatement_equals_reference_modules(tables_default, time_map):
    """oracle/_ref (the unmodified reference modules) against the NumPy restatement on fresh inputs."""
    rp = ref_chain.RefPath(x_map=tables_default.x_map)
    assert rp.tables_match(tables_default)
    assert np.array_equal(rp.time_map.projector_time_map_rectified, time_map)
    frames = [orc.synth_events(77, 150_000, 640, 480), orc.synth_plane_events(tables_default, time_map, 0.6, repeat=3, jitter_us=4, seed=1)]
    for ev in frames:
        for view in (0, 1):
            assert np.array_equal(rp.frame_depth(ev, view), orc.frame_depth(tables_default, ev, view))
