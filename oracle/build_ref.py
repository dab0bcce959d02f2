#!/usr/bin/env python
"""Recipe for ``oracle/_ref/``: the reference's OWN implementation of the hot path, unmodified.

TEST INFRASTRUCTURE ONLY (see oracle/xmaps_oracle.py).  fraunhoferhhi/X-maps is pure Python: there is
nothing to compile, its path "builds" by placing the five modules the path consists of next to each other
so that they import each other (``x_maps_disparity`` imports ``x_map`` and ``cam_proj_calibration``):

    python/x_maps_disparity.py   compute_disparity / XMapsDisparity            
    python/cam_proj_calibration.py   CamProjMaps.rectify_* / compute_disp_map_* (rows A1, A3)
    python/disp_to_depth.py      remap_rectified_disp_map_to_proj, disparity_to_depth_rectified (A4, A5)
    python/proj_time_map.py, python/x_map.py   set-up tables the constructors need
    data/ESL_calib_hhi.yaml      the calibration BASELINE.json's configs name

The copies go to ``oracle/_ref/`` only, which is git-ignored (nothing of the reference enters the history)
but NOT gpurun-ignored, so it travels to the GPU box like the built ``.so``; ``/root/reference`` does not
exist there.  ``__graft_entry__.build()`` runs this when ``/root/reference`` is present.  The files are used
by ``oracle/ref_chain.py`` as (a) the CPU baseline that ``bench.py`` times (``cpu_baseline.kind =
"reference"``) and (b) a second checker next to the NumPy restatement.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("XMAPS_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
FILES = [
    ("python/x_maps_disparity.py", "x_maps_disparity.py"),
    ("python/cam_proj_calibration.py", "cam_proj_calibration.py"),
    ("python/disp_to_depth.py", "disp_to_depth.py"),
    ("python/proj_time_map.py", "proj_time_map.py"),
    ("python/x_map.py", "x_map.py"),
    ("data/ESL_calib_hhi.yaml", "ESL_calib_hhi.yaml"),
    # The reference's per-frame driver and the two helper modules it imports, in a directory of their own
    # (oracle/_ref/pipe): tests/test_gpu_pipeline.py puts the drop-in modules AHEAD of it on sys.path and runs the
    # unmodified DepthReprojectionPipe.process_ev_frame / process_events body over the CUDA path.
    ("python/depth_reprojection_pipe.py", "pipe/depth_reprojection_pipe.py"),
    ("python/timing_watchdog.py", "pipe/timing_watchdog.py"),
    ("python/event_buf_pool.py", "pipe/event_buf_pool.py"),
]


def build(verbose=False):
    """Returns True if oracle/_ref is in place (freshly copied or already there), False if the reference
    tree is absent and nothing was shipped."""
    have_src = all(os.path.exists(os.path.join(REF_ROOT, s)) for s, _ in FILES)
    if not have_src:
        return all(os.path.exists(os.path.join(OUT, d)) for _, d in FILES)
    os.makedirs(OUT, exist_ok=True)
    digest = {}
    for src, dst in FILES:
        a, b = os.path.join(REF_ROOT, src), os.path.join(OUT, dst)
        os.makedirs(os.path.dirname(b), exist_ok=True)
        shutil.copyfile(a, b)
        os.chmod(b, 0o644)
        with open(b, "rb") as fh:
            digest[dst] = hashlib.sha256(fh.read()).hexdigest()[:16]
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as fh:
        json.dump({"source": REF_ROOT, "sha256_16": digest}, fh, indent=1, sort_keys=True)
    if verbose:
        print("oracle/_ref:", ", ".join(sorted(digest)))
    return True


if __name__ == "__main__":
    ok = build(verbose=True)
    sys.exit(0 if ok else 1)
