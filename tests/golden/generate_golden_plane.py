#!/usr/bin/env python
"""Golden vectors of the ~100 %-inlier "plane" workload and of the full-size HD frame, from the REAL reference.

Run in the authoring container only (needs /root/reference):

    python tests/golden/generate_golden_plane.py

Adds ``frame_default_plane_z050.npz`` (depth frames of both views for the fronto-parallel plane at Z = 0.5 m,
SURVEY.md §8c "sanity oracle" / §8d input 1: one event per lit camera pixel, the shape of the reference's real
input in python/eval/compute_depth_x_maps.py:83-96) and extends ``manifest.json`` with hashes for the plane at
Z = 0.3 / 0.5 / 0.8, for the 16-events-per-pixel burst variant that ``bench.py --workload plane`` times, and for
the HD geometry (BASELINE config 3) at its full 20 M events.  The event generator is the oracle's seeded
``synth_plane_events``; everything downstream of the events is the unmodified reference.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import generate_golden as gg  # noqa: E402  (imports the reference, unmodified)
import numpy as np  # noqa: E402

from oracle.xmaps_oracle import OracleTables, synth_events, synth_plane_events  # noqa: E402


def oracle_tables(p, maps, xd):
    return OracleTables(
        lut_x=maps.disp_cam_mapx_i16, lut_y=maps.disp_cam_mapy_i16, x_map=xd.proj_x_map,
        remap_xy=maps.disp_proj_mapxy_i16, rect_w=p.rect_image_width, rect_h=p.rect_image_height,
        t_px_scale=xd.T_PX_SCALE, x_offset=xd.X_OFFSET, depth_scale=float(maps.P2[0, 3]),
    )


def main():
    with open(os.path.join(HERE, "manifest.json")) as fh:
        manifest = json.load(fh)

    p = gg.CamProjCalibrationParams.from_yaml(gg.CALIB_YAML, 640, 480, 720, 1280)
    maps, tm, xd, d2d = gg.build_reference_objects(p)
    tables = oracle_tables(p, maps, xd)
    tmr = tm.projector_time_map_rectified
    plane = {}
    for z in (0.3, 0.5, 0.8):
        ev = synth_plane_events(tables, tmr, z)
        proj = gg.run_reference_frame(maps, xd, d2d, ev, camera_view=False)
        cam = gg.run_reference_frame(maps, xd, d2d, ev, camera_view=True)
        vals = cam["depth"][cam["depth"] > 0]
        tag = "z%03d" % round(z * 100)
        plane[tag] = {
            "z": z,
            "n_events": int(len(ev)),
            "n_inliers": int(proj["mask"].sum()),
            "median_depth_cam": float(np.median(vals)),
            "events": gg.sha(ev),
            "disp": gg.sha(proj["disp"]),
            "depth_proj": gg.sha(proj["depth"]),
            "depth_cam": gg.sha(cam["depth"]),
            "bgr_proj": gg.sha(proj["bgr"]),
        }
        if tag == "z050":
            np.savez_compressed(
                os.path.join(HERE, "frame_default_plane_z050.npz"),
                depth_proj=proj["depth"], depth_cam=cam["depth"], n_inliers=np.array([proj["mask"].sum()], np.int64),
                n_events=np.array([len(ev)], np.int64),
            )
    # burst variant (bench.py --workload plane): 16 events per lit pixel, +-8 us jitter
    ev = synth_plane_events(tables, tmr, 0.5, repeat=16, jitter_us=8, seed=5)
    proj = gg.run_reference_frame(maps, xd, d2d, ev, camera_view=False)
    cam = gg.run_reference_frame(maps, xd, d2d, ev, camera_view=True)
    plane["z050_x16"] = {
        "z": 0.5, "repeat": 16, "jitter_us": 8, "seed": 5, "n_events": int(len(ev)), "n_inliers": int(proj["mask"].sum()),
        "events": gg.sha(ev), "depth_proj": gg.sha(proj["depth"]), "depth_cam": gg.sha(cam["depth"]),
    }
    manifest["configs"]["default"]["plane"] = plane

    # 5 M-event uniform frame of the bench (seed 1000, the CPU arm's frame): hash only
    ev5 = synth_events(1000, 5_000_000, 640, 480)
    r5 = gg.run_reference_frame(maps, xd, d2d, ev5, camera_view=False)
    manifest["configs"]["default"]["hash"]["depth_proj_seed1000_5m"] = gg.sha(r5["depth"])

    # HD geometry at its full 20 M events (BASELINE config 3), hashes only
    ph = gg.scaled_params(1280, 720, 1080, 1920, cam_scale=2.0, proj_scale=1.0, cy_shift=-120.0)
    hmaps, htm, hxd, hd2d = gg.build_reference_objects(ph)
    evh = synth_events(3, 20_000_000, 1280, 720)
    hp = gg.run_reference_frame(hmaps, hxd, hd2d, evh, camera_view=False)
    manifest["configs"]["hd"]["hash"]["depth_proj_seed3_20m"] = gg.sha(hp["depth"])
    manifest["configs"]["hd"]["n_inliers_20m"] = int(hp["mask"].sum())
    htables = oracle_tables(ph, hmaps, hxd)
    evp = synth_plane_events(htables, htm.projector_time_map_rectified, 0.5)
    hpp = gg.run_reference_frame(hmaps, hxd, hd2d, evp, camera_view=False)
    manifest["configs"]["hd"]["plane_z050"] = {
        "n_events": int(len(evp)), "n_inliers": int(hpp["mask"].sum()), "events": gg.sha(evp), "depth_proj": gg.sha(hpp["depth"]),
    }

    with open(os.path.join(HERE, "manifest.json"), "w") as fh:
        json.dump(manifest, fh, indent=1, sort_keys=True)
    print(json.dumps(plane, indent=1))
    print(json.dumps(manifest["configs"]["hd"], indent=1))


if __name__ == "__main__":
    main()
