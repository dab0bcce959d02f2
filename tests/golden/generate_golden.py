#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by running the REAL reference.

Run in the authoring container only (it needs /root/reference, which does not exist on the GPU
box):

    python tests/golden/generate_golden.py

It imports the reference's own modules unmodified from /root/reference/python, feeds them the
seeded inputs of SURVEY.md §8c/§8d and stores what they return.  The committed ``*.npz`` files
are then the pin for ``oracle/xmaps_oracle.py`` (tests/test_oracle_golden.py) and, through the
oracle, for the CUDA path (tests/test_gpu_parity.py).  Nothing here is product code.
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/xmaps_numba_cache")
sys.path.insert(0, os.path.join(REF, "python"))
sys.path.insert(1, ROOT)

import cv2  # noqa: E402
import numpy as np  # noqa: E402

# ---- the reference, unmodified -------------------------------------------------------------
from cam_proj_calibration import CamProjCalibrationParams, CamProjMaps  # noqa: E402
from disp_to_depth import DisparityToDepth, disparity_to_depth_rectified  # noqa: E402
from proj_time_map import ProjectorTimeMap  # noqa: E402
from x_maps_disparity import XMapsDisparity  # noqa: E402

from oracle.xmaps_oracle import synth_events  # noqa: E402  (seeded generator only)

CALIB_YAML = os.path.join(REF, "data", "ESL_calib_hhi.yaml")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


class _NullStats:
    class _T:
        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

    def measure_time(self, key):
        return self._T()


def scaled_params(cam_w, cam_h, proj_w, proj_h, cam_scale, proj_scale, cy_shift=0.0):
    """ESL_calib_hhi with intrinsics scaled, built through the reference's own from_yaml
    (camera K x2, cy x2 - 120)."""
    p = CamProjCalibrationParams.from_yaml(CALIB_YAML, cam_w, cam_h, proj_w, proj_h)
    K = p.camera_K.copy()
    K[:2, :] *= cam_scale
    K[1, 2] += cy_shift
    p.camera_K = K
    Kp = p.projector_K.copy()
    Kp[:2, :] *= proj_scale
    p.projector_K = Kp
    return p


def build_reference_objects(params):
    maps = CamProjMaps(params)
    tm = ProjectorTimeMap.from_calib(params, maps)
    xd = XMapsDisparity(calib_params=params, cam_proj_maps=maps, proj_time_map_rect=tm.projector_time_map_rectified)
    d2d = DisparityToDepth(stats=_NullStats(), calib_params=params, calib_maps=maps, z_near=0.1, z_far=1.0)
    return maps, tm, xd, d2d


def run_reference_frame(maps, xd, d2d, events, camera_view):
    """The per-frame chain of python/depth_reprojection_pipe.py:121-167 with the polarity filter
    restated as p == 1, returning every intermediate."""
    evs = events[events["p"] == 1]
    xr, yr = maps.rectify_cam_coords_i16(evs)
    disp, mask = xd.compute_event_disparity(events=evs, ev_x_rect_i16=xr, ev_y_rect_i16=yr)
    out = {"n_pos": len(evs), "xr": xr, "yr": yr, "disp": disp, "mask": mask}
    if camera_view:
        dm = maps.compute_disp_map_camera_view(events=evs, inlier_mask=mask, ev_disparity_f32=disp)
        out["disp_map"] = dm
    else:
        rect = maps.compute_disp_map_projector_view(
            ev_x_rect_i16=xr, ev_y_rect_i16=yr, inlier_mask=mask, ev_disparity_f32=disp
        )
        out["rect_map"] = rect
        dm = d2d.remap_rectified_disp_map_to_proj(rect)
        out["disp_map"] = dm
    out["depth"] = disparity_to_depth_rectified(dm, maps.P2)
    out["bgr"] = d2d.colorize_depth_from_disp(dm)
    return out


def tables_payload(params, maps, tm, xd):
    return dict(
        lut_x=maps.disp_cam_mapx_i16,
        lut_y=maps.disp_cam_mapy_i16,
        x_map=xd.proj_x_map,
        remap_x=np.ascontiguousarray(maps.disp_proj_mapxy_i16[..., 0]),
        remap_y=np.ascontiguousarray(maps.disp_proj_mapxy_i16[..., 1]),
        rect_wh=np.array([params.rect_image_width, params.rect_image_height], np.int64),
        consts=np.array([xd.T_PX_SCALE, xd.X_OFFSET, xd.X_MAP_WIDTH], np.int64),
        depth_scale=np.array([maps.P2[0, 3]], np.float64),
        Q=maps.Q,
    )


def frame_payload(res, full):
    pay = dict(
        n_pos=np.array([res["n_pos"]], np.int64),
        disp=res["disp"],
        mask_bits=np.packbits(res["mask"]),
        depth=res["depth"],
        disp_map=res["disp_map"],
    )
    if full:
        pay["xr"], pay["yr"] = res["xr"], res["yr"]
        pay["bgr"] = res["bgr"]
        if "rect_map" in res:
            pay["rect_map"] = res["rect_map"]
    return pay


def main():
    manifest = {"numpy": np.__version__, "opencv": cv2.__version__, "configs": {}}

    # ------------------------------------------------------------------ config 1/2: default
    p = CamProjCalibrationParams.from_yaml(CALIB_YAML, 640, 480, 720, 1280)
    maps, tm, xd, d2d = build_reference_objects(p)
    np.savez_compressed(os.path.join(HERE, "tables_default.npz"), **tables_payload(p, maps, tm, xd))
    ev = synth_events(0, 100_000, 640, 480)
    proj = run_reference_frame(maps, xd, d2d, ev, camera_view=False)
    cam = run_reference_frame(maps, xd, d2d, ev, camera_view=True)
    np.savez_compressed(os.path.join(HERE, "frame_default_100k_proj.npz"), **frame_payload(proj, full=True))
    np.savez_compressed(os.path.join(HERE, "frame_default_100k_cam.npz"), **frame_payload(cam, full=False), bgr=cam["bgr"])
    manifest["configs"]["default"] = {
        "geometry": [640, 480, 720, 1280],
        "hash": {
            "lut_x": sha(maps.disp_cam_mapx_i16),
            "lut_y": sha(maps.disp_cam_mapy_i16),
            "lut_x_f32": sha(maps.disp_cam_mapx_f32),
            "lut_y_f32": sha(maps.disp_cam_mapy_f32),
            "x_map": sha(xd.proj_x_map),
            "remap_xy": sha(maps.disp_proj_mapxy_i16),
            "time_map_rect": sha(tm.projector_time_map_rectified),
            "events_seed0_100k": sha(ev),
            "disp": sha(proj["disp"]),
            "mask": sha(proj["mask"]),
            "rect_map": sha(proj["rect_map"]),
            "remapped": sha(proj["disp_map"]),
            "depth_proj": sha(proj["depth"]),
            "depth_cam": sha(cam["depth"]),
            "bgr_proj": sha(proj["bgr"]),
            "bgr_cam": sha(cam["bgr"]),
        },
        "P2_03": float(maps.P2[0, 3]),
        "n_pos": int(proj["n_pos"]),
        "n_inliers": int(proj["mask"].sum()),
    }
    # a 1 M-event frame pinned by hash only (size-independent check of the same tables)
    ev1m = synth_events(1, 1_000_000, 640, 480)
    r1p = run_reference_frame(maps, xd, d2d, ev1m, camera_view=False)
    r1c = run_reference_frame(maps, xd, d2d, ev1m, camera_view=True)
    manifest["configs"]["default"]["hash"].update(
        {"depth_proj_seed1_1m": sha(r1p["depth"]), "depth_cam_seed1_1m": sha(r1c["depth"]), "disp_seed1_1m": sha(r1p["disp"])}
    )

    # ------------------------------------------------------------------ small config (everything stored)
    ps = scaled_params(160, 120, 180, 320, cam_scale=0.25, proj_scale=0.25)
    smaps, stm, sxd, sd2d = build_reference_objects(ps)
    np.savez_compressed(
        os.path.join(HERE, "tables_small.npz"),
        **tables_payload(ps, smaps, stm, sxd),
        time_map_rect=stm.projector_time_map_rectified,
        lut_x_f32=smaps.disp_cam_mapx_f32,
        lut_y_f32=smaps.disp_cam_mapy_f32,
    )
    evs = synth_events(2, 20_000, 160, 120)
    sp = run_reference_frame(smaps, sxd, sd2d, evs, camera_view=False)
    sc = run_reference_frame(smaps, sxd, sd2d, evs, camera_view=True)
    np.savez_compressed(os.path.join(HERE, "frame_small_20k_proj.npz"), **frame_payload(sp, full=True))
    np.savez_compressed(os.path.join(HERE, "frame_small_20k_cam.npz"), **frame_payload(sc, full=False), bgr=sc["bgr"])
    # unsorted timestamps (frame filters may reorder events, python/x_maps_disparity.py:10-11)
    evu = evs.copy()
    np.random.default_rng(7).shuffle(evu)
    su = run_reference_frame(smaps, sxd, sd2d, evu, camera_view=False)
    np.savez_compressed(os.path.join(HERE, "frame_small_20k_shuffled_proj.npz"), **frame_payload(su, full=False))
    manifest["configs"]["small"] = {
        "geometry": [160, 120, 180, 320],
        "cam_scale": 0.25,
        "proj_scale": 0.25,
        "hash": {
            "lut_x": sha(smaps.disp_cam_mapx_i16),
            "x_map": sha(sxd.proj_x_map),
            "remap_xy": sha(smaps.disp_proj_mapxy_i16),
            "time_map_rect": sha(stm.projector_time_map_rectified),
            "depth_proj": sha(sp["depth"]),
            "depth_cam": sha(sc["depth"]),
        },
        "n_inliers": int(sp["mask"].sum()),
    }

    # ------------------------------------------------------------------ config 3: HD (hashes only)
    ph = scaled_params(1280, 720, 1080, 1920, cam_scale=2.0, proj_scale=1.0, cy_shift=-120.0)
    hmaps, htm, hxd, hd2d = build_reference_objects(ph)
    evh = synth_events(3, 1_000_000, 1280, 720)
    hp = run_reference_frame(hmaps, hxd, hd2d, evh, camera_view=False)
    hc = run_reference_frame(hmaps, hxd, hd2d, evh, camera_view=True)
    manifest["configs"]["hd"] = {
        "geometry": [1280, 720, 1080, 1920],
        "cam_scale": 2.0,
        "cy_shift": -120.0,
        "rect_wh": [ph.rect_image_width, ph.rect_image_height],
        "hash": {
            "lut_x": sha(hmaps.disp_cam_mapx_i16),
            "lut_y": sha(hmaps.disp_cam_mapy_i16),
            "x_map": sha(hxd.proj_x_map),
            "remap_xy": sha(hmaps.disp_proj_mapxy_i16),
            "time_map_rect": sha(htm.projector_time_map_rectified),
            "depth_proj_seed3_1m": sha(hp["depth"]),
            "depth_cam_seed3_1m": sha(hc["depth"]),
        },
        "P2_03": float(hmaps.P2[0, 3]),
        "n_inliers": int(hp["mask"].sum()),
    }

    with open(os.path.join(HERE, "manifest.json"), "w") as fh:
        json.dump(manifest, fh, indent=1, sort_keys=True)
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
