#!/usr/bin/env python
"""Convert an OpenCV-matrix calibration YAML (the format read by the reference's
`CamProjCalibrationParams.from_yaml`, /root/reference/python/cam_proj_calibration.py:77-108)
into the flat JSON fixture `data/esl_calib_hhi.json` that travels with this repo.

Only the numeric calibration values are carried over; the repo never contains the
reference's YAML itself.  Usage:

    python tools/convert_calib.py /root/reference/data/ESL_calib_hhi.yaml data/esl_calib_hhi.json
"""
import json
import sys

import yaml

WANTED = (
    "camera_intrinsic_matrix",
    "camera_distortion_coefficients",
    "projector_intrinsic_matrix",
    "projector_distortion_coefficients",
    "relative_rotation",
    "relative_translation",
    "F",
    "fundamental_matrix",
)


def main(src, dst):
    with open(src, "r") as f:
        doc = yaml.safe_load(f)
    out = {"source": "ESL_calib_hhi (fraunhoferhhi/X-maps data fixture)", "matrices": {}}
    for key in WANTED:
        node = doc.get(key)
        if isinstance(node, dict) and node.get("type-id") == "opencv_matrix":
            out["matrices"][key] = {
                "rows": int(node["rows"]),
                "cols": int(node["cols"]),
                # repr() round-trips float64 exactly
                "data": [float(v) for v in node["data"]],
            }
    with open(dst, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", dst, "with", sorted(out["matrices"]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
