"""Calibration parameters and rectification tables (host side, set-up time only).

Mirrors the call surface of the reference's ``cam_proj_calibration`` module
(/root/reference/python/cam_proj_calibration.py): ``CamProjCalibrationParams`` (:57-140) and
``CamProjMaps`` (:143-331).  Table *construction* is out of scope for the GPU work (SURVEY.md §2
row 4): it happens once, on the host, by handing the same arguments to the same OpenCV entry
points the reference uses (``stereoRectify``, ``initUndistortRectifyMap``, ``undistortPoints``),
so the tables are bit-identical to the reference's.  The *per-frame* methods
(``rectify_cam_coords_i16`` :277-281, ``compute_disp_map_projector_view`` :299-303,
``compute_disp_map_camera_view`` :312-317, ``construct_point_cloud`` :319-331) run on the
GPU through the C-ABI library (see ``engine.py``); there is no CPU fallback.
"""
from __future__ import annotations

import json
from dataclasses import dataclass, field
from typing import Optional

import cv2
import numpy as np

# scale of the rectified image relative to the camera image (reference: from_yaml :84) or the
# projector image (reference: from_ESL_yaml :117)
_RECT_SCALE_CAMERA = 2.75
_RECT_SCALE_ESL = 3


def _matrix_from_node(doc: dict, key: str) -> np.ndarray:
    """Decode one ``opencv_matrix`` node of a calibration document (reference: read_cv_matrix :17-28)."""
    node = doc.get(key)
    if not isinstance(node, dict) or node.get("type-id", "opencv_matrix") != "opencv_matrix":
        raise ValueError(f"Could not read matrix {key} from calibration data")
    return np.asarray(node["data"], dtype=np.float64).reshape(int(node["rows"]), int(node["cols"]))


def _load_calibration_document(path: str) -> dict:
    """YAML in the reference's format, or this repo's JSON fixture (tools/convert_calib.py)."""
    if str(path).lower().endswith(".json"):
        with open(path, "r") as fh:
            return json.load(fh)["matrices"]
    import yaml

    with open(path, "r") as fh:
        return yaml.safe_load(fh)


@dataclass
class CamProjCalibrationParams:
    camera_width: int
    camera_height: int
    projector_width: int
    projector_height: int
    rect_image_width: int
    rect_image_height: int
    camera_K: np.ndarray
    camera_D: np.ndarray
    projector_K: np.ndarray
    projector_D: np.ndarray
    cam2proj_R: np.ndarray
    cam2proj_T: np.ndarray
    F: Optional[np.ndarray] = None

    @staticmethod
    def from_yaml(calibration_yaml_path, camera_width, camera_height, projector_width, projector_height):
        """Reference: from_yaml :77-108.  The projector distortion coefficients are read but
        replaced by zeros, as upstream does (:86-89)."""
        doc = _load_calibration_document(calibration_yaml_path)
        _matrix_from_node(doc, "projector_distortion_coefficients")  # must exist, value unused
        fundamental = _matrix_from_node(doc, "F" if "F" in doc else "fundamental_matrix")
        return CamProjCalibrationParams(
            camera_width=camera_width,
            camera_height=camera_height,
            projector_width=projector_width,
            projector_height=projector_height,
            rect_image_width=round(camera_width * _RECT_SCALE_CAMERA),
            rect_image_height=round(camera_height * _RECT_SCALE_CAMERA),
            camera_K=_matrix_from_node(doc, "camera_intrinsic_matrix"),
            camera_D=_matrix_from_node(doc, "camera_distortion_coefficients"),
            projector_K=_matrix_from_node(doc, "projector_intrinsic_matrix"),
            projector_D=np.zeros((5,)),
            cam2proj_R=_matrix_from_node(doc, "relative_rotation"),
            cam2proj_T=_matrix_from_node(doc, "relative_translation"),
            F=fundamental,
        )

    @staticmethod
    def from_ESL_yaml(calibration_yaml_path, camera_width, camera_height, projector_width, projector_height):
        """Reference: from_ESL_yaml :110-140 (OpenCV FileStorage keys cam_K, cam_kc, proj_K, proj_kc, R, T)."""
        print("Reading calibration from file: {0}".format(calibration_yaml_path))
        store = cv2.FileStorage(calibration_yaml_path, cv2.FILE_STORAGE_READ)
        read = lambda name: store.getNode(name).mat()  # noqa: E731
        return CamProjCalibrationParams(
            camera_width=camera_width,
            camera_height=camera_height,
            projector_width=projector_width,
            projector_height=projector_height,
            rect_image_width=round(projector_width * _RECT_SCALE_ESL),
            rect_image_height=round(projector_height * _RECT_SCALE_ESL),
            camera_K=read("cam_K"),
            camera_D=read("cam_kc"),
            projector_K=read("proj_K"),
            projector_D=read("proj_kc"),
            cam2proj_R=read("R"),
            cam2proj_T=read("T"),
        )


def inverse_rectify_map(K, D, R, P, size, device=None):
    """Map every pixel of an *unrectified* ``size = (W, H)`` image to rectified coordinates
    (reference: initUndistortRectifyMapInverse :31-41): ``cv2.undistortPoints`` over the full
    pixel grid in float32 -- on the host with OpenCV as the reference does, or with ``device`` set by the
    CUDA builder (``engine.build_inverse_lut``, bit-identical)."""
    width, height = size
    if device is not None:
        from .engine import build_inverse_lut

        mx, my = build_inverse_lut(K, D, R, P, size, device=device)
        return mx.cpu().numpy(), my.cpu().numpy()
    grid = np.empty((height * width, 1, 2), dtype=np.float32)
    grid[:, 0, 0] = np.tile(np.arange(width, dtype=np.float32), height)
    grid[:, 0, 1] = np.repeat(np.arange(height, dtype=np.float32), width)
    pts = cv2.undistortPoints(grid, K, D, None, R, P).reshape(height, width, 2)
    return pts[..., 0], pts[..., 1]


def round_map_to_i16(map_f32: np.ndarray) -> np.ndarray:
    """Round-half-even to int16 with a range check (reference: mapf_to_i16 :44-48)."""
    if map_f32.dtype != np.float32:
        raise TypeError("expected a float32 map")
    rounded = np.rint(map_f32)
    info = np.iinfo(np.int16)
    if rounded.min() < info.min or rounded.max() > info.max:
        raise OverflowError("rectification map does not fit int16")
    return rounded.astype(np.int16)


@dataclass
class CamProjMaps:
    """Rectification maps of a camera/projector pair + the per-frame operators that use them."""

    calib: CamProjCalibrationParams
    cam_is_left: bool = False
    zero_undistort_proj_map: bool = False
    table_device: Optional[str] = None  # e.g. "cuda:0": build the inverse LUTs on the GPU (bit-identical to the OpenCV path)

    R1: np.ndarray = field(init=False, repr=False)
    R2: np.ndarray = field(init=False, repr=False)
    P1: np.ndarray = field(init=False, repr=False)
    P2: np.ndarray = field(init=False, repr=False)
    Q: np.ndarray = field(init=False, repr=False)

    def __post_init__(self):
        c = self.calib
        rect_size = (c.rect_image_width, c.rect_image_height)
        first, second = ("camera", "projector") if self.cam_is_left else ("projector", "camera")
        # reference :176-217 — the projector is "camera 1" of the stereo pair by default
        self.R1, self.R2, self.P1, self.P2, self.Q, _roi1, _roi2 = cv2.stereoRectify(
            cameraMatrix1=getattr(c, first + "_K"),
            distCoeffs1=getattr(c, first + "_D"),
            cameraMatrix2=getattr(c, second + "_K"),
            distCoeffs2=getattr(c, second + "_D"),
            imageSize=rect_size,
            R=c.cam2proj_R,
            T=c.cam2proj_T,
            alpha=-1,
        )
        # forward maps rect <- cam / rect <- proj (reference :224-244); note the reference pairs
        # the camera with (R1, P1) and the projector with (R2, P2) irrespective of cam_is_left
        self.camera_mapx, self.camera_mapy = cv2.initUndistortRectifyMap(
            c.camera_K, c.camera_D, self.R1, self.P1, rect_size, cv2.CV_32FC1
        )
        proj_dist = np.zeros(5) if self.zero_undistort_proj_map else c.projector_D
        self.projector_mapx, self.projector_mapy = cv2.initUndistortRectifyMap(
            c.projector_K, proj_dist, self.R2, self.P2, rect_size, cv2.CV_32FC1
        )
        # inverse LUTs cam -> rect (reference :246-254)
        self.disp_cam_mapx_f32, self.disp_cam_mapy_f32 = inverse_rectify_map(
            c.camera_K, c.camera_D, self.R1, self.P1, (c.camera_width, c.camera_height), device=self.table_device
        )
        self.disp_cam_mapx_i16 = round_map_to_i16(self.disp_cam_mapx_f32)
        self.disp_cam_mapy_i16 = round_map_to_i16(self.disp_cam_mapy_f32)
        # inverse LUT proj -> rect, interleaved (x, y) int16 (reference :262-270)
        px, py = inverse_rectify_map(
            c.projector_K, c.projector_D, self.R2, self.P2, (c.projector_width, c.projector_height), device=self.table_device
        )
        self.disp_proj_mapxy_i16 = np.stack((round_map_to_i16(px), round_map_to_i16(py)), axis=-1)

        self._x_map = None  # registered by XMapsDisparity
        self._x_map_consts = None
        self._engine = None

    # ------------------------------------------------------------------ engine plumbing
    def register_x_map(self, x_map: np.ndarray, t_px_scale: int, x_offset: int):
        """Called by ``XMapsDisparity`` once the X-map exists; (re)creates the device context lazily."""
        self._x_map = np.ascontiguousarray(x_map, dtype=np.int16)
        self._x_map_consts = (int(t_px_scale), int(x_offset))
        self._engine = None

    def engine(self, device=None):
        """The device context holding this calibration's tables (created on first use)."""
        from .engine import DepthEngine, TableSet

        if self._x_map is None:
            raise RuntimeError("no X-map registered yet: construct XMapsDisparity(cam_proj_maps=...) first")
        if self._engine is None or (device is not None and self._engine.device_index != _dev_index(device)):
            t_px_scale, x_offset = self._x_map_consts if self._x_map_consts else (0, 0)
            tables = TableSet(
                lut_x=self.disp_cam_mapx_i16,
                lut_y=self.disp_cam_mapy_i16,
                x_map=self._x_map,
                remap_xy=self.disp_proj_mapxy_i16,
                rect_w=self.calib.rect_image_width,
                rect_h=self.calib.rect_image_height,
                t_px_scale=t_px_scale,
                x_offset=x_offset,
                depth_scale=float(self.P2[0, 3]),
                lut_x_f32=self.disp_cam_mapx_f32,
                lut_y_f32=self.disp_cam_mapy_f32,
            )
            self._engine = DepthEngine(tables, device=device)
            from .lazy import register_engine

            register_engine(self._engine)
        return self._engine

    # ------------------------------------------------------------------ per-frame operators
    def _ticket(self, events):
        """Frame ticket for ``events`` (re-used while the same event object flows through the stages)."""
        from .events import DeviceEvents
        from .lazy import FrameTicket, as_device_events

        # Only device buffers are remembered by identity: a host array (Metavision's pooled buffers, a preallocated
        # frame array) may be refilled in place between two calls and must be uploaded again; between the stages of
        # one frame the ticket travels on the lazy handles instead.
        if isinstance(events, DeviceEvents):
            last = getattr(self, "_last_ticket", None)
            if last is not None and last[0] is events:
                return last[1]
            ticket = FrameTicket(self.engine(events.device), events)
            self._last_ticket = (events, ticket)
            return ticket
        self._last_ticket = None
        dev = as_device_events(events)
        return FrameTicket(self.engine(dev.device), dev)

    def rectify_cam_coords_i16(self, events):
        """Reference :277-281.  Returns lazy device columns (x_rect, y_rect) bound to ``events``."""
        from .lazy import DeviceArray

        t = self._ticket(events)
        n = len(t.events)
        return (
            DeviceArray(lambda: t.rect_i16()[0], length=n, ticket=t, role="x_rect_i16"),
            DeviceArray(lambda: t.rect_i16()[1], length=n, ticket=t, role="y_rect_i16"),
        )

    def rectify_cam_coords_f32(self, events):
        """Reference :272-275."""
        from .lazy import DeviceArray

        t = self._ticket(events)
        cache = {}

        def both():
            if "v" not in cache:
                cache["v"] = t.engine.rectify_f32(t.events)
            return cache["v"]

        n = len(t.events)
        return (
            DeviceArray(lambda: both()[0], length=n, ticket=t, role="x_rect_f32"),
            DeviceArray(lambda: both()[1], length=n, ticket=t, role="y_rect_f32"),
        )

    def compute_disp_map_projector_view(self, ev_x_rect_i16, ev_y_rect_i16, inlier_mask, ev_disparity_f32):
        """Reference :299-303.  With the handles produced by ``XMapsDisparity.compute_event_disparity``
        nothing is computed yet (the fused kernel scatters straight from the event buffer);
        with plain arrays the stage-by-stage kernels run."""
        import torch

        from .lazy import DeviceArray, LazyDispMap, to_tensor

        ticket = getattr(ev_disparity_f32, "ticket", None)
        if ticket is not None and getattr(ev_disparity_f32, "role", "") == "disparity":
            return LazyDispMap(ticket, "rect")
        eng = self.engine()
        xr = to_tensor(ev_x_rect_i16, eng.device, torch.int16)
        yr = to_tensor(ev_y_rect_i16, eng.device, torch.int16)
        m = to_tensor(inlier_mask, eng.device).bool()
        d = to_tensor(ev_disparity_f32, eng.device, torch.int16)
        xpr = (xr[m] + d).to(torch.int16)
        c = self.calib
        return DeviceArray.of(eng.scatter_last_wins(yr[m].contiguous(), xpr.contiguous(), d, c.rect_image_height, c.rect_image_width))

    def compute_disp_map_camera_view(self, events, inlier_mask, ev_disparity_f32):
        """Reference :312-317."""
        import torch

        from .lazy import DeviceArray, LazyDispMap, to_tensor

        ticket = getattr(ev_disparity_f32, "ticket", None)
        if ticket is not None and getattr(ev_disparity_f32, "role", "") == "disparity":
            return LazyDispMap(ticket, "cam")
        t = self._ticket(events)
        eng = t.engine
        m = to_tensor(inlier_mask, eng.device).bool()
        d = to_tensor(ev_disparity_f32, eng.device, torch.int16)
        x = t.events["x"][m].to(torch.int16).contiguous()
        y = t.events["y"][m].to(torch.int16).contiguous()
        return DeviceArray.of(eng.scatter_last_wins(y, x, d, self.calib.camera_height, self.calib.camera_width))

    def construct_point_cloud(self, xpr_f32, ypr_f32, disp_f32):
        """Reference :319-331: ``Q @ [x + d, y, -d, 1]`` in float32, dehomogenised, y and z negated."""
        import torch

        from .lazy import DeviceArray, to_tensor

        eng = self.engine()
        x = to_tensor(xpr_f32, eng.device, torch.float32)
        y = to_tensor(ypr_f32, eng.device, torch.float32)
        d = to_tensor(disp_f32, eng.device, torch.float32)
        return DeviceArray.of(eng.point_cloud(x, y, d, self.Q))


def _dev_index(device):
    import torch

    d = torch.device(device)
    return d.index if d.index is not None else torch.cuda.current_device()
