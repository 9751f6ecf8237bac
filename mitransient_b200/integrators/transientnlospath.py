"""``transient_nlos_path`` -- host mirror of mitransient/integrators/transientnlospath.py (TransientNLOSPath).

Plugin parameters, defaults and rejections follow transientnlospath.py:201-249; ``prepare`` (:251-383: hidden
geometry area distribution, per-pixel sensor targets, laser targets) and ``sample`` (:673-927) run on the device
(csrc/tr_render.cu, NLOS branch).  Exhaustive captures need an ``exhaustive_scan`` film: the illuminated points are
the sensor's scanned points (``force_equal_illumination_scanning=True``, the default) or a ``laser_scan_width x
laser_scan_height`` grid shot through a copy of the projector with the field of view ``illumination_scan_fov``; AD is
out of scope.
"""
from __future__ import annotations

import enum
import warnings

from .. import _abi
from ..sensors.nloscapturemeter import NLOSCaptureMeter
from .common import TransientADIntegrator


class CaptureType(enum.IntEnum):      # transientnlospath.py:9-13
    Single = 1
    Confocal = 2
    Exhaustive = 3


_CAPTURE_TO_ABI = {CaptureType.Single: _abi.TR_CAPTURE_SINGLE, CaptureType.Confocal: _abi.TR_CAPTURE_CONFOCAL,
                   CaptureType.Exhaustive: _abi.TR_CAPTURE_EXHAUSTIVE}


class TransientNLOSPath(TransientADIntegrator):
    integrator_id = _abi.TR_INTEGRATOR_TRANSIENT_NLOS_PATH
    plugin_name = "transient_nlos_path"

    def __init__(self, props):
        super().__init__(props)
        self.filter_depth = int(props.get("filter_depth", -1))
        self.filter_bounces = int(props.get("filter_bounces", -1))
        if self.filter_depth != -1 and self.filter_bounces != -1:
            raise AssertionError("Only use one of filter_depth or filter_bounces")        # :206-208
        if self.filter_bounces != -1:
            self.filter_depth = self.filter_bounces + 1
        if self.filter_depth != -1 and self.max_depth != 0xFFFFFFFF and self.filter_depth >= self.max_depth:
            warnings.warn("You have set filter_depth >= max_depth. This will cause the final image to be all zero.")
        self.discard_direct_paths = bool(props.get("discard_direct_paths", False))
        ct = props.get("capture_type", CaptureType.Single)                                  # :217-229
        if isinstance(ct, CaptureType):
            self.capture_type = ct
        elif isinstance(ct, int):
            self.capture_type = CaptureType(ct)
        elif isinstance(ct, str):
            self.capture_type = CaptureType[ct.capitalize()]
        else:
            raise AssertionError("capture_type must be either an int, a string or a CaptureType enum")
        self.force_equal_grids = bool(props.get("force_equal_illumination_scanning", True))
        self.illumination_scan_fov = float(props.get("illumination_scan_fov", 20.0))
        self.laser_sampling = bool(props.get("nlos_laser_sampling", False))
        self.hg_sampling = bool(props.get("nlos_hidden_geometry_sampling", False))
        self.hg_sampling_do_rroulette = bool(props.get("nlos_hidden_geometry_sampling_do_rroulette", False)) \
            and self.hg_sampling
        self.hg_sampling_includes_relay_wall = bool(
            props.get("nlos_hidden_geometry_sampling_includes_relay_wall", False)) and self.hg_sampling
        self.account_first_and_last_bounces = bool(props.get("account_first_and_last_bounces", False))

    def check_scene_(self, scene, sensor):
        if self.camera_unwarp:                                                              # :725-727
            raise AssertionError("Do not use camera_unwarp with TransientNLOSPath. Use "
                                 "account_first_and_last_bounces instead for the same purpose.")
        from ..scene import PerspectiveSensor
        if not isinstance(sensor, (NLOSCaptureMeter, PerspectiveSensor)):
            raise NotImplementedError("transient_nlos_path is implemented for nlos_capture_meter and perspective sensors")
        ems = scene.emitters()
        if len(ems) != 1:                                                                   # :257-260
            raise ValueError(f"You have defined multiple ({len(ems)}) emitters in the scene with a NLOS capture "
                             "meter. You should have only 1.")
        if ems[0]["type"] != _abi.TR_EMITTER_PROJECTOR:
            raise NotImplementedError("transient_nlos_path expects a `projector` emitter (the laser)")
        film = sensor.film()
        if self.capture_type == CaptureType.Exhaustive:
            if not getattr(film, "exhaustive_scan", False):
                raise ValueError("Exhaustive captures need a transient_hdr_film with exhaustive_scan=True "
                                 "(laser_scan_width / laser_scan_height = the laser grid)")
            if self.force_equal_grids and self.laser_sampling and \
                    (film.laser_scan_width, film.laser_scan_height) != tuple(film.size()):      # :345-348
                raise AssertionError("Sensor and laser scan resolution must be equal if "
                                     "force_equal_illumination_scanning is set to True")
        if self.hg_sampling:                                                                # :277-289
            areas = [0.0 if (s.is_sensor and not self.hg_sampling_includes_relay_wall) else s.surface_area()
                     for s in scene.shapes()]
            if not areas:
                raise AssertionError("Hidden geometry sampling is activated, but there are no hidden geometries "
                                     "in the scene!")
            if sum(areas) < 1e-7:
                raise AssertionError("Hidden geometry sampling is activated, but the hidden geometry in the scene "
                                     "has zero surface area?")

    def render_flags(self) -> int:
        f = super().render_flags()
        if self.laser_sampling:
            f |= _abi.TR_FLAG_NLOS_LASER_SAMPLING
        if self.hg_sampling:
            f |= _abi.TR_FLAG_NLOS_HG_SAMPLING
        if self.hg_sampling_do_rroulette:
            f |= _abi.TR_FLAG_NLOS_HG_SAMPLING_DO_RROULETTE
        if self.hg_sampling_includes_relay_wall:
            f |= _abi.TR_FLAG_NLOS_HG_SAMPLING_INCLUDES_RELAY_WALL
        if self.account_first_and_last_bounces:
            f |= _abi.TR_FLAG_ACCOUNT_FIRST_AND_LAST_BOUNCES
        if self.discard_direct_paths:
            f |= _abi.TR_FLAG_DISCARD_DIRECT_PATHS
        return f

    def fill_params_(self, params, scene, sensor):
        params.capture_type = _CAPTURE_TO_ABI[self.capture_type]
        params.filter_depth = self.filter_depth
        params.force_equal_grids = int(self.force_equal_grids)
        # the widened copy of the projector that scans the illuminated points (transientnlospath.py:349-356):
        # inverse of Mitsuba's perspective_projection(1x1 texture, fov, 1e-4, 1e4)
        from ..transforms import parse_fov, perspective_projection
        fov_x = float(parse_fov(self.illumination_scan_fov, "x", 1.0))
        s2c = perspective_projection((1, 1), (1, 1), (0, 0), fov_x, 1e-4, 1e4).inverse()
        params.laser_scan_sample_to_camera[:] = s2c.matrix.astype("float32").reshape(-1).tolist()
