"""``transient_prbvolpath`` -- host mirror of mitransient/integrators/transient_prbvolpath.py.

Primal render only (the reference's Forward mode raises, :131-133; Backward is AD and out of scope).  The loop
(:170-408), ``sample_emitter`` (:414-512) and ``first_surface`` (:515-528) run on the device.  Supported media:
``homogeneous`` with ``hg`` / ``isotropic`` phase functions behind ``null`` boundaries (what the reference's
participating-media example examples/transient/cornell-box/cbox_volumetric.xml uses).
"""
from __future__ import annotations

from .. import _abi
from ..scene import PerspectiveSensor
from .common import TransientADIntegrator


class TransientPRBVolpathIntegrator(TransientADIntegrator):
    integrator_id = _abi.TR_INTEGRATOR_TRANSIENT_PRBVOLPATH
    plugin_name = "transient_prbvolpath"

    def __init__(self, props):
        super().__init__(props)
        self.use_nee = True                      # forced on by prepare_scene (:110-111)
        self.nee_handle_homogeneous = True
        self.handle_null_scattering = False      # homogeneous media only

    def check_scene_(self, scene, sensor):
        if not isinstance(sensor, PerspectiveSensor):
            raise NotImplementedError("transient_prbvolpath is implemented for the `perspective` sensor")
        if any(e["type"] != _abi.TR_EMITTER_AREA for e in scene.emitters()):
            raise NotImplementedError("transient_prbvolpath supports `area` emitters only")

    def to_string(self):                         # :530-531
        md = -1 if self.max_depth == 0xFFFFFFFF else self.max_depth
        return f"PRBVolpathIntegrator[max_depth = {md}]"

    __repr__ = to_string
