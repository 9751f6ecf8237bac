"""``transient_path`` -- host mirror of mitransient/integrators/transientpath.py (class TransientPath).

The per-bounce loop (transientpath.py:141-320) runs inside ``render_transient_path_kernel``
(csrc/tr_render.cu); this class only carries the plugin parameters (transientpath.py:43-86).
"""
from __future__ import annotations

from .. import _abi
from ..scene import PerspectiveSensor
from .common import TransientADIntegrator


class TransientPath(TransientADIntegrator):
    integrator_id = _abi.TR_INTEGRATOR_TRANSIENT_PATH
    plugin_name = "transient_path"

    def check_scene_(self, scene, sensor):
        if not isinstance(sensor, PerspectiveSensor):
            raise NotImplementedError("transient_path is implemented for the `perspective` sensor; NLOS setups use "
                                      "transient_nlos_path + nlos_capture_meter")
        if not scene.emitters():
            return
        if any(e["type"] not in (_abi.TR_EMITTER_AREA, _abi.TR_EMITTER_ANGULARAREA) for e in scene.emitters()):
            raise NotImplementedError("transient_path supports `area` and `angulararea` emitters")
