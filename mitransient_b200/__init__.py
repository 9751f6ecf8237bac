"""mitransient_b200 -- B200-native transient path tracing behind mitransient's plugin surface.

Usage mirrors the reference (examples/transient/0-render_cbox_diffuse.ipynb)::

    import mitransient_b200 as mitr
    scene = mitr.load_dict(mitr.cornell_box())
    steady, transient = mitr.render(scene, spp=256, seed=0)      # (H,W,3), (H,W,T,3) CUDA tensors

All arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI in
``include/mitransient_b200.h``; there is no CPU fallback.
"""
from __future__ import annotations

from . import _abi
from . import films, integrators, sensors            # registers the plugins
from ._lib import TransientError, lib_path
from .films.transient_hdr_film import TransientHDRFilm
from .films.phasor_hdr_film import PhasorHDRFilm
from .integrators.common import TransientADIntegrator
from .integrators.transientnlospath import CaptureType, TransientNLOSPath
from .integrators.transientpath import TransientPath
from .integrators.transient_prbvolpath import TransientPRBVolpathIntegrator
from . import nlos
from . import vis
from .plugins import register_film, register_integrator, register_sensor
from .scene import PluginSpec, Properties, ScalarColor3d, ScalarColor3f, Scene, load_dict, load_obj, load_ply
from .sensors.nloscapturemeter import NLOSCaptureMeter
from .transforms import ScalarTransform4f
from .xml import load_file, xml_to_dict
from .utils import (cornell_box, cornell_box_config, cornell_box_mesh_config, cornell_box_sphere_config, cornell_box_volumetric, icosphere, letter_z_mesh, nlos_z_scene,
                    speed_of_light)

__version__ = "0.1.0"


def render(scene: Scene, params=None, sensor=0, seed: int = 0, spp: int = 0, **kwargs):
    """``mi.render`` for transient scenes: ``scene.integrator().render(scene, sensor, seed, spp)``.
    Returns ``(steady, transient)`` like the reference (mitransient/integrators/common.py:212-213)."""
    integrator = scene.integrator()
    if integrator is None:
        raise ValueError("the scene has no integrator")
    return integrator.render(scene, sensor=sensor, seed=seed, spp=spp, **kwargs)


def set_variant(*names):
    """``mi.set_variant`` for source compatibility.  Everything runs in CUDA fp32; the only thing a variant name
    selects is the colour mode: names containing ``mono`` (``llvm_ad_mono``, ``cuda_mono`` ...) switch to
    monochromatic rendering (colours become their luminance; required by ``phasor_hdr_film``), every other name
    is RGB.  Returns the active variant, ``cuda_rgb`` or ``cuda_mono``."""
    from . import colour_mode as _v
    name = names[0] if names else "cuda_rgb"
    _v.set_monochromatic("mono" in str(name))
    return variant()


def variant() -> str:
    from . import colour_mode as _v
    return "cuda_mono" if _v.is_monochromatic() else "cuda_rgb"
