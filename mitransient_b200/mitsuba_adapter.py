"""Optional adapter: the B200 path behind a LIVE Mitsuba 3 + mitransient install.

``register()`` re-registers the reference's integrator plugin names (``transient_path`` -- transientpath.py:329,
``transient_nlos_path`` -- transientnlospath.py:930-931, ``transient_prbvolpath`` -- transient_prbvolpath.py:534-535)
with classes whose ``render()`` exports the already-loaded ``mi.Scene`` to this package's scene description
(``scene_to_dict``), renders it through ``libmitransient_b200.so`` and hands the two tensors back as ``mi.TensorXf``.
The film and the sensor stay the reference's own objects (``TransientHDRFilm``, ``NLOSCaptureMeter``): only their
parameters are read.

STATUS: Mitsuba cannot be installed in the environment this repository is built and tested in, so
this module has only been exercised against the small stand-in of tests/test_host.py -- the parameter names it reads
through ``mi.traverse`` follow the Mitsuba 3.6-3.8 plugin documentation and are NOT verified against a live install.
It supports what the C2 / C3 benchmark scenes contain -- triangle meshes (``cube``, ``obj``, ``ply``), ``rectangle``,
``diffuse`` / ``conductor`` / ``dielectric`` / ``null`` / ``twosided`` BSDFs, ``area`` emitters, a ``projector``, a ``perspective`` sensor or the reference's
``nlos_capture_meter`` -- and raises ``NotImplementedError`` for everything else instead of guessing.  The parity path
of this repository does not depend on it.
"""
from __future__ import annotations

from typing import Any, Dict

import numpy as np

INTEGRATOR_NAMES = ("transient_path", "transient_nlos_path", "transient_prbvolpath")


def available() -> bool:
    try:
        import mitsuba as mi
    except ImportError:
        return False
    return mi.variant() is not None


def _np(x, shape=None) -> np.ndarray:
    a = np.array(x, dtype=np.float64)
    a = np.squeeze(a)
    return a.reshape(shape) if shape is not None else a


def _matrix(t) -> np.ndarray:
    """4x4 row-major matrix of a Mitsuba transform (scalar or width-1 vectorised)."""
    m = getattr(t, "matrix", t)
    return _np(m).reshape(4, 4)


def _kind(obj) -> str:
    """Plugin class of a Mitsuba object from its printed form: 'Rectangle[...', 'SmoothDiffuse[...', 'TwoSided[...'."""
    return str(obj).split("[", 1)[0].strip()


def _colour(params, key: str, default=None):
    if key + ".value" in params:
        v = _np(params[key + ".value"]).reshape(-1)
        return {"type": "rgb", "value": [float(v[0]), float(v[min(1, len(v) - 1)]), float(v[min(2, len(v) - 1)])]}
    if key in params:
        return float(_np(params[key]))
    if default is not None:
        return default
    raise NotImplementedError(f"only uniform / rgb textures are supported (parameter {key!r})")


def _bsdf_to_dict(mi, bsdf) -> Dict[str, Any]:
    kind = _kind(bsdf)
    p = mi.traverse(bsdf)
    if kind == "TwoSided":
        inner = {k[len("brdf_0."):]: p[k] for k in p.keys() if k.startswith("brdf_0.")}
        if "eta.value" in inner or ("eta" in inner and "k.value" in inner):
            return {"type": "twosided", "bsdf": {"type": "conductor", "eta": _colour(inner, "eta"), "k": _colour(inner, "k"),
                                                 "specular_reflectance": _colour(inner, "specular_reflectance", 1.0)}}
        if any(k in ("reflectance.value", "reflectance") for k in inner):
            return {"type": "twosided", "bsdf": {"type": "diffuse", "reflectance": _colour(inner, "reflectance")}}
        raise NotImplementedError("twosided: only a nested diffuse or conductor BSDF is exported")
    if kind == "SmoothDiffuse":
        return {"type": "diffuse", "reflectance": _colour(p, "reflectance")}
    if kind == "Null":
        return {"type": "null"}
    if kind == "SmoothConductor":
        return {"type": "conductor", "eta": _colour(p, "eta"), "k": _colour(p, "k"),
                "specular_reflectance": _colour(p, "specular_reflectance", 1.0)}
    if kind == "SmoothDielectric":     # the plugin exposes the relative index int_ior / ext_ior as `eta`
        return {"type": "dielectric", "int_ior": float(_np(p["eta"])), "ext_ior": 1.0,
                "specular_reflectance": _colour(p, "specular_reflectance", 1.0),
                "specular_transmittance": _colour(p, "specular_transmittance", 1.0)}
    raise NotImplementedError(f"BSDF {kind!r} is not exported by the adapter")


def _shape_to_dict(mi, shape) -> Dict[str, Any]:
    p = mi.traverse(shape)
    if shape.is_mesh():              # Mitsuba bakes to_world into the vertex buffer when it loads a mesh
        v = _np(p["vertex_positions"]).reshape(-1, 3)
        f = np.array(p["faces"], dtype=np.int64).reshape(-1, 3)
        d = {"type": "trimesh", "vertices": v, "faces": f}
    elif _kind(shape) == "Rectangle":
        d = {"type": "rectangle", "to_world": _matrix(p["to_world"])}
    else:
        raise NotImplementedError(f"shape {_kind(shape)!r} is not exported by the adapter")
    d["bsdf"] = _bsdf_to_dict(mi, shape.bsdf())
    for side in ("interior", "exterior"):        # media are not exported yet: refuse instead of rendering empty space
        med = getattr(shape, side + "_medium")()
        if med is not None:
            raise NotImplementedError(f"shape {shape.id()!r} has an {side} medium: participating media are not exported "
                                      "by the adapter (build the scene with mitransient_b200.load_file / load_dict)")
    em = shape.emitter()
    if em is not None:
        if _kind(em) != "AreaLight":
            raise NotImplementedError(f"emitter {_kind(em)!r} on a shape is not exported by the adapter")
        d["emitter"] = {"type": "area", "radiance": _colour(mi.traverse(em), "radiance")}
    return d


def _film_to_dict(film) -> Dict[str, Any]:
    """The reference's TransientHDRFilm (a Python object): transient_hdr_film.py:104-121."""
    size, crop, off = film.size(), film.crop_size(), film.crop_offset()
    d = {"type": "transient_hdr_film", "width": int(size[0]), "height": int(size[1]),
         "crop_width": int(crop[0]), "crop_height": int(crop[1]), "crop_offset_x": int(off[0]), "crop_offset_y": int(off[1]),
         "temporal_bins": int(film.temporal_bins), "start_opl": float(film.start_opl), "bin_width_opl": float(film.bin_width_opl),
         "rfilter": {"type": "box"}}
    if getattr(film, "exhaustive_scan", False):
        d.update(exhaustive_scan=True, laser_scan_width=int(film.laser_scan_width), laser_scan_height=int(film.laser_scan_height))
    if not film.rfilter().is_box_filter():
        raise NotImplementedError("only the box reconstruction filter is supported (transient_image_block.py:150-151)")
    return d


def scene_to_dict(mi_scene, sensor, integrator: Dict[str, Any]) -> Dict[str, Any]:
    """``mi.Scene`` + one of its sensors -> the dictionary ``mitransient_b200.load_dict`` reads."""
    import mitsuba as mi
    out: Dict[str, Any] = {"type": "scene", "integrator": dict(integrator)}
    film = _film_to_dict(sensor.film())
    sampler = {"type": "independent", "sample_count": int(sensor.sampler().sample_count())}
    wall = sensor.get_shape() if hasattr(sensor, "get_shape") else None
    for i, shape in enumerate(mi_scene.shapes()):
        d = _shape_to_dict(mi, shape)
        if wall is not None and shape is wall:                      # the relay wall carries the nlos_capture_meter
            d["nlos_sensor"] = {"type": "nlos_capture_meter", "sampler": sampler, "transient_film": film,
                                "sensor_origin": [float(v) for v in _np(sensor.world_transform().translation()).reshape(3)]}
            if getattr(sensor, "is_confocal", False):
                d["nlos_sensor"].update(original_film_width=int(sensor.film_size[0]), original_film_height=int(sensor.film_size[1]))
                # runtime parameters of the 1x1 confocal meter (nloscapturemeter.py:219-228)
                d["nlos_sensor"]["laser_target"] = [float(v) for v in _np(sensor.laser_target).reshape(3)]
                if hasattr(sensor, "laser_bounce_opl"):
                    d["nlos_sensor"]["laser_bounce_opl"] = float(_np(sensor.laser_bounce_opl).reshape(-1)[0])
        out[shape.id() or f"shape{i}"] = d
    if wall is None:
        sp = mi.traverse(sensor)
        out["sensor"] = {"type": "perspective", "to_world": _matrix(sp["to_world"]), "fov": float(_np(sp["x_fov"])),
                         "fov_axis": "x", "near_clip": float(_np(sp["near_clip"])), "far_clip": float(_np(sp["far_clip"])),
                         "sampler": sampler, "film": film}
    for i, em in enumerate(mi_scene.emitters()):
        if em.get_shape() is not None:
            continue                                                # area lights were exported with their shapes
        if _kind(em) != "Projector":
            raise NotImplementedError(f"emitter {_kind(em)!r} is not exported by the adapter")
        ep = mi.traverse(em)
        out[em.id() or f"emitter{i}"] = {"type": "projector", "to_world": _matrix(ep["to_world"]), "fov": float(_np(ep["x_fov"])),
                                         "fov_axis": "x", "irradiance": _colour(ep, "irradiance")}
    return out


def _scene_stamp(mi, mi_scene, sensor) -> bytes:
    """Bytes of every transform an update() can edit after the scene was loaded (cache key of the exported scene)."""
    parts = []
    for obj in list(mi_scene.emitters()) + [sensor]:
        p = mi.traverse(obj)
        for k in p.keys():
            if k.endswith("to_world"):
                parts.append(_matrix(p[k]).tobytes())
    if getattr(sensor, "is_confocal", False):
        parts.append(_np(sensor.laser_target).tobytes())
    return b"".join(parts)


def _make_integrator_class(mi, plugin: str):
    import torch

    import mitransient_b200 as mitr

    class B200Integrator(mi.SamplingIntegrator):
        """``render()`` with the signature of TransientADIntegrator.render (common.py:122-129)."""
        def __init__(self, props):
            super().__init__(props)
            self._props = {k: props[k] for k in props.property_names()} if hasattr(props, "property_names") else dict(props)
            self._impl = mitr.load_dict({"type": plugin, **self._props})      # same names, defaults and rejections
            self._cache = None

        def render(self, scene, sensor=0, seed=0, spp=0, develop=True, evaluate=True, progress_callback=None):
            if not develop:
                raise Exception("develop=True must be specified when invoking AD integrators")       # common.py:130-132
            if isinstance(sensor, int):
                sensor = scene.sensors()[sensor]
            # the exported scene is cached per (scene, sensor) AND the values mi.traverse(...).update() can change between
            # renders (emitter / sensor transforms: mitransient.nlos.focus_emitter_at_relay_wall_*, laser_target)
            stamp = _scene_stamp(mi, scene, sensor)
            if self._cache is None or self._cache[0] is not scene or self._cache[1] is not sensor or self._cache[3] != stamp:
                ours = mitr.load_dict(scene_to_dict(scene, sensor, {"type": plugin, **self._props}))
                self._cache = (scene, sensor, ours, stamp)
            ours = self._cache[2]
            steady, transient = self._impl.render(ours, sensor=0, seed=seed, spp=spp, develop=True, evaluate=evaluate,
                                                  progress_callback=progress_callback)
            if not mi.variant().startswith("cuda"):
                steady, transient = steady.cpu().numpy(), transient.cpu().numpy()
            else:
                torch.cuda.current_stream().synchronize()
            return mi.TensorXf(steady), mi.TensorXf(transient)

        def to_string(self):
            return f"B200Integrator[{plugin}]"

    B200Integrator.__name__ = "B200_" + plugin
    return B200Integrator


def register(prefix: str = "") -> list:
    """Register the B200 integrators with Mitsuba under ``prefix + name`` (``prefix=""`` replaces the reference's
    own registrations, which is what a drop-in does).  Returns the registered names."""
    if not available():
        raise RuntimeError("mitsuba is not importable (or no variant is set): nothing to register with")
    import mitsuba as mi
    names = []
    for plugin in INTEGRATOR_NAMES:
        cls = _make_integrator_class(mi, plugin)
        mi.register_integrator(prefix + plugin, lambda props, cls=cls: cls(props))
        names.append(prefix + plugin)
    return names
