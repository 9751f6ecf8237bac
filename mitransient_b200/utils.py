"""Benchmark scene + small helpers (counterpart of mitransient/utils.py).

``cornell_box()`` returns the same scene *description* as the reference's (mitransient/utils.py:78-220:
pinhole at z = 3.9, five one-sided diffuse walls, two cubes, a 0.46 x 0.38 area light under the
ceiling, transient film 256 x 256 x 300 bins over OPL [3.5, 9.5)) so that configurations are expressed
identically; it is consumed by :func:`mitransient_b200.load_dict`.
"""
from __future__ import annotations

from .scene import ScalarColor3d
from .transforms import ScalarTransform4f as T

speed_of_light = 299792458.0
"""Speed of light in meters/second"""

_CBOX_ALBEDO = {
    "white": (0.885809, 0.698859, 0.666422),
    "green": (0.105421, 0.37798, 0.076425),
    "red": (0.570068, 0.0430135, 0.0443706),
}
_CBOX_LIGHT_RADIANCE = (18.387, 13.9873, 6.75357)


def _wall(to_world, bsdf):
    return {"type": "rectangle", "to_world": to_world, "bsdf": {"type": "ref", "id": bsdf}}


def _box(to_world):
    return {"type": "cube", "to_world": to_world, "bsdf": {"type": "ref", "id": "white"}}


def cornell_box():
    """Dictionary describing the Cornell-box scene for transient rendering."""
    scene = {
        "type": "scene",
        "integrator": {"type": "transient_path", "camera_unwarp": False, "max_depth": 8,
                       "temporal_filter": "box", "gaussian_stddev": 2.0},
        "sensor": {
            "type": "perspective", "fov_axis": "smaller", "near_clip": 0.001, "far_clip": 100.0,
            "focus_distance": 1000, "fov": 39.3077,
            "to_world": T().look_at(origin=[0, 0, 3.90], target=[0, 0, 0], up=[0, 1, 0]),
            "sampler": {"type": "independent", "sample_count": 256},
            "film": {"type": "transient_hdr_film", "width": 256, "height": 256, "rfilter": {"type": "box"},
                     "temporal_bins": 300, "start_opl": 3.5, "bin_width_opl": 0.02},
        },
    }
    for name, rgb in _CBOX_ALBEDO.items():
        scene[name] = {"type": "diffuse", "reflectance": {"type": "rgb", "value": ScalarColor3d(*rgb)}}
    light = _wall(T().translate([0.0, 0.99, 0.01]).rotate([1, 0, 0], 90).scale([0.23, 0.19, 0.19]), "white")
    light["emitter"] = {"type": "area", "radiance": {"type": "rgb", "value": ScalarColor3d(*_CBOX_LIGHT_RADIANCE)}}
    scene["light"] = light
    scene["floor"] = _wall(T().translate([0.0, -1.0, 0.0]).rotate([1, 0, 0], -90), "white")
    scene["ceiling"] = _wall(T().translate([0.0, 1.0, 0.0]).rotate([1, 0, 0], 90), "white")
    scene["back"] = _wall(T().translate([0.0, 0.0, -1.0]), "white")
    scene["green-wall"] = _wall(T().translate([1.0, 0.0, 0.0]).rotate([0, 1, 0], -90), "green")
    scene["red-wall"] = _wall(T().translate([-1.0, 0.0, 0.0]).rotate([0, 1, 0], 90), "red")
    scene["small-box"] = _box(T().translate([0.335, -0.7, 0.38]).rotate([0, 1, 0], -17).scale(0.3))
    scene["large-box"] = _box(T().translate([-0.33, -0.4, -0.28]).rotate([0, 1, 0], 18.25).scale([0.3, 0.61, 0.3]))
    return scene


def cornell_box_config(width, height, temporal_bins, spp, max_depth=8, window=(3.5, 9.5)):
    """The Cornell box with the film overrides BASELINE.md section 4 uses (same OPL window [3.5, 9.5))."""
    d = cornell_box()
    film = d["sensor"]["film"]
    film["width"], film["height"], film["temporal_bins"] = int(width), int(height), int(temporal_bins)
    film["start_opl"] = float(window[0])
    film["bin_width_opl"] = (float(window[1]) - float(window[0])) / int(temporal_bins)
    d["sensor"]["sampler"]["sample_count"] = int(spp)
    d["integrator"]["max_depth"] = int(max_depth)
    return d


def _unit_mesh(kind):
    """Explicit triangles of Mitsuba's unit `rectangle` ([-1,1]^2 in z = 0, normal +z) or `cube` ([-1,1]^3, outward
    normals): (vertices, faces)."""
    from .scene import _CUBE_FACES
    quads = [((-1, -1, 0), (2, 0, 0), (0, 2, 0))] if kind == "rectangle" else _CUBE_FACES
    verts, faces = [], []
    for o, u, v in quads:
        p = [tuple(float(o[i]) for i in range(3)), tuple(float(o[i] + u[i]) for i in range(3)),
             tuple(float(o[i] + u[i] + v[i]) for i in range(3)), tuple(float(o[i] + v[i]) for i in range(3))]
        k = len(verts)
        verts += p
        faces += [(k, k + 1, k + 2), (k, k + 2, k + 3)]
    return verts, faces


def cornell_box_mesh_config(width, height, temporal_bins, spp, max_depth=8, window=(3.5, 9.5), merge_boxes=False):
    """The Cornell box of :func:`cornell_box_config` with every rectangle and cube given as an explicit triangle
    mesh (36 triangles; the light is a mesh area emitter) -- the mesh flavour of the benchmark scene, like the
    reference's examples/transient/cornell-box/cbox_diffuse.xml builds its box from OBJ files.  It exercises the mesh
    pipeline: coplanar triangle pairs -> parallelograms, mesh position sampling by the area CDF; with ``merge_boxes``
    the loader may also turn the two closed cubes back into box primitives (``scene.merge_box``), without it the scene
    keeps its 18 parallelograms."""
    d = cornell_box_config(width, height, temporal_bins, spp, max_depth, window)
    for key, val in list(d.items()):
        if isinstance(val, dict) and val.get("type") in ("rectangle", "cube"):
            v, f = _unit_mesh(val["type"])
            mesh = {k: x for k, x in val.items() if k != "type"}
            mesh.update({"type": "trimesh", "vertices": v, "faces": f, "box_primitive": bool(merge_boxes)})
            d[key] = mesh
    return d


def icosphere(subdivisions=2):
    """Unit sphere as a subdivided icosahedron: (vertices (N,3), faces (M,3)); 20 * 4^subdivisions triangles."""
    import numpy as np
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6), (7, 1, 8),
         (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    v = [np.array(p, dtype=np.float64) / np.linalg.norm(p) for p in v]
    for _ in range(int(subdivisions)):
        mid, nf = {}, []

        def midpoint(a, b):
            key = (min(a, b), max(a, b))
            if key not in mid:
                m = v[a] + v[b]
                v.append(m / np.linalg.norm(m))
                mid[key] = len(v) - 1
            return mid[key]
        for a, b, c in f:
            ab, bc, ca = midpoint(a, b), midpoint(b, c), midpoint(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    return np.array(v), np.array(f, dtype=np.int64)


def cornell_box_sphere_config(width, height, temporal_bins, spp, subdivisions=3, max_depth=8, window=(3.5, 9.5)):
    """The Cornell box with the small box replaced by a triangulated sphere (20 * 4^subdivisions triangles: 320 at 2,
    1280 at 3, 5120 at 4): the scene the BVH path is measured on (more primitives than the lock-step limit of 40)."""
    d = cornell_box_config(width, height, temporal_bins, spp, max_depth, window)
    v, f = icosphere(subdivisions)
    d["small-box"] = {"type": "trimesh", "vertices": v, "faces": f, "merge_quads": False,
                      "to_world": T().translate([0.335, -0.7, 0.38]).scale(0.3), "bsdf": {"type": "ref", "id": "white"}}
    return d


def indent(obj, amount=2):
    lines = str(obj).splitlines(keepends=True)
    if len(lines) <= 1:
        return "".join(lines)
    return "".join(line + " " * amount for line in lines)


def letter_z_mesh(half=0.4, bar=0.1, z=0.0):
    """A flat letter "Z" (two bars + a diagonal, 6 triangles) in the plane ``z``, facing -z -- a procedural
    stand-in for the reference's examples/transient-nlos/Z.obj hidden object."""
    import numpy as np
    h, b = float(half), float(bar)
    quads = [
        [(-h, h - b), (h, h - b), (h, h), (-h, h)],                       # top bar
        [(-h, -h), (h, -h), (h, -h + b), (-h, -h + b)],                   # bottom bar
        [(-h, -h + b), (-h + 2.4 * b, -h + b), (h, h - b), (h - 2.4 * b, h - b)],   # diagonal
    ]
    verts, faces = [], []
    for q in quads:
        i = len(verts)
        verts += [(x, y, z) for x, y in q]
        faces += [(i, i + 2, i + 1), (i, i + 3, i + 2)]                   # clockwise seen from +z: normal -z
    return np.array(verts, dtype=np.float64), np.array(faces, dtype=np.int64)


def nlos_z_scene(width=64, height=64, temporal_bins=300, spp=1000, start_opl=1.85, window=1.8, capture_type="confocal",
                 **integrator):
    """The NLOS "letter Z" setup of examples/transient-nlos/1-simple-nlos-scenes.ipynb (cells 3, 17): hidden Z at
    z = 1 facing the unit relay wall at z = 0, laser and sensor at (-0.5, 0, 0.25)."""
    v, f = letter_z_mesh()
    integ = {"type": "transient_nlos_path", "nlos_laser_sampling": True, "nlos_hidden_geometry_sampling": True,
             "nlos_hidden_geometry_sampling_do_rroulette": False, "account_first_and_last_bounces": False,
             "capture_type": capture_type, "temporal_filter": "box"}
    integ.update(integrator)
    return {
        "type": "scene",
        "geometry": {"type": "trimesh", "vertices": v, "faces": f, "to_world": T().translate([0.0, 0.0, 1.0]),
                     "bsdf": {"type": "diffuse", "reflectance": 1.0}},
        "emitter": {"type": "projector", "irradiance": 100.0, "fov": 0.2, "to_world": T().translate([-0.5, 0.0, 0.25])},
        "relay_wall": {"type": "rectangle", "bsdf": {"type": "diffuse", "reflectance": 1.0},
                       "nlos_sensor": {"type": "nlos_capture_meter",
                                       "sampler": {"type": "independent", "sample_count": int(spp)},
                                       "sensor_origin": [-0.5, 0.0, 0.25],
                                       "transient_film": {"type": "transient_hdr_film", "width": int(width),
                                                          "height": int(height), "temporal_bins": int(temporal_bins),
                                                          "bin_width_opl": float(window) / int(temporal_bins),
                                                          "start_opl": float(start_opl), "rfilter": {"type": "box"}}}},
        "integrator": integ,
    }


def cornell_box_volumetric(width=256, height=384, temporal_bins=512, spp=64, max_depth=16, window=(3.5, 13.5)):
    """The Cornell box with both boxes turned into homogeneous participating media behind `null` boundaries -- the
    dictionary counterpart of examples/transient/cornell-box/cbox_volumetric.xml (albedos / Henyey-Greenstein g of
    that file; sigma_t rescaled from its 556-unit box to this 2-unit box: 0.05 * 278, 0.02 * 278)."""
    d = cornell_box_config(width, height, temporal_bins, spp, max_depth, window)
    d["integrator"] = {"type": "transient_prbvolpath", "max_depth": int(max_depth), "camera_unwarp": False,
                       "temporal_filter": "box"}
    media = {"small-box": ((0.99, 0.4, 0.46), 0.05 * 278.0, 0.1), "large-box": ((0.44, 0.4, 0.96), 0.02 * 278.0, 0.7)}
    for name, (albedo, sigma_t, g) in media.items():
        d[name]["bsdf"] = {"type": "null"}
        d[name]["interior"] = {"type": "homogeneous", "albedo": {"type": "rgb", "value": ScalarColor3d(*albedo)},
                               "sigma_t": sigma_t, "phase": {"type": "hg", "g": g}}
    return d
