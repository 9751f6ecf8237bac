"""Scene ingestion: Mitsuba-style scene dictionaries -> the flat ``TrSceneDesc`` the C ABI takes.

The reference hands ``mi.load_dict(mitr.cornell_box())`` to ``mi.render`` (mitransient/utils.py:78-220,
examples/transient/0-render_cbox_diffuse.ipynb cell 3).  Mitsuba is not available here, so this module
understands the subset of the dictionary format the hot path's configurations use and flattens it into
triangles + shape/bsdf/emitter tables + one sensor/film (include/mitransient_b200.h).

Supported plugin types
  shapes     rectangle, cube, obj (v / f records; ``l`` records are ignored), trimesh (in-memory arrays)
  bsdfs      diffuse, twosided(diffuse), null, ref
  emitters   area (nested in a shape), projector (top level; transient_nlos_path's laser)
  sensors    perspective, nlos_capture_meter (nested in the relay-wall shape)
  films      transient_hdr_film          samplers   independent          rfilter   box
  integrators transient_path, transient_nlos_path, transient_prbvolpath
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Any, Dict, List, Optional

import numpy as np

from . import _abi
from . import colour_mode as _variant
from .transforms import ScalarTransform4f, parse_fov, perspective_projection


def ScalarColor3d(*v):
    """``mi.ScalarColor3d`` stand-in: just a float64 triple."""
    if len(v) == 1:
        v = v[0]
    a = np.asarray(v, dtype=np.float64).reshape(-1)
    if a.size == 1:
        a = np.repeat(a, 3)
    return a


ScalarColor3f = ScalarColor3d


class Properties(dict):
    """Dictionary with the ``mi.Properties.get(name, default)`` access pattern the plugins use."""

    def keys_unqueried(self):
        return [k for k in self.keys() if k not in getattr(self, "_queried", set())]

    def get(self, key, default=None):
        if not hasattr(self, "_queried"):
            self._queried = set()
        self._queried.add(key)
        return super().get(key, default)


# ------------------------------------------------------------------------------------ geometry
_CUBE_FACES = (  # (origin corner, u edge, v edge) with u x v pointing outwards; corners in [-1,1]^3
    ((1, -1, -1), (0, 2, 0), (0, 0, 2)),    # +x
    ((-1, -1, -1), (0, 0, 2), (0, 2, 0)),   # -x
    ((-1, 1, -1), (0, 0, 2), (2, 0, 0)),    # +y
    ((-1, -1, -1), (2, 0, 0), (0, 0, 2)),   # -y
    ((-1, -1, 1), (2, 0, 0), (0, 2, 0)),    # +z
    ((-1, -1, -1), (0, 2, 0), (2, 0, 0)),   # -z
)


def _with_p3(prims3: np.ndarray, types: np.ndarray) -> np.ndarray:
    """[n,3,3] planar primitives -> [n,4,3] (the fourth vertex is only meaningful for boxes)."""
    p3 = np.where((types == _abi.TR_PRIM_QUAD)[:, None], prims3[:, 1] + prims3[:, 2] - prims3[:, 0], prims3[:, 0])
    return np.concatenate([prims3, p3[:, None, :]], axis=1)


def _cube_prims(as_box: bool = True):
    """Mitsuba `cube` ([-1,1]^3): ONE solid box primitive (corner + three edges), or six outward-facing
    parallelogram primitives when ``as_box`` is False."""
    if as_box:
        c = np.array([[(-1, -1, -1), (1, -1, -1), (-1, 1, -1), (-1, -1, 1)]], dtype=np.float64)
        return c, np.full(1, _abi.TR_PRIM_BOX, dtype=np.uint8)
    prims = []
    for o, u, v in _CUBE_FACES:
        o, u, v = np.array(o, float), np.array(u, float), np.array(v, float)
        prims.append([o, o + u, o + v])
    types = np.full(len(prims), _abi.TR_PRIM_QUAD, dtype=np.uint8)
    return _with_p3(np.array(prims, dtype=np.float64), types), types


def _rectangle_prims():
    """Mitsuba `rectangle` ([-1,1]^2 in z = 0, normal +z): one parallelogram primitive."""
    return (np.array([[(-1, -1, 0), (1, -1, 0), (-1, 1, 0), (1, 1, 0)]], dtype=np.float64),
            np.full(1, _abi.TR_PRIM_QUAD, dtype=np.uint8))


def load_obj(path: str) -> np.ndarray:
    """Wavefront OBJ -> float64 [n_tris, 3, 3] triangle primitives.  Polygons are fan-triangulated; ``l``, ``vt``, ``vn``, ``o``,
    ``s`` ... records are ignored (face normals are used, as Mitsuba does for meshes without ``vn``)."""
    verts: List[List[float]] = []
    tris: List[List[int]] = []
    with open(path, "r") as f:
        for line in f:
            if line.startswith("v "):
                p = line.split()
                verts.append([float(p[1]), float(p[2]), float(p[3])])
            elif line.startswith("f "):
                idx = []
                for tok in line.split()[1:]:
                    i = int(tok.split("/")[0])
                    idx.append(i - 1 if i > 0 else len(verts) + i)
                for k in range(1, len(idx) - 1):
                    tris.append([idx[0], idx[k], idx[k + 1]])
    v = np.asarray(verts, dtype=np.float64)
    t = np.asarray(tris, dtype=np.int64)
    if t.size == 0:
        return np.zeros((0, 3, 3))
    return v[t]


_PLY_TYPES = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2", "ushort": "u2",
              "uint16": "u2", "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4", "float": "f4", "float32": "f4",
              "double": "f8", "float64": "f8"}


def load_ply(path: str) -> np.ndarray:
    """Stanford PLY (ascii, binary_little_endian or binary_big_endian) -> float64 [n_tris, 3, 3] triangle primitives.
    Vertex properties other than x, y, z (normals, uv, colours) are skipped -- face normals are used, as for OBJ; polygon
    faces are fan-triangulated; elements other than `vertex` and `face` are skipped."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt, elements = None, []
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: unterminated PLY header")
            tok = line.decode("ascii", "replace").split()
            if not tok or tok[0] in ("comment", "obj_info"):
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                elements.append({"name": tok[1], "count": int(tok[2]), "props": []})
            elif tok[0] == "property":
                if tok[1] == "list":
                    elements[-1]["props"].append(("list", _PLY_TYPES[tok[2]], _PLY_TYPES[tok[3]], tok[4]))
                else:
                    elements[-1]["props"].append(("scalar", _PLY_TYPES[tok[1]], None, tok[2]))
            elif tok[0] == "end_header":
                break
        if fmt not in ("ascii", "binary_little_endian", "binary_big_endian"):
            raise ValueError(f"{path}: unsupported PLY format {fmt!r}")
        verts, tris = None, []
        if fmt == "ascii":
            tokens = f.read().split()
            pos = 0
            for el in elements:
                rows = []
                for _ in range(el["count"]):
                    row = {}
                    for kind, t0, t1, name in el["props"]:
                        if kind == "scalar":
                            row[name] = float(tokens[pos]); pos += 1
                        else:
                            n = int(tokens[pos]); pos += 1
                            row[name] = [int(x) for x in tokens[pos:pos + n]]; pos += n
                    rows.append(row)
                if el["name"] == "vertex":
                    verts = np.array([[r["x"], r["y"], r["z"]] for r in rows], dtype=np.float64).reshape(-1, 3)
                elif el["name"] == "face":
                    key = next(n for k, _, _, n in el["props"] if k == "list")
                    for r in rows:
                        idx = r[key]
                        tris.extend([idx[0], idx[k], idx[k + 1]] for k in range(1, len(idx) - 1))
        else:
            e = "<" if fmt == "binary_little_endian" else ">"
            for el in elements:
                if all(k == "scalar" for k, _, _, _ in el["props"]):
                    dt = np.dtype([(name, e + t0) for _, t0, _, name in el["props"]])
                    data = np.frombuffer(f.read(dt.itemsize * el["count"]), dtype=dt, count=el["count"])
                    if el["name"] == "vertex":
                        verts = np.stack([data["x"], data["y"], data["z"]], axis=1).astype(np.float64)
                    continue
                for _ in range(el["count"]):           # an element with list properties: row by row
                    for kind, t0, t1, name in el["props"]:
                        if kind == "scalar":
                            f.read(np.dtype(t0).itemsize)
                            continue
                        n = int(np.frombuffer(f.read(np.dtype(t0).itemsize), dtype=e + t0)[0])
                        idx = np.frombuffer(f.read(np.dtype(t1).itemsize * n), dtype=e + t1).astype(np.int64)
                        if el["name"] == "face" and name in ("vertex_indices", "vertex_index"):
                            tris.extend([idx[0], idx[k], idx[k + 1]] for k in range(1, n - 1))
    if verts is None:
        raise ValueError(f"{path}: PLY file without a vertex element")
    t = np.asarray(tris, dtype=np.int64).reshape(-1, 3)
    if t.size == 0:
        return np.zeros((0, 3, 3))
    if t.min() < 0 or t.max() >= len(verts):
        raise ValueError(f"{path}: face index out of range")
    return verts[t]


def merge_quads(tris: np.ndarray, rel_eps: float = 1e-6):
    """Greedily merge pairs of triangles that share an edge and together form a parallelogram into one
    TR_PRIM_QUAD (p0, p1, p2 with the fourth corner p1 + p2 - p0).  Geometry and orientation are unchanged;
    a ray tests one primitive instead of two.  Returns (prims [n,3,3], types [n])."""
    n = len(tris)
    used = np.zeros(n, dtype=bool)
    scale = max(1e-30, float(np.abs(tris).max())) if n else 1.0
    eps = rel_eps * scale
    key = lambda p: tuple(np.round(p / max(eps, 1e-30)).astype(np.int64))   # noqa: E731
    edge_map: Dict[Any, List[int]] = {}
    for i in range(n):
        ks = [key(tris[i, k]) for k in range(3)]
        for a, b in ((0, 1), (1, 2), (2, 0)):
            edge_map.setdefault(frozenset((ks[a], ks[b])), []).append(i)
    prims, types = [], []
    for i in range(n):
        if used[i]:
            continue
        merged = False
        ki = [key(tris[i, k]) for k in range(3)]
        for a, b, c in ((0, 1, 2), (1, 2, 0), (2, 0, 1)):        # shared edge (a, b), apex c of triangle i
            for j in edge_map.get(frozenset((ki[a], ki[b])), []):
                if j == i or used[j]:
                    continue
                kj = [key(tris[j, k]) for k in range(3)]
                other = [k for k in range(3) if kj[k] not in (ki[a], ki[b])]
                if len(other) != 1:
                    continue
                pa, s1, s2, pb = tris[i, c], tris[i, a], tris[i, b], tris[j, other[0]]
                if np.abs((pa + pb) - (s1 + s2)).max() > 4 * eps:
                    continue                                      # not a parallelogram
                ni = np.cross(tris[i, 1] - tris[i, 0], tris[i, 2] - tris[i, 0])
                nj = np.cross(tris[j, 1] - tris[j, 0], tris[j, 2] - tris[j, 0])
                if np.dot(ni, nj) <= 0:
                    continue                                      # opposite orientation
                prims.append([pa, s1, s2])                        # cyclic order (c, a, b) keeps triangle i's normal
                types.append(_abi.TR_PRIM_QUAD)
                used[i] = used[j] = True
                merged = True
                break
            if merged:
                break
        if not merged:
            used[i] = True
            prims.append(tris[i])
            types.append(_abi.TR_PRIM_TRIANGLE)
    return (np.array(prims, dtype=np.float64).reshape(-1, 3, 3), np.array(types, dtype=np.uint8))


def merge_box(prims: np.ndarray, types: np.ndarray, rel_eps: float = 1e-6):
    """Six outward-facing parallelograms that are the faces of ONE parallelepiped (an OBJ / PLY cube after
    :func:`merge_quads`, like the boxes of the reference's examples/transient/cornell-box/meshes) -> one TR_PRIM_BOX
    primitive (corner + three edges), which a ray tests with three slab intervals instead of six plane tests.
    Returns ``(box [1,4,3], types [1])`` or ``None`` when the primitives are anything else."""
    if len(prims) != 6 or not np.all(types == _abi.TR_PRIM_QUAD):
        return None
    q = prims[:, :3, :]
    corners = np.concatenate([q, (q[:, 1] + q[:, 2] - q[:, 0])[:, None, :]], axis=1).reshape(-1, 3)      # 24 corners
    eps = rel_eps * max(1e-30, float(np.abs(corners).max()))
    keys = np.round(corners / eps).astype(np.int64)
    uniq, inverse = np.unique(keys, axis=0, return_inverse=True)
    if len(uniq) != 8:
        return None
    pts = np.array([corners[inverse == i].mean(axis=0) for i in range(8)])
    # the corner with the smallest key and its three neighbours (the corners it shares a face edge with)
    adj = {i: set() for i in range(8)}
    idx = inverse.reshape(6, 4)                       # per face: p0, p1, p2, p3 = p1 + p2 - p0  -> edges p0p1, p0p2, p1p3, p2p3
    for f in idx:
        for a, b in ((0, 1), (0, 2), (1, 3), (2, 3)):
            adj[f[a]].add(f[b]); adj[f[b]].add(f[a])
    if any(len(v) != 3 for v in adj.values()):
        return None
    p0 = pts[0]
    e = [pts[j] - p0 for j in sorted(adj[0])]
    if np.linalg.det(np.array(e)) < 0:
        e[1], e[2] = e[2], e[1]
    want = np.array([p0 + a * e[0] + b * e[1] + c * e[2] for a in (0, 1) for b in (0, 1) for c in (0, 1)])
    if any(np.abs(want - p).max(axis=1).min() > 4 * eps for p in pts):
        return None                                   # not a parallelepiped
    centre = pts.mean(axis=0)
    for f in q:                                       # every face must look outwards (a room built from a cube stays a mesh)
        n = np.cross(f[1] - f[0], f[2] - f[0])
        if np.dot(n, (f[1] + f[2]) * 0.5 - centre) <= 0:
            return None
    box = np.array([[p0, p0 + e[0], p0 + e[1], p0 + e[2]]], dtype=np.float64)
    return box, np.full(1, _abi.TR_PRIM_BOX, dtype=np.uint8)


def _apply(to_world: ScalarTransform4f, prims: np.ndarray, flip: bool) -> np.ndarray:
    m = to_world.matrix
    out = prims @ m[:3, :3].T + m[:3, 3]
    if (np.linalg.det(m[:3, :3]) < 0) != bool(flip):
        out = out[:, [0, 2, 1, 3], :]  # swap e1/e2: keeps the geometric normal equal to the transformed normal
    return np.ascontiguousarray(out)


def _rgb(v, default=None) -> np.ndarray:
    if v is None:
        v = default
    if isinstance(v, dict):
        t = v.get("type", "rgb")
        if t not in ("rgb", "uniform", "spectrum"):
            raise ValueError(f"unsupported spectrum/texture type {t!r} (only constant rgb values)")
        v = v.get("value")
    return ScalarColor3d(v)


def _to_transform(v) -> ScalarTransform4f:
    if v is None:
        return ScalarTransform4f()
    if isinstance(v, ScalarTransform4f):
        return v
    return ScalarTransform4f(np.asarray(v, dtype=np.float64))


# ------------------------------------------------------------------------------------ deferred plugins
_SHAPE_TYPES = ("rectangle", "cube", "obj", "ply", "trimesh")
_BSDF_TYPES = ("diffuse", "twosided", "null", "conductor", "dielectric")

# Named indices of refraction accepted by `dielectric` (int_ior / ext_ior), Mitsuba's ior.h table [upstream]
_IOR_TABLE = {"vacuum": 1.0, "helium": 1.00004, "hydrogen": 1.00013, "air": 1.000277, "carbon dioxide": 1.00045,
              "water": 1.3330, "acetone": 1.36, "ethanol": 1.361, "carbon tetrachloride": 1.461, "glycerol": 1.4729,
              "benzene": 1.501, "silicone oil": 1.52045, "bromine": 1.661, "water ice": 1.31, "fused quartz": 1.458,
              "pyrex": 1.470, "acrylic glass": 1.49, "polypropylene": 1.49, "bk7": 1.5046, "sodium chloride": 1.544,
              "amber": 1.55, "pet": 1.575, "diamond": 2.419}


def _ior(v, default: str) -> float:
    v = default if v is None else v
    if isinstance(v, str):
        if v.lower() not in _IOR_TABLE:
            raise ValueError(f"unknown index of refraction name {v!r}")
        return float(_IOR_TABLE[v.lower()])
    return float(v)


class PluginSpec(dict):
    """What ``load_dict`` returns for a shape / emitter / bsdf / medium dictionary: the dictionary itself, kept
    until a scene is assembled (the reference builds its NLOS scenes this way: ``geometry = mi.load_dict({...})``
    first, ``mi.load_dict({'type': 'scene', 'geometry': geometry, ...})`` later).  Mutable on purpose: e.g.
    ``mitransient_b200.nlos.focus_emitter_at_relay_wall_pixel`` rewrites a projector's ``to_world``."""

    def plugin_type(self):
        return self.get("type")

    # -- the bits of the mi.Shape / mi.Emitter interface the reference's helpers use
    def world_transform(self) -> ScalarTransform4f:
        return _to_transform(self.get("to_world"))

    def sensor(self):
        """The sensor attached to this shape (instantiated once, shared with the scene)."""
        for k, v in self.items():
            if _type_of(v) == "nlos_capture_meter":
                if isinstance(v, dict):
                    from . import plugins
                    v = plugins.instantiate(v)
                    self[k] = v
                return v
        raise ValueError("this shape carries no sensor")

    def sample_position(self, time, uv, active=True):
        """Shape::sample_position for a `rectangle`: to_world * (2u - 1, 2v - 1, 0)."""
        if self.get("type") != "rectangle":
            raise NotImplementedError("sample_position is provided for rectangles (the relay wall)")
        p = self.world_transform().transform_affine([2.0 * float(uv[0]) - 1.0, 2.0 * float(uv[1]) - 1.0, 0.0])

        class _PS:
            pass
        ps = _PS()
        ps.p = p
        return ps


def _type_of(v):
    if isinstance(v, dict):
        return v.get("type")
    return getattr(v, "plugin_name", None)


def _find_child(d, types):
    """First nested plugin (dictionary or instance, under ANY key) whose type is in ``types``."""
    for k, v in d.items():
        if k in ("interior", "exterior"):
            continue
        if _type_of(v) in types:
            return k, v
    return None, None


# ------------------------------------------------------------------------------------ host-side scene objects
class Shape:
    def __init__(self, name, kind, prims, prim_types, to_world, bsdf_index, flip=False):
        self.name, self.kind, self.prims, self.prim_types, self.to_world = name, kind, prims, prim_types, to_world
        self.bsdf_index = bsdf_index
        self.emitter_index = -1
        self.is_sensor = False
        self.interior_medium = -1
        self.exterior_medium = -1
        self.flip = flip

    def surface_area(self) -> float:
        e1 = self.prims[:, 1] - self.prims[:, 0]
        e2 = self.prims[:, 2] - self.prims[:, 0]
        e3 = self.prims[:, 3] - self.prims[:, 0]
        a12 = np.linalg.norm(np.cross(e1, e2), axis=1)
        box = 2.0 * (a12 + np.linalg.norm(np.cross(e2, e3), axis=1) + np.linalg.norm(np.cross(e3, e1), axis=1))
        area = np.where(self.prim_types == _abi.TR_PRIM_BOX, box,
                        np.where(self.prim_types == _abi.TR_PRIM_QUAD, 1.0, 0.5) * a12)
        return float(area.sum())


class Scene:
    """What ``mi.load_dict`` returns for our purposes: the parsed plugins + the flattened description."""

    def __init__(self):
        self.shapes_: List[Shape] = []
        self.bsdfs_: List[Dict[str, Any]] = []
        self.emitters_: List[Dict[str, Any]] = []
        self.media_: List[Dict[str, Any]] = []
        self.sensors_: List[Any] = []
        self.integrator_ = None
        self._handle = None          # device scene (mitransient_b200._lib.DeviceScene)
        self._handle_key = None
        self._keepalive: List[Any] = []

    # Mitsuba-like accessors used by the reference's integrators
    def sensors(self):
        return self.sensors_

    def integrator(self):
        return self.integrator_

    def shapes(self):
        return self.shapes_

    def emitters(self):
        return self.emitters_

    # -- flattening ----------------------------------------------------------------------
    def describe(self, sensor) -> _abi.TrSceneDesc:
        """Build the ctypes ``TrSceneDesc`` for ``sensor`` (arrays are kept alive on ``self``)."""
        film = sensor.film()
        prims = [s.prims for s in self.shapes_]
        prim_verts = np.ascontiguousarray(np.concatenate(prims, axis=0) if prims else np.zeros((0, 4, 3)),
                                          dtype=np.float32)
        prim_types = np.ascontiguousarray(np.concatenate([s.prim_types for s in self.shapes_])
                                          if prims else np.zeros(0), dtype=np.uint8)
        shapes = (_abi.TrShapeDesc * max(1, len(self.shapes_)))()
        begin = 0
        for i, s in enumerate(self.shapes_):
            d = shapes[i]
            d.type = _abi.TR_SHAPE_RECTANGLE if s.kind == "rectangle" else _abi.TR_SHAPE_MESH
            d.bsdf, d.emitter = s.bsdf_index, s.emitter_index
            d.flags = _abi.TR_SHAPE_FLAG_SENSOR if s.is_sensor else 0
            d.interior_medium, d.exterior_medium = s.interior_medium, s.exterior_medium
            d.prim_begin, d.prim_count = begin, len(s.prims)
            d.to_world[:] = s.to_world.affine3x4().reshape(-1).tolist()
            begin += len(s.prims)
        bsdfs = (_abi.TrBsdfDesc * max(1, len(self.bsdfs_)))()
        for i, b in enumerate(self.bsdfs_):
            bsdfs[i].type, bsdfs[i].flags = b["type"], b["flags"]
            bsdfs[i].reflectance[:] = [float(x) for x in _variant.colour(b["reflectance"])]
            if "eta" in b:
                bsdfs[i].eta[:] = [float(x) for x in _variant.colour(b["eta"])]
            if "k" in b:
                bsdfs[i].k[:] = [float(x) for x in _variant.colour(b["k"])]
            if "specular_transmittance" in b:
                bsdfs[i].specular_transmittance[:] = [float(x) for x in _variant.colour(b["specular_transmittance"])]
        emitters = (_abi.TrEmitterDesc * max(1, len(self.emitters_)))()
        for i, e in enumerate(self.emitters_):
            emitters[i].type, emitters[i].shape = e["type"], e["shape"]
            emitters[i].radiance[:] = [float(x) for x in _variant.colour(e["radiance"])]
            for k in ("cutoff_angle", "beam_width", "cos_cutoff_angle", "cos_beam_width", "inv_transition_width"):
                setattr(emitters[i], k, e.get(k, 0.0))
            spec = e.get("spec")
            if spec is not None:      # projector: read the CURRENT state (nlos.focus_* may have re-aimed it)
                emitters[i].radiance[:] = [float(x) for x in
                                           _variant.colour(_rgb(spec.get("irradiance"), 1.0) * float(spec.get("scale", 1.0)))]
                emitters[i].fov_x = float(parse_fov(float(spec.get("fov", 45.0)), spec.get("fov_axis", "x"), 1.0))
                tw = _to_transform(spec.get("to_world"))
                emitters[i].to_world[:] = tw.affine3x4().reshape(-1).tolist()
                emitters[i].world_to_local[:] = tw.inverse().affine3x4().reshape(-1).tolist()
                c2s = perspective_projection((1, 1), (1, 1), (0, 0), float(emitters[i].fov_x), 1e-4, 1e4)
                emitters[i].camera_to_sample[:] = c2s.matrix.astype(np.float32).reshape(-1).tolist()
        media = (_abi.TrMediumDesc * max(1, len(self.media_)))()
        for i, m in enumerate(self.media_):
            media[i].sigma_t[:] = [float(x) for x in _variant.colour(m["sigma_t"])]
            media[i].albedo[:] = [float(x) for x in _variant.colour(m["albedo"])]
            media[i].g, media[i].scale = float(m["g"]), float(m.get("scale", 1.0))

        desc = _abi.TrSceneDesc()
        desc.abi_version = _abi.TR_ABI_VERSION
        desc.n_prims = prim_verts.shape[0]
        desc.prim_verts = prim_verts.ctypes.data_as(C.POINTER(C.c_float))
        desc.prim_types = prim_types.ctypes.data_as(C.POINTER(C.c_uint8))
        desc.n_shapes, desc.shapes = len(self.shapes_), shapes
        desc.n_bsdfs, desc.bsdfs = len(self.bsdfs_), bsdfs
        desc.n_emitters, desc.emitters = len(self.emitters_), emitters
        desc.n_media, desc.media = len(self.media_), media
        desc.sensor = sensor.describe(self)
        desc.film = film.describe()
        self._keepalive = [prim_verts, prim_types, shapes, bsdfs, emitters, media, film]
        return desc

    def description_key(self, sensor):
        f = sensor.film()
        em = tuple(_to_transform(e["spec"].get("to_world")).matrix.tobytes() for e in self.emitters_ if e.get("spec"))
        return (id(sensor), f.size(), f.crop_size(), f.crop_offset(), f.temporal_bins, float(f.start_opl), float(f.bin_width_opl),
                sensor.state_key(), em, _variant.is_monochromatic(), f.state_key())

    def __repr__(self):
        return (f"Scene[shapes={len(self.shapes_)}, primitives={sum(len(s.prims) for s in self.shapes_)}, "
                f"bsdfs={len(self.bsdfs_)}, emitters={len(self.emitters_)}, sensors={len(self.sensors_)}]")


class PerspectiveSensor:
    """Mitsuba ``perspective`` sensor (pinhole): parameters only; rays are generated on the device."""

    def __init__(self, props: Properties, film, sampler):
        self.to_world = _to_transform(props.get("to_world"))
        self.fov = float(props.get("fov", 45.0))
        self.fov_axis = props.get("fov_axis", "x")
        self.near_clip = float(props.get("near_clip", 1e-2))
        self.far_clip = float(props.get("far_clip", 1e4))
        props.get("focus_distance")
        self._film, self._sampler = film, sampler
        self.medium = -1

    def film(self):
        return self._film

    def sampler(self):
        return self._sampler

    def world_transform(self):
        return self.to_world

    def state_key(self):
        return ("perspective", self.to_world.matrix.tobytes(), self.fov, self.fov_axis, self.near_clip,
                self.far_clip)

    def describe(self, scene) -> _abi.TrSensorDesc:
        w, h = self._film.size()
        d = _abi.TrSensorDesc()
        d.type, d.shape = _abi.TR_SENSOR_PERSPECTIVE, -1
        d.to_world[:] = self.to_world.affine3x4().reshape(-1).tolist()
        fov_x = parse_fov(self.fov, self.fov_axis, w / h)
        cam_to_sample = perspective_projection((w, h), self._film.crop_size(), self._film.crop_offset(), fov_x,
                                               self.near_clip, self.far_clip)
        d.sample_to_camera[:] = cam_to_sample.inverse().matrix.astype(np.float32).reshape(-1).tolist()
        d.near_clip, d.far_clip = self.near_clip, self.far_clip
        d.medium = self.medium
        return d


class IndependentSampler:
    def __init__(self, props: Properties):
        self._sample_count = int(props.get("sample_count", 4))
        self.base_seed = int(props.get("seed", 0))

    def sample_count(self):
        return self._sample_count

    def set_sample_count(self, spp):
        self._sample_count = int(spp)

    def clone(self):
        c = IndependentSampler(Properties())
        c._sample_count, c.base_seed = self._sample_count, self.base_seed
        return c


# ------------------------------------------------------------------------------------ load_dict
def load_dict(d: Dict[str, Any], base_dir: Optional[str] = None) -> Any:
    """``mi.load_dict`` for the supported subset.  Returns a :class:`Scene` for ``type == 'scene'`` and the
    plugin object for a film / integrator / sensor dictionary."""
    from . import plugins  # late import: plugins registers the mitransient classes

    t = d.get("type")
    if t != "scene":
        if t in plugins.INTEGRATORS or t in plugins.FILMS or t in plugins.SENSORS:
            return plugins.instantiate(d)
        if t in _SHAPE_TYPES or t in _BSDF_TYPES or t in ("projector", "area", "angulararea", "homogeneous", "perspective"):
            return PluginSpec(d)
        raise ValueError(f"load_dict(): unsupported plugin type {t!r}")

    scene = Scene()
    named_bsdf: Dict[str, int] = {}
    named_medium: Dict[str, int] = {}

    def add_bsdf(spec) -> int:
        if isinstance(spec, dict) and spec.get("type") == "ref":
            rid = spec["id"]
            if rid not in named_bsdf:
                if rid in d and isinstance(d[rid], dict):
                    named_bsdf[rid] = add_bsdf(d[rid])
                else:
                    raise ValueError(f"unresolved bsdf reference {rid!r}")
            return named_bsdf[rid]
        flags = 0
        bt = spec.get("type")
        if bt == "twosided":
            _, inner = _find_child(spec, _BSDF_TYPES + ("ref",))
            if inner is None:
                raise ValueError("twosided bsdf without a nested bsdf")
            if inner.get("type") == "ref":
                inner = d[inner["id"]]
            flags |= _abi.TR_BSDF_FLAG_TWOSIDED
            spec, bt = inner, inner.get("type")
        if bt == "diffuse":
            entry = {"type": _abi.TR_BSDF_DIFFUSE, "flags": flags, "reflectance": _rgb(spec.get("reflectance"), 0.5)}
        elif bt == "null":
            entry = {"type": _abi.TR_BSDF_NULL, "flags": flags, "reflectance": ScalarColor3d(0.0)}
        elif bt == "conductor":      # Mitsuba `conductor`: material "none" = perfect mirror (eta 0, k 1), or explicit eta / k
            material = spec.get("material", None)
            if material is not None and str(material).lower() != "none":
                raise ValueError(f"conductor material {material!r}: only 'none' or explicit eta / k are supported "
                                 "(the measured spectral IOR tables are not shipped)")
            has_ior = "eta" in spec or "k" in spec
            entry = {"type": _abi.TR_BSDF_CONDUCTOR, "flags": flags,
                     "reflectance": _rgb(spec.get("specular_reflectance"), 1.0),
                     "eta": _rgb(spec.get("eta"), 0.0) if has_ior else ScalarColor3d(0.0),
                     "k": _rgb(spec.get("k"), 1.0) if has_ior else ScalarColor3d(1.0)}
        elif bt == "dielectric":     # Mitsuba `dielectric`: int_ior "bk7", ext_ior "air" by default
            if flags & _abi.TR_BSDF_FLAG_TWOSIDED:
                raise ValueError("twosided cannot wrap a transmissive bsdf (dielectric)")
            eta = _ior(spec.get("int_ior"), "bk7") / _ior(spec.get("ext_ior"), "air")
            entry = {"type": _abi.TR_BSDF_DIELECTRIC, "flags": flags,
                     "reflectance": _rgb(spec.get("specular_reflectance"), 1.0), "eta": ScalarColor3d(eta),
                     "specular_transmittance": _rgb(spec.get("specular_transmittance"), 1.0)}
        else:
            raise ValueError(f"unsupported bsdf type {bt!r} (supported: diffuse, conductor, dielectric, null, "
                             "twosided(diffuse | conductor))")
        scene.bsdfs_.append(entry)
        return len(scene.bsdfs_) - 1

    def add_medium(spec) -> int:
        if isinstance(spec, dict) and spec.get("type") == "ref":
            rid = spec["id"]
            if rid not in named_medium:
                named_medium[rid] = add_medium(d[rid])
            return named_medium[rid]
        if spec.get("type") != "homogeneous":
            raise ValueError(f"unsupported medium type {spec.get('type')!r} (supported: homogeneous)")
        scale = float(spec.get("scale", 1.0))
        phase = spec.get("phase", {"type": "isotropic"})
        g = float(phase.get("g", 0.0)) if phase.get("type") == "hg" else 0.0
        if phase.get("type") not in ("hg", "isotropic"):
            raise ValueError(f"unsupported phase function {phase.get('type')!r}")
        scene.media_.append({"sigma_t": _rgb(spec.get("sigma_t"), 1.0), "albedo": _rgb(spec.get("albedo"), 0.75),
                             "g": g, "scale": scale})
        return len(scene.media_) - 1

    default_bsdf = None
    # first pass: named bsdfs / media keep their dictionary order (ids are the dict keys)
    for key, val in d.items():
        if isinstance(val, dict) and val.get("type") in _BSDF_TYPES:
            named_bsdf[key] = add_bsdf(val)
        elif isinstance(val, dict) and val.get("type") == "homogeneous":
            named_medium[key] = add_medium(val)

    for key, val in d.items():
        t = _type_of(val)
        if t is None:
            continue
        if t in _SHAPE_TYPES:
            to_world = _to_transform(val.get("to_world"))
            flip = bool(val.get("flip_normals", False))
            if t == "rectangle":
                base, ptypes = _rectangle_prims()
            elif t == "cube":
                base, ptypes = _cube_prims(as_box=bool(val.get("box_primitive", True)))
            elif t == "trimesh":   # in-memory triangle mesh (no Mitsuba counterpart): vertices (N,3), faces (M,3)
                v = np.asarray(val["vertices"], dtype=np.float64).reshape(-1, 3)
                f = np.asarray(val["faces"], dtype=np.int64).reshape(-1, 3)
                base = v[f]
                ptypes = np.full(len(base), _abi.TR_PRIM_TRIANGLE, dtype=np.uint8)
            else:
                fn = val["filename"]
                if not os.path.isabs(fn) and base_dir:
                    fn = os.path.join(base_dir, fn)
                base = load_ply(fn) if t == "ply" else load_obj(fn)
                ptypes = np.full(len(base), _abi.TR_PRIM_TRIANGLE, dtype=np.uint8)
            if t in ("obj", "ply", "trimesh") and bool(val.get("merge_quads", True)) and len(base) > 1:
                base, ptypes = merge_quads(base)     # coplanar triangle pairs -> parallelogram primitives
                # a closed parallelepiped (six such quads) -> one box primitive, like `cube`; not for emitters (their
                # position sampling follows the mesh's primitive list) and not when normals are flipped
                is_emitter = any(_type_of(cv) in ("area", "angulararea") for cv in val.values())
                if bool(val.get("box_primitive", True)) and not flip and not is_emitter:
                    as_box = merge_box(base, ptypes)
                    if as_box is not None:
                        base, ptypes = as_box
            if base.shape[1] == 3:
                base = _with_p3(base, ptypes)
            for ck, cv in val.items():      # every nested plugin of a shape must be one this build understands
                ct = _type_of(cv)
                if ct is not None and ck not in ("interior", "exterior") and ct not in _BSDF_TYPES + (
                        "ref", "area", "angulararea", "homogeneous") + tuple(plugins.SENSORS):
                    raise ValueError(f"shape {key!r}: unsupported nested plugin type {ct!r} (supported bsdfs: diffuse, "
                                     "conductor, dielectric, twosided, null; emitters: area; sensors: nlos_capture_meter)")
            _, bspec = _find_child(val, _BSDF_TYPES + ("ref",))
            if bspec is None:
                if default_bsdf is None:
                    default_bsdf = add_bsdf({"type": "diffuse"})
                bidx = default_bsdf
            else:
                bidx = add_bsdf(bspec)
            shape = Shape(key, t, _apply(to_world, base, flip), ptypes, to_world, bidx, flip)
            if "interior" in val:
                shape.interior_medium = add_medium(val["interior"])
            if "exterior" in val:
                shape.exterior_medium = add_medium(val["exterior"])
            scene.shapes_.append(shape)
            sidx = len(scene.shapes_) - 1
            _, em = _find_child(val, ("area", "angulararea"))
            if em is not None:
                entry = {"type": _abi.TR_EMITTER_AREA, "shape": sidx, "radiance": _rgb(em.get("radiance"), 1.0)}
                if em.get("type") == "angulararea":      # mitransient/emitters/angulararea.py:54-69 (degrees -> radians)
                    cutoff = float(em.get("cutoff_angle", 10))
                    beam = float(em.get("beam_width", cutoff))
                    if not cutoff >= beam:
                        raise AssertionError("angulararea: cutoff_angle must be >= beam_width")
                    c_rad, b_rad = cutoff * (math.pi / 180.0), beam * (math.pi / 180.0)       # dr.deg2rad on Python floats
                    entry.update(type=_abi.TR_EMITTER_ANGULARAREA, cutoff_angle=c_rad, beam_width=b_rad,
                                 cos_cutoff_angle=math.cos(c_rad), cos_beam_width=math.cos(b_rad),
                                 inv_transition_width=(math.inf if c_rad - b_rad == 0.0 else 1.0 / (c_rad - b_rad)))
                scene.emitters_.append(entry)
                shape.emitter_index = len(scene.emitters_) - 1
            sk, se = _find_child(val, tuple(plugins.SENSORS))
            if se is not None:
                if isinstance(se, dict):
                    se = plugins.instantiate(se)
                    val[sk] = se          # share the instance with PluginSpec.sensor()
                shape.is_sensor = True
                se.set_shape(sidx, shape)
                scene.sensors_.append(se)
        elif t == "projector":
            spec = val if isinstance(val, PluginSpec) else PluginSpec(val)
            scene.emitters_.append({"type": _abi.TR_EMITTER_PROJECTOR, "shape": -1, "spec": spec,
                                    "radiance": _rgb(spec.get("irradiance"), 1.0)})
        elif t == "perspective":
            _, fspec = _find_child(val, tuple(plugins.FILMS))
            film = fspec if (fspec is not None and not isinstance(fspec, dict)) else plugins.instantiate(
                fspec if fspec is not None else {"type": "transient_hdr_film"})
            _, sspec = _find_child(val, ("independent",))
            sampler = IndependentSampler(Properties(sspec or {}))
            sensor = PerspectiveSensor(Properties(val), film, sampler)
            if "medium" in val:
                sensor.medium = add_medium(val["medium"])
            scene.sensors_.append(sensor)
        elif t in plugins.INTEGRATORS:
            scene.integrator_ = val if not isinstance(val, dict) else plugins.instantiate(val)
        elif t in _BSDF_TYPES or t in ("homogeneous", "scene"):
            pass
        else:
            raise ValueError(f"load_dict(): unsupported plugin type {t!r} for key {key!r}")
    return scene
