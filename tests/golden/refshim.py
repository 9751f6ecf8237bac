"""A NumPy stand-in for the slice of `mitsuba` / `drjit` that mitransient's HOST-SIDE film, block, plugin-parameter
and scene-dictionary code touches, so that those reference functions can be EXECUTED here (Mitsuba itself is not
installable in this image) and their outputs committed as golden vectors (make_reference_golden.py).

What this makes runnable, unmodified, straight from /root/reference/mitransient:
    films/transient_hdr_film.py   TransientHDRFilm.__init__, prepare_transient_, create_block, add_transient_data,
                                  develop_transient_
    render/transient_image_block.py  TransientImageBlock.clear / put / put_ / accum   (plain and exhaustive_scan)
    films/phasor_hdr_film.py      PhasorHDRFilm.__init__ (frequency selection), prepare, add_transient_data,
                                  develop_phasors_
    render/phasor_image_block.py  PhasorImageBlock.put  (its base class mi.ImageBlock.put is the box-filter
                                  accumulate stated below)
    integrators/*.py              the __init__ parameter parsing and TransientADIntegrator.prepare's pass split
    utils.py                      cornell_box()
    unpolarized_visualization.py  tonemap_transient
    emitters/angulararea.py       AngularAreaLight._fallof_curve

What it does NOT provide: ray tracing, BSDFs, emitters, samplers -- the Mitsuba C++ side of the path.  Those parts of
the oracle stay unpinned ("Oracle").

All arrays are float32 / int64 NumPy arrays; arithmetic follows NumPy's float32 rules (Python scalars are weak), which
is what Dr.Jit's Float does.  TEST INFRASTRUCTURE ONLY -- needs /root/reference, so it runs in the build container only.
"""
from __future__ import annotations

import enum
import importlib.util
import os
import sys
import types

import numpy as np

REFERENCE = "/root/reference"


# ------------------------------------------------------------------------------------------ small vectors
class Vec:
    """mi.Point3f / Vector3f / Point3u / Vector2f / Color3f ...: a tuple of NumPy components."""
    caster = staticmethod(lambda c: np.asarray(c, dtype=np.float32))

    def __init__(self, *c):
        if len(c) == 1 and isinstance(c[0], Vec):
            c = c[0].c
        elif len(c) == 1 and isinstance(c[0], (list, tuple)):
            c = tuple(c[0])
        elif len(c) == 1:
            c = (c[0],) * self.dim
        self.c = tuple(self.caster(x) for x in c)

    dim = 3
    x = property(lambda s: s.c[0])
    y = property(lambda s: s.c[1])
    z = property(lambda s: s.c[2])

    def __getitem__(self, i):
        return self.c[i]

    def __len__(self):
        return len(self.c)

    def __iter__(self):
        return iter(self.c)

    def _zip(self, o, f):
        oc = o.c if isinstance(o, Vec) else (o,) * len(self.c)
        out = Vec.__new__(type(self) if not isinstance(o, Vec) or type(o) is type(self) else VecF)
        out.c = tuple(f(a, b) for a, b in zip(self.c, oc))
        return out

    def __add__(self, o): return self._zip(o, lambda a, b: a + b)
    def __radd__(self, o): return self._zip(o, lambda a, b: b + a)
    def __sub__(self, o): return self._zip(o, lambda a, b: a - b)
    def __mul__(self, o): return self._zip(o, lambda a, b: a * b)
    def __rmul__(self, o): return self._zip(o, lambda a, b: b * a)
    def __truediv__(self, o): return self._zip(o, lambda a, b: a / b)
    def __lt__(self, o): return self._zip(o, lambda a, b: a < b)
    def __le__(self, o): return self._zip(o, lambda a, b: a <= b)
    def __gt__(self, o): return self._zip(o, lambda a, b: a > b)
    def __ge__(self, o): return self._zip(o, lambda a, b: a >= b)
    def __eq__(self, o): return self._zip(o, lambda a, b: a == b)
    def __and__(self, o): return self._zip(o, lambda a, b: a & b)
    __hash__ = None

    def __repr__(self):
        return f"{type(self).__name__}{tuple(np.asarray(x).tolist() for x in self.c)}"


class VecF(Vec):
    pass


class Vec2F(Vec):
    dim = 2


class VecI(Vec):       # unsigned / signed integer vectors: float -> int like Dr.Jit's Point3u(Point3f) conversion
    @staticmethod
    def caster(c):
        with np.errstate(invalid="ignore"):          # inf / nan lanes are masked out by the caller (put_: active &= ...)
            return np.asarray(c).astype(np.int64)


class Vec2I(VecI):
    dim = 2


def Float(v=0.0):
    return np.asarray(v, dtype=np.float32)


def UInt32(v=0):
    a = np.asarray(v)
    return (a.astype(np.float64) if a.dtype.kind == "f" else a).astype(np.int64) & 0xFFFFFFFF


def Int32(v=0):
    return np.asarray(v).astype(np.int64)


def Bool(v=True):
    return np.asarray(v, dtype=bool)


class TensorXf:
    def __init__(self, array, shape=None):
        self.array = np.asarray(array, dtype=np.float32).reshape(-1)
        self.shape = tuple(int(s) for s in (shape if shape is not None else np.shape(array)))

    def numpy(self):
        return self.array.reshape(self.shape)


# ------------------------------------------------------------------------------------------ drjit
def make_drjit():
    dr = types.ModuleType("drjit")
    dr.pi = np.pi
    dr.inf = float("inf")
    dr.floor = lambda v: v._zip(0, lambda a, b: np.floor(a)) if isinstance(v, Vec) else np.floor(v)
    dr.fma = lambda a, b, c: a * b + c            # integer index arithmetic in the film code
    dr.all = lambda v: np.logical_and.reduce([np.asarray(x) for x in v.c]) if isinstance(v, Vec) else np.all(v)
    dr.any = lambda v: np.logical_or.reduce([np.asarray(x) for x in v.c]) if isinstance(v, Vec) else np.any(v)
    dr.prod = lambda v: int(np.prod([int(x) for x in (v.c if isinstance(v, Vec) else v)]))
    dr.zeros = lambda t, n=1: np.zeros(int(n), dtype=np.float32)
    dr.isfinite = np.isfinite
    dr.cos = lambda x: np.cos(np.asarray(x, dtype=np.float32))
    dr.sin = lambda x: np.sin(np.asarray(x, dtype=np.float32))
    dr.acos = lambda x: np.arccos(np.asarray(x, dtype=np.float32))
    dr.deg2rad = lambda x: np.float32(x) * np.float32(np.pi / 180.0)
    dr.rad2deg = lambda x: np.float32(x) * np.float32(180.0 / np.pi)
    dr.select = lambda m, a, b: np.where(m, np.asarray(a, dtype=np.float32), np.asarray(b, dtype=np.float32))
    dr.arange = lambda t, n: np.arange(int(n), dtype=np.int64)
    dr.gather = lambda t, src, idx, active=True: np.asarray(src)[np.asarray(idx, dtype=np.int64)]
    dr.mean = lambda t, axis=-1: TensorXf(t.numpy().mean(axis=axis))
    dr.normalize = lambda v: v * (1.0 / np.sqrt(sum(c * c for c in v.c)))
    dr.epsilon = lambda t: 2.0 ** -24
    dr.opaque = lambda t, v: v
    dr.syntax = lambda f=None, **kw: (f if f is not None else (lambda g: g))
    dr.hint = lambda v, **kw: v

    class ReduceOp(enum.Enum):
        Add = 1
    dr.ReduceOp = ReduceOp

    def scatter_reduce(op, target, value, index, active=True):
        """dr.scatter_reduce(ReduceOp.Add, ...): sequential float32 adds in lane order."""
        index = np.asarray(index, dtype=np.int64).reshape(-1)
        value = np.broadcast_to(np.asarray(value, dtype=np.float32), index.shape)
        active = np.broadcast_to(np.asarray(active, dtype=bool), index.shape)
        np.add.at(target, index[active], value[active])
    dr.scatter_reduce = scatter_reduce

    class ADMode(enum.Enum):
        Primal = 0
        Forward = 1
        Backward = 2
    dr.ADMode = ADMode
    dr.scalar = types.SimpleNamespace(ArrayXf=list, ArrayXu=list, ArrayXi=list, Array2f=Vec2F, Array2u=Vec2I, Array3f=VecF)
    return dr


# ------------------------------------------------------------------------------------------ mitsuba
class Properties(dict):
    def get(self, key, default=None):
        return dict.get(self, key, default)

    def get_texture(self, key, default=None):
        return dict.get(self, key, default)


class _BoxFilter:
    def is_box_filter(self):
        return True

    def border_size(self):
        return 0


class Object:
    def __init__(self, *a, **k):
        pass


class Film(Object):
    """mi.Film base: size / crop / rfilter from the properties, flags() empty (no FilmFlags.Alpha)."""
    def __init__(self, props):
        self._size = Vec2I(int(props.get("width", 768)), int(props.get("height", 576)))
        self._crop_size = Vec2I(int(props.get("crop_width", self._size.x)), int(props.get("crop_height", self._size.y)))
        self._crop_offset = Vec2I(int(props.get("crop_offset_x", 0)), int(props.get("crop_offset_y", 0)))
        self._sample_border = bool(props.get("sample_border", False))

    def size(self): return self._size
    def crop_size(self): return self._crop_size
    def crop_offset(self): return self._crop_offset
    def sample_border(self): return self._sample_border
    def rfilter(self): return _BoxFilter()
    def flags(self): return 0
    def traverse(self, cb): pass
    def parameters_changed(self, keys=None): pass


class _SteadyFilm:
    def prepare(self, aovs): return 3
    def clear(self): pass
    def base_channels_count(self): return 3
    def develop(self, raw=False): return None


class ImageBlock(Object):
    """mi.ImageBlock with rfilter=None (box): put(pos, values, active) accumulates `values` at floor(pos)
    [upstream imageblock.cpp, box-filter fast path]; tensor (H, W, C)."""
    def __init__(self, size, offset, channel_count, rfilter=None, border=False, normalize=False, coalesce=False,
                 compensate=False, warn_negative=False, warn_invalid=False):
        self._size, self._offset, self._cc = size, offset, int(channel_count)
        self._tensor = TensorXf(np.zeros(int(size.y) * int(size.x) * self._cc, dtype=np.float32),
                                (int(size.y), int(size.x), self._cc))

    def tensor(self):
        return self._tensor

    def put(self, pos, values, active=True):
        x = np.floor(np.asarray(pos.x, dtype=np.float32)).astype(np.int64) - int(self._offset.x)
        y = np.floor(np.asarray(pos.y, dtype=np.float32)).astype(np.int64) - int(self._offset.y)
        ok = np.asarray(active, dtype=bool) & (x >= 0) & (y >= 0) & (x < int(self._size.x)) & (y < int(self._size.y))
        base = (y * int(self._size.x) + x) * self._cc
        ok = np.broadcast_to(ok, base.shape)
        for k, v in enumerate(values):
            v = np.broadcast_to(np.asarray(v, dtype=np.float32), base.shape)
            np.add.at(self._tensor.array, base[ok] + k, v[ok])


class ADIntegrator(Object):
    """mitsuba.ad.integrators.common.ADIntegrator.__init__ [upstream]: max_depth (default 6, -1 -> 0xffffffff),
    rr_depth (default 5)."""
    def __init__(self, props=Properties()):
        max_depth = props.get("max_depth", 6)
        if max_depth < 0 and max_depth != -1:
            raise Exception('"max_depth" must be set to -1 (infinite) or a value >= 0')
        self.max_depth = max_depth if max_depth != -1 else 0xFFFFFFFF
        self.rr_depth = props.get("rr_depth", 5)
        if self.rr_depth <= 0:
            raise Exception('"rr_depth" must be set to a value greater than zero!')


class LogLevel(enum.Enum):
    Trace = 0; Debug = 100; Info = 200; Warn = 300; Error = 400


class LogError(RuntimeError):
    pass


def Log(level, msg):
    if level == LogLevel.Error:
        raise LogError(msg)


def install(monochromatic: bool = False, transform_cls=None, color_fn=None):
    """Install the stand-ins as `mitsuba`, `drjit`, and `mitransient` (namespace over the reference tree, without
    running its __init__: that file insists on a real Mitsuba).  Returns (mi, dr)."""
    for name in [m for m in sys.modules if m == "mitransient" or m.startswith("mitransient.")
                 or m in ("mitsuba", "drjit") or m.startswith("mitsuba.") or m.startswith("drjit.")]:
        del sys.modules[name]
    dr = make_drjit()
    mi = types.ModuleType("mitsuba")
    mi.__version__ = "3.7.0"
    mi.MI_VERSION_MAJOR, mi.MI_VERSION_MINOR, mi.MI_VERSION_PATCH = 3, 7, 0
    mi.variant = lambda: "llvm_ad_mono" if monochromatic else "llvm_ad_rgb"
    mi.is_monochromatic, mi.is_polarized, mi.is_spectral, mi.is_rgb = monochromatic, False, False, not monochromatic
    mi.Float, mi.UInt32, mi.UInt, mi.Int32, mi.Int, mi.Bool = Float, UInt32, UInt32, Int32, Int32, Bool
    mi.ScalarInt32 = int
    Float.__module__ = "drjit.llvm.ad"                  # utils.get_module(mi.Float) -> the array namespace below
    for n in ("Point3f", "Vector3f", "Normal3f", "Color3f", "Spectrum", "UnpolarizedSpectrum", "ScalarPoint3f", "ScalarVector3f"):
        setattr(mi, n, VecF)
    for n in ("Point2f", "Vector2f", "ScalarPoint2f", "ScalarVector2f"):
        setattr(mi, n, Vec2F)
    for n in ("Point3u", "Vector3u", "Point3i", "Vector3i", "ScalarVector3u", "ScalarPoint3i", "ScalarPoint3u"):
        setattr(mi, n, VecI)
    for n in ("Point2u", "Vector2u", "ScalarVector2u", "ScalarPoint2u", "ScalarPoint2i", "ScalarVector2i"):
        setattr(mi, n, Vec2I)
    mi.TensorXf, mi.Properties, mi.Object, mi.Film, mi.ImageBlock = TensorXf, Properties, Object, Film, ImageBlock
    mi.Emitter = mi.Sensor = mi.SamplingIntegrator = mi.Texture = mi.ReconstructionFilter = Object
    mi.Transform4f = mi.ScalarTransform4f = transform_cls
    mi.ScalarColor3d = mi.ScalarColor3f = color_fn
    mi.unpolarized_spectrum = lambda v: v
    mi.has_flag = lambda flags, f: bool(int(flags) & int(f))
    mi.FilmFlags = types.SimpleNamespace(Alpha=1, Spectral=2, Special=4)
    mi.EmitterFlags = types.SimpleNamespace(Surface=16)
    mi.ParamFlags = types.SimpleNamespace(NonDifferentiable=1, Differentiable=0, Discontinuous=2)
    mi.Log, mi.LogLevel = Log, LogLevel
    mi.load_dict = lambda d: _SteadyFilm() if d.get("type") == "hdrfilm" else d
    for n in ("register_film", "register_integrator", "register_sensor", "register_emitter"):
        setattr(mi, n, lambda name, factory: None)
    mi.is_spectral = False
    ad = types.ModuleType("mitsuba.ad"); integ = types.ModuleType("mitsuba.ad.integrators")
    common = types.ModuleType("mitsuba.ad.integrators.common")
    common.ADIntegrator = ADIntegrator
    common.mis_weight = lambda a, b: a * a / (a * a + b * b)
    ad.integrators, integ.common, mi.ad = integ, common, ad
    util = types.ModuleType("mitsuba.util")
    mi.util = util
    sys.modules.update({"mitsuba": mi, "mitsuba.ad": ad, "mitsuba.ad.integrators": integ,
                        "mitsuba.ad.integrators.common": common, "mitsuba.util": util, "drjit": dr,
                        "drjit.scalar": types.ModuleType("drjit.scalar")})
    sys.modules["drjit.scalar"].Array3f = VecF
    llvm = types.ModuleType("drjit.llvm"); llvm_ad = types.ModuleType("drjit.llvm.ad")
    llvm_ad.ArrayXf = llvm_ad.ArrayXu = llvm_ad.ArrayXi = list
    llvm_ad.Array2f, llvm_ad.Array2u, llvm_ad.Array3f = Vec2F, Vec2I, VecF
    llvm.ad, dr.llvm = llvm_ad, llvm
    sys.modules.update({"drjit.llvm": llvm, "drjit.llvm.ad": llvm_ad})
    pkg = types.ModuleType("mitransient")
    pkg.__path__ = [os.path.join(REFERENCE, "mitransient")]
    sys.modules["mitransient"] = pkg
    for sub in ("render", "films", "integrators", "sensors", "emitters"):       # skip the packages' plugin-discovery __init__
        m = types.ModuleType(f"mitransient.{sub}")
        m.__path__ = [os.path.join(REFERENCE, "mitransient", sub)]
        sys.modules[f"mitransient.{sub}"] = m
        setattr(pkg, sub, m)
    return mi, dr


def ref_module(dotted: str):
    """Import a module of the reference (e.g. 'mitransient.films.transient_hdr_film') from where it lies."""
    import importlib
    return importlib.import_module(dotted)
