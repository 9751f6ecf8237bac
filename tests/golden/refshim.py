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


# ==========================================================================================================
#  Part 2 -- scene services: enough of mi.Scene / SurfaceInteraction3f / BSDF / Emitter / Sampler / Ray3f for
#  the reference's INTEGRATOR code (integrators/common.py render(), transientpath.py sample()) to run lane-wise.
#  The Mitsuba primitives behind these objects (ray casting, emitter / BSDF sampling, sensor rays) are the CPU
#  oracle's own functions, called through its osv_* exports -- so executing the reference integrator over this
#  module checks OUR READING OF MITRANSIENT'S LOOP (draw order, masks, MIS, roulette, depth and OPL bookkeeping,
#  deposit conditions) against the reference code itself; the primitives stay the oracle's (unpinned) restatement
#  of published Mitsuba algorithms.
# ==========================================================================================================
import ctypes as C

LANES = None          # wavefront width while an integrator runs (set by ShimADIntegrator.sample_rays)


class DrArray(np.ndarray):
    """A Dr.Jit-flavoured ndarray: x[mask] is a full-width masked view (so `x[mask] += 1`, `x[mask] = y` work lane-wise)
    and float & mask selects (dr.rcp(p) & (p > 0))."""
    def __new__(cls, a):
        return np.asarray(a).view(cls)

    @staticmethod
    def _is_mask(k):
        return isinstance(k, np.ndarray) and k.dtype == bool

    def __getitem__(self, k):
        if DrArray._is_mask(k):
            return np.array(self).view(DrArray)
        return np.ndarray.__getitem__(self, k)

    def __setitem__(self, k, v):
        if DrArray._is_mask(k):
            np.copyto(self, np.asarray(v, dtype=self.dtype), where=np.asarray(k))
        else:
            np.ndarray.__setitem__(self, k, v)

    def __and__(self, o):
        if self.dtype.kind == "f":
            return np.where(np.asarray(o, dtype=bool), np.asarray(self), np.zeros((), self.dtype)).view(DrArray)
        return np.ndarray.__and__(self, o)


def _lanes(v, dtype):
    a = np.array(v, dtype=dtype, copy=True)
    if LANES is not None and a.ndim == 0:
        a = np.full(LANES, a, dtype=dtype)
    return a.view(DrArray)


def LFloat(v=0.0): return _lanes(v, np.float32)
def LUInt(v=0): return _lanes(v, np.int64)
def LBool(v=True): return _lanes(v, bool)


class LVec(Vec):
    """Vec with lane-wide components and Dr.Jit masked assignment."""
    caster = staticmethod(lambda c: np.array(c, dtype=np.float32, copy=True).view(DrArray))

    x = property(lambda s: s.c[0], lambda s, v: setattr(s, "c", (np.asarray(v, dtype=np.float32).view(DrArray),) + s.c[1:]))

    def __getitem__(self, k):
        if DrArray._is_mask(k):
            out = type(self).__new__(type(self)); out.c = tuple(np.array(x).view(DrArray) for x in self.c); return out
        return self.c[k]

    def __setitem__(self, k, v):
        vc = v.c if isinstance(v, Vec) else (v,) * len(self.c)
        self.c = tuple(np.where(np.asarray(k), np.asarray(b, dtype=np.float32), np.asarray(a)).view(DrArray)
                       for a, b in zip(self.c, vc))

    def __neg__(self):
        out = type(self).__new__(type(self)); out.c = tuple((-x).view(DrArray) for x in self.c); return out

    def _zip(self, o, f):
        oc = o.c if isinstance(o, Vec) else (o,) * len(self.c)
        out = type(self).__new__(type(self))
        out.c = tuple(np.asarray(f(a, b)).view(DrArray) for a, b in zip(self.c, oc))
        return out

    def arr(self, n=None):                                 # (n, dim) float32, contiguous: what the osv_* calls take
        return np.ascontiguousarray(np.stack([np.broadcast_to(np.asarray(x, dtype=np.float32), (n or LANES,)) for x in self.c], axis=1))

    @classmethod
    def from_arr(cls, a):
        out = cls.__new__(cls); out.c = tuple(np.ascontiguousarray(a[:, k]).view(DrArray) for k in range(a.shape[1])); return out


class LVec2(LVec):
    dim = 2


class Osv:
    """ctypes face of the oracle's scene services (oracle/transient_oracle.cpp, osv_*)."""
    def __init__(self, desc):
        from oracle import oracle
        self.lib = oracle.load()
        self.lib.osv_open.restype = C.c_void_p
        self.lib.osv_open.argtypes = [C.c_void_p]
        self.lib.osv_shape_area.restype = C.c_float
        self.lib.osv_n_shapes.restype = C.c_uint32
        self.lib.osv_sensor_shape.restype = C.c_int32
        self.h = self.lib.osv_open(C.byref(desc))
        if not self.h:
            raise RuntimeError("osv_open failed: " + self.lib.oracle_last_error().decode())
        self._desc = desc

    def call(self, name, *args, n=None):
        fn = getattr(self.lib, name)
        fn.restype = None
        conv = [C.c_void_p(self.h), C.c_uint64(LANES if n is None else n)]
        for a in args:
            conv.append(C.c_void_p(a.ctypes.data) if isinstance(a, np.ndarray) else C.c_int(int(a)))
        fn(*conv)


def _f(n, k=None): return np.zeros((n,) if k is None else (n, k), dtype=np.float32)
def _mask(m, n=None): return np.ascontiguousarray(np.broadcast_to(np.asarray(m, dtype=bool), (n or LANES,))).astype(np.uint8)
def _flt(a, n=None): return np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float32), (n or LANES,)))
def _i32(a, n=None): return np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.int32), (n or LANES,)))


def width(*vals):
    """Dr.Jit width of a set of values: 1 when everything is scalar (then results stay scalar and broadcast later)."""
    w = 1
    for v in vals:
        for c in (v.c if isinstance(v, Vec) else (v,)):
            a = np.asarray(c)
            if a.ndim > 0:
                w = max(w, a.shape[0])
    return w


def _out_vec(a, n):
    v = LVec.from_arr(a)
    if n == 1:
        v.c = tuple(np.asarray(x[0]).view(DrArray) for x in v.c)
    return v


def _out(a, n):
    return (np.asarray(a[0]) if n == 1 else a).view(DrArray)


class Ray3f:
    def __init__(self, o=None, d=None, maxt=None, wavelengths=None):
        time = 0.0
        if isinstance(o, Ray3f):
            o, d, maxt = o.o, o.d, o.maxt
        elif maxt is not None and wavelengths is not None:        # RayDifferential3f(o, d, time, wavelengths)
            time, maxt = maxt, None
        self.o, self.d = LVec(o), LVec(d)
        self.maxt = np.float32(np.finfo(np.float32).max) if maxt is None else maxt
        self.time, self.wavelengths = (LFloat(0.0) if LANES else np.float32(0.0)), None


class BSDFPtr:
    Smooth, Delta, Reflection = 0x1, 0x2, 0x4      # stand-ins for the BSDFFlags unions the integrators test

    def eval(self, ctx, si, wo, active=True):
        return self.eval_pdf(ctx, si, wo, active)[0]

    def __init__(self, osv, shape): self.osv, self.shape = osv, shape

    def flags(self):
        sm, nl = np.zeros(LANES, np.uint8), np.zeros(LANES, np.uint8)
        self.osv.call("osv_bsdf_smooth", _i32(self.shape), sm, nl)
        return np.where(sm != 0, BSDFPtr.Smooth, BSDFPtr.Delta).astype(np.int64).view(DrArray)

    def eval_pdf(self, ctx, si, wo, active=True):
        val, pdf = _f(LANES, 3), _f(LANES)
        self.osv.call("osv_bsdf_eval_pdf", _i32(self.shape), si.wi.arr(), wo.arr(), _mask(active), val, pdf)
        return LVec.from_arr(val), pdf.view(DrArray)

    def sample(self, ctx, si, sample1, sample2, active=True):
        wo, pdf, eta, delta, w = _f(LANES, 3), _f(LANES), _f(LANES), np.zeros(LANES, np.uint8), _f(LANES, 3)
        self.osv.call("osv_bsdf_sample", _i32(self.shape), si.wi.arr(), _flt(sample1), sample2.arr(), _mask(active), wo, pdf, eta, delta, w)
        bs = types.SimpleNamespace(wo=LVec.from_arr(wo), pdf=pdf.view(DrArray), eta=eta.view(DrArray),
                                   sampled_type=np.where(delta != 0, BSDFPtr.Delta, BSDFPtr.Smooth).astype(np.int64).view(DrArray))
        return bs, LVec.from_arr(w)


class EmitterPtr:
    def __init__(self, osv, shape): self.osv, self.shape = osv, shape

    def eval(self, si, active=True):
        rgb = _f(LANES, 3)
        self.osv.call("osv_emitter_eval", _i32(self.shape), si.wi.arr(), _mask(active), rgb)
        return LVec.from_arr(rgb)


class SurfaceInteraction3f:
    wavelengths = None

    def is_valid(self): return self.valid
    def bsdf(self, ray=None): return BSDFPtr(self.osv, self.shape)
    def emitter(self, scene=None): return EmitterPtr(self.osv, self.shape)

    def spawn_ray(self, d):
        o, maxt = _f(LANES, 3), _f(LANES)
        self.osv.call("osv_spawn_ray", self.p.arr(), self.n.arr(), d.arr(), o, maxt)
        return Ray3f(LVec.from_arr(o), d, maxt.view(DrArray))

    def spawn_ray_to(self, target):
        o, d, maxt = _f(LANES, 3), _f(LANES, 3), _f(LANES)
        self.osv.call("osv_spawn_ray_to", self.p.arr(), self.n.arr(), target.arr(), o, d, maxt)
        return Ray3f(LVec.from_arr(o), LVec.from_arr(d), maxt.view(DrArray))

    def to_world(self, v):
        out = _f(LANES, 3); self.osv.call("osv_to_world", self.s.arr(), self.tt.arr(), self.n.arr(), v.arr(), out); return LVec.from_arr(out)

    def to_local(self, v):
        out = _f(LANES, 3); self.osv.call("osv_to_local", self.s.arr(), self.tt.arr(), self.n.arr(), v.arr(), out); return LVec.from_arr(out)

    def to_world_mueller(self, m, a, b): return m       # unpolarized: identity


def _zero_si(osv):
    si = SurfaceInteraction3f(); si.osv = osv
    si.valid, si.t, si.shape = LBool(False), LFloat(np.inf), np.full(LANES, -1, np.int32)
    si.p = si.n = si.s = si.tt = si.wi = LVec(0.0)
    return si


class DirectionSample3f:
    def __init__(self, scene=None, si=None, ref=None):
        if scene is None:
            return
        d, dist = _f(LANES, 3), _f(LANES)
        scene.osv.call("osv_direction_to", si.p.arr(), ref.p.arr(), d, dist)
        self.p, self.n, self.d, self.dist = si.p, si.n, LVec.from_arr(d), dist.view(DrArray)
        self.pdf, self.delta, self.emitter = LFloat(0.0), LBool(False), si.emitter(scene)
        self._shape = si.shape


class ShimScene:
    """mi.Scene over the oracle's scene services."""
    def __init__(self, our_scene, sensor):
        self.our_scene, self.our_sensor = our_scene, sensor
        self.osv = Osv(our_scene.describe(sensor))

    def ray_intersect(self, ray, ray_flags=None, coherent=None, active=True):
        n = width(ray.o, ray.d, ray.maxt)
        valid, t, p, nr, s, tt, shape, wi = np.zeros(n, np.uint8), _f(n), _f(n, 3), _f(n, 3), _f(n, 3), _f(n, 3), np.zeros(n, np.int32), _f(n, 3)
        self.osv.call("osv_ray_intersect", ray.o.arr(n), ray.d.arr(n), _flt(ray.maxt, n), valid, t, p, nr, s, tt, shape, wi, n=n)
        si = SurfaceInteraction3f(); si.osv = self.osv
        si.valid, si.t, si.shape = _out(valid != 0, n), _out(t, n), shape
        si.p, si.n, si.s, si.tt, si.wi = (_out_vec(x, n) for x in (p, nr, s, tt, wi))
        si.time = LFloat(0.0) if n != 1 else np.float32(0.0)
        return si

    def ray_test(self, ray, active=True):
        n = width(ray.o, ray.d, ray.maxt, active)
        hit = np.zeros(n, np.uint8)
        self.osv.call("osv_ray_test", ray.o.arr(n), ray.d.arr(n), _flt(ray.maxt, n), hit, n=n)
        return _out((hit != 0) & np.broadcast_to(np.asarray(active, dtype=bool), (n,)), n)

    def emitters(self): return [ProjectorShim(self)] if self.our_scene.emitters()[0].get("spec") is not None else []
    def shapes(self): return [ShapeShim(self, i) for i in range(self.osv.lib.osv_n_shapes(C.c_void_p(self.osv.h)))]
    def shapes_dr(self): return self

    def sample_emitter_direction(self, ref, sample, test_visibility=True, active=True):
        n = LANES
        dp, dn, dd, dist, pdf, w = _f(n, 3), _f(n, 3), _f(n, 3), _f(n), _f(n), _f(n, 3)
        self.osv.call("osv_sample_emitter_direction", ref.p.arr(), ref.n.arr(), sample.arr(), 1 if test_visibility else 0, _mask(active),
                      dp, dn, dd, dist, pdf, w)
        ds = DirectionSample3f()
        ds.p, ds.n, ds.d = LVec.from_arr(dp), LVec.from_arr(dn), LVec.from_arr(dd)
        ds.dist, ds.pdf, ds.delta, ds.emitter = dist.view(DrArray), pdf.view(DrArray), LBool(False), None
        return ds, LVec.from_arr(w)

    def pdf_emitter_direction(self, ref, ds, active=True):
        pdf = _f(LANES)
        self.osv.call("osv_pdf_emitter_direction", _i32(ds._shape), ds.d.arr(), ds.n.arr(), _flt(ds.dist), _mask(active), pdf)
        return pdf.view(DrArray)


class ShimSampler:
    """Mitsuba's `independent` sampler [upstream]: one PCG32 stream per lane, seeded through TEA from (seed, lane)."""
    MULT = np.uint64(0x5851F42D4C957F2D)

    def __init__(self, spp=4, base_seed=0): self.spp, self.base_seed = spp, base_seed
    def clone(self): return ShimSampler(self.spp, self.base_seed)
    def set_sample_count(self, n): self.spp = int(n)
    def sample_count(self): return self.spp
    def set_samples_per_wavefront(self, n): pass
    def wavefront_size(self): return len(self.state)

    def seed(self, seed, wavefront_size):
        idx = np.arange(int(wavefront_size), dtype=np.uint64)
        v0 = np.full(idx.shape, (int(seed) + self.base_seed) & 0xFFFFFFFF, dtype=np.uint64); v1 = idx.copy()
        s, M = np.uint64(0), np.uint64(0xFFFFFFFF)
        for _ in range(4):                                   # sample_tea_32
            s = (s + np.uint64(0x9E3779B9)) & M
            v0 = (v0 + ((((v1 << np.uint64(4)) & M) + np.uint64(0xA341316C)) ^ ((v1 + s) & M) ^ ((v1 >> np.uint64(5)) + np.uint64(0xC8013EA4)))) & M
            v1 = (v1 + ((((v0 << np.uint64(4)) & M) + np.uint64(0xAD90777D)) ^ ((v0 + s) & M) ^ ((v0 >> np.uint64(5)) + np.uint64(0x7E95761E)))) & M
        with np.errstate(over="ignore"):
            self.inc = (v1 << np.uint64(1)) | np.uint64(1)
            self.state = np.zeros_like(idx)
            self.state = self.state * self.MULT + self.inc
            self.state = self.state + v0
            self.state = self.state * self.MULT + self.inc

    def next_1d(self, active=True):
        old = self.state
        with np.errstate(over="ignore"):
            new = old * self.MULT + self.inc
        self.state = np.where(np.broadcast_to(np.asarray(active, dtype=bool), old.shape), new, old)
        xorshifted = (((old >> np.uint64(18)) ^ old) >> np.uint64(27)) & np.uint64(0xFFFFFFFF)
        rot = (old >> np.uint64(59)).astype(np.uint64)
        out = ((xorshifted >> rot) | (xorshifted << ((np.uint64(32) - rot) & np.uint64(31)))) & np.uint64(0xFFFFFFFF)
        bits = ((out >> np.uint64(9)) | np.uint64(0x3F800000)).astype(np.uint32)
        return (bits.view(np.float32) - np.float32(1.0)).view(DrArray)

    def next_2d(self, active=True):
        x = self.next_1d(active); y = self.next_1d(active)
        out = LVec2.__new__(LVec2); out.c = (x, y); return out


class ShimSteadyBlock:
    def __init__(self, film): self.film = film; self.offset = film.offset; self.data = np.zeros((film.h, film.w, 4), dtype=np.float32)
    def coalesce(self): return False
    def set_coalesce(self, v): pass


class ShimSteadyFilm:
    """hdrfilm + ImageBlock (box): put accumulates [R, G, B, weight] at floor(pos); develop = colour / weight."""
    def __init__(self, w, h, offset=(0, 0)): self.w, self.h, self.offset, self.acc = w, h, offset, np.zeros((h, w, 4), dtype=np.float32)
    def prepare(self, aovs): self.acc[...] = 0; return 3
    def clear(self): self.acc[...] = 0
    def base_channels_count(self): return 3
    def create_block(self): return ShimSteadyBlock(self)
    def put_block(self, block): self.acc += block.data

    def develop(self, raw=False):
        w = self.acc[..., 3:4]
        return self.acc[..., :3] / np.where(w == 0, 1, w)


def _steady_film_from_dict(d):
    """mi.load_dict({'type': 'hdrfilm', ...}) as TransientHDRFilm.prepare issues it (transient_hdr_film.py:131-144): sized by
    the crop window."""
    return ShimSteadyFilm(int(d.get("crop_width", d["width"])), int(d.get("crop_height", d["height"])),
                          (int(d.get("crop_offset_x", 0)), int(d.get("crop_offset_y", 0))))


class ShimSensor:
    def __init__(self, film, spp): self._film, self._sampler = film, ShimSampler(spp)
    def film(self): return self._film
    def sampler(self): return self._sampler


class ShimADIntegrator(ADIntegrator):
    """mitsuba.ad.integrators.common.ADIntegrator [upstream]: __init__ (max_depth, rr_depth, hide_emitters),
    sample_rays (one lane per sample: idx // spp = pixel, pos + next_2d jitter, sensor.sample_ray_differential) and
    _splat_to_block (box filter: accumulate value and weight at the pixel)."""
    def __init__(self, props=Properties()):
        super().__init__(props)
        self.hide_emitters = props.get("hide_emitters", False)

    def aov_names(self): return []

    def sample_rays(self, scene, sensor, sampler, reparam=None):
        global LANES
        film = sensor.film()                                   # film_size = film.crop_size(); pos += film.crop_offset()
        w, h = int(film.crop_size().x), int(film.crop_size().y)
        ox, oy = int(film.crop_offset().x), int(film.crop_offset().y)
        spp = sampler.sample_count()
        LANES = w * h * spp
        idx = np.arange(LANES, dtype=np.int64) // spp
        py, px = idx // w, idx % w
        jitter = sampler.next_2d()
        pos = LVec2.__new__(LVec2); pos.c = ((px + ox).astype(np.float32).view(DrArray), (py + oy).astype(np.float32).view(DrArray))
        if hasattr(sensor, "sample_ray_differential"):        # a reference sensor class (nlos_capture_meter): run ITS code
            scale = LVec2(np.float32(1.0) / np.float32(w), np.float32(1.0) / np.float32(h))       # rcp(film_size)
            pos_adjusted = (pos + jitter) * scale
            ray, weight = sensor.sample_ray_differential(LFloat(0.0), LFloat(0.0), pos_adjusted, LVec2(0.0), LBool(True))
            return ray, LVec(weight), pos
        o, d, maxt = _f(LANES, 3), _f(LANES, 3), _f(LANES)
        scene.osv.call("osv_camera_rays", px.astype(np.uint32), py.astype(np.uint32), jitter.arr(), o, d, maxt)
        ray = Ray3f(LVec.from_arr(o), LVec.from_arr(d), maxt.view(DrArray))
        return ray, LVec(1.0), pos

    @staticmethod
    def _splat_to_block(block, film, pos, value, weight, alpha, aovs, wavelengths):
        x, y = np.asarray(pos.x).astype(np.int64) - block.offset[0], np.asarray(pos.y).astype(np.int64) - block.offset[1]
        for k, v in enumerate((value.x, value.y, value.z, np.full(LANES, weight, dtype=np.float32))):
            np.add.at(block.data[..., k], (y, x), np.asarray(v, dtype=np.float32))


def install_scene_services(monochromatic=False, transform_cls=None, color_fn=None):
    """install() + the lane-wise array types and scene objects of part 2."""
    mi, dr = install(monochromatic, transform_cls, color_fn)
    mi.Float, mi.UInt32, mi.UInt, mi.Int32, mi.Bool, mi.Mask = LFloat, LUInt, LUInt, LUInt, LBool, LBool
    LFloat.__module__ = "drjit.llvm.ad"
    for n in ("Point3f", "Vector3f", "Normal3f", "Color3f", "Spectrum", "UnpolarizedSpectrum"):
        setattr(mi, n, LVec)
    for n in ("Point2f", "Vector2f"):
        setattr(mi, n, LVec2)
    mi.Ray3f, mi.SurfaceInteraction3f, mi.DirectionSample3f = Ray3f, SurfaceInteraction3f, DirectionSample3f
    mi.BSDFContext = lambda *a, **k: None
    mi.RayFlags = types.SimpleNamespace(All=0)
    mi.BSDFFlags = types.SimpleNamespace(Smooth=BSDFPtr.Smooth, Delta=BSDFPtr.Delta)
    mi.has_flag = lambda flags, f: ((np.asarray(flags) & int(f)) != 0).view(DrArray) if isinstance(flags, np.ndarray) else bool(int(flags) & int(f))
    mi.Sampler = ShimSampler
    common = sys.modules["mitsuba.ad.integrators.common"]
    common.ADIntegrator = ShimADIntegrator

    def mis_weight(a, b):
        a2, b2 = np.asarray(a, dtype=np.float32) ** 2, np.asarray(b, dtype=np.float32) ** 2
        with np.errstate(invalid="ignore", divide="ignore"):
            w = a2 / (a2 + b2)
        return np.where(np.isfinite(w), w, np.float32(0)).view(DrArray)
    common.mis_weight = mis_weight

    def select(m, a, b):
        if isinstance(a, Vec) or isinstance(b, Vec):
            ref = a if isinstance(a, Vec) else b
            ac = a.c if isinstance(a, Vec) else (a,) * len(ref.c)
            bc = b.c if isinstance(b, Vec) else (b,) * len(ref.c)
            out = type(ref).__new__(type(ref))
            out.c = tuple(np.where(m, np.asarray(x, dtype=np.float32), np.asarray(y, dtype=np.float32)).view(DrArray) for x, y in zip(ac, bc))
            return out
        return np.where(m, np.asarray(a, dtype=np.float32), np.asarray(b, dtype=np.float32)).view(DrArray)
    dr.select = select
    dr.max = lambda v: np.maximum(np.maximum(v.x, v.y), v.z).view(DrArray) if isinstance(v, Vec) else np.max(v)
    dr.minimum = lambda a, b: np.minimum(a, np.float32(b) if np.isscalar(b) else b).view(DrArray)
    dr.maximum = lambda a, b: np.maximum(a, np.float32(b) if np.isscalar(b) else b).view(DrArray)
    dr.rcp = lambda x: (np.float32(1.0) / np.asarray(x, dtype=np.float32)).view(DrArray)
    dr.detach = lambda x, *a: x
    dr.copy = lambda x: x
    dr.zeros = lambda t, n=1: (_zero_si(_CURRENT["osv"]) if isinstance(t, type) and issubclass(t, SurfaceInteraction3f)
                               else np.zeros(int(n), dtype=np.float32))
    dr.hint = lambda v, **kw: bool(np.any(v)) if isinstance(v, np.ndarray) else v

    class _Ctx:
        def __init__(self, *a, **k): pass
        def __enter__(self): return self
        def __exit__(self, *a): return False
    dr.resume_grad = dr.suspend_grad = _Ctx
    return mi, dr


_CURRENT = {"osv": None}


def reference_render(our_scene, integrator_module, integrator_cls, props, spp, seed=0, monochromatic=False):
    """Run the REFERENCE'S render() (integrators/common.py:122-213) + sample() for `our_scene` (a mitransient_b200
    Scene) and return (steady (H,W,3), transient raw tensor, deposits (N,6) in the order the reference produced them)."""
    import mitransient_b200 as mitr
    global LANES
    LANES = None
    mi, dr = install_scene_services(monochromatic, mitr.ScalarTransform4f, mitr.ScalarColor3d)
    thf = ref_module("mitransient.films.transient_hdr_film")
    mod = ref_module(integrator_module)
    sensor = our_scene.sensors()[0]
    f = sensor.film()
    if hasattr(f, "frequencies"):                     # the reference's PhasorHDRFilm collects the deposits (monochromatic variant)
        dr.cos = lambda x: np.cos(np.asarray(x, dtype=np.float32))
        dr.sin = lambda x: np.sin(np.asarray(x, dtype=np.float32))
        phf = ref_module("mitransient.films.phasor_hdr_film")
        film = phf.PhasorHDRFilm(mi.Properties(width=f.size()[0], height=f.size()[1], temporal_bins=f.temporal_bins, start_opl=f.start_opl,
                                               bin_width_opl=f.bin_width_opl, wl_mean=f.wl_mean, wl_sigma=f.wl_sigma))
        film.develop_transient_ = lambda raw=True: film.develop_phasors_(raw=raw)
    else:
        film = thf.TransientHDRFilm(mi.Properties(width=f.size()[0], height=f.size()[1], temporal_bins=f.temporal_bins,
                                                  start_opl=f.start_opl, bin_width_opl=f.bin_width_opl,
                                                  crop_width=f.crop_size()[0], crop_height=f.crop_size()[1],
                                                  crop_offset_x=f.crop_offset()[0], crop_offset_y=f.crop_offset()[1]))
    log = []
    inner = film.add_transient_data

    def logged(pos, distance, wavelengths, spec, ray_weight, active, **kw):
        a = np.broadcast_to(np.asarray(active, dtype=bool), (LANES,))
        val = spec * ray_weight
        rec = np.stack([np.asarray(pos.x), np.asarray(pos.y), np.asarray(distance), np.asarray(val.x), np.asarray(val.y),
                        np.asarray(val.z)], axis=1).astype(np.float32)
        nz = a & ((rec[:, 3] != 0) | (rec[:, 4] != 0) | (rec[:, 5] != 0))
        log.append(rec[nz])
        return inner(pos, distance, wavelengths, spec, ray_weight, active, **kw)
    film.add_transient_data = logged
    _real_load = mi.load_dict
    mi.load_dict = lambda d: _steady_film_from_dict(d) if d.get("type") == "hdrfilm" else _real_load(d)
    scene = ShimScene(our_scene, sensor)
    _CURRENT["osv"] = scene.osv
    shim_sensor = ShimSensor(film, spp)
    shim_sensor._sampler.base_seed = sensor.sampler().base_seed
    scene.sensors = lambda: [shim_sensor]
    integ = getattr(mod, integrator_cls)(mi.Properties(**props))
    steady, transient = integ.render(scene, shim_sensor, seed=seed, spp=spp)
    LANES = None
    return np.asarray(steady), film.develop_transient_(raw=True).numpy(), np.concatenate(log, axis=0) if log else np.zeros((0, 6), np.float32)


# ---------------------------------------------------------------------------------------------------------
#  Part 3 -- what transient_nlos_path / nlos_capture_meter need on top: shapes, the projector, transforms,
#  DiscreteDistribution, mi.traverse, meshgrid / linspace / gather, and dr.syntax's array-`if`.
# ---------------------------------------------------------------------------------------------------------
import ast
import copy as _copy
import inspect
import textwrap


class ShimTransform4f:
    """mi.Transform4f as the reference uses it: translate(), translation(), `@ Vector3f`; float32 entries."""
    def __init__(self, m=None): self.m = np.eye(4, dtype=np.float64) if m is None else np.asarray(m, dtype=np.float64)

    @property
    def matrix(self):                              # world_transform().matrix.to_numpy()[..., 0]  (transientnlospath.py:355)
        return types.SimpleNamespace(to_numpy=lambda: self.m[..., None])

    def translate(self, v):
        t = np.eye(4); t[:3, 3] = [float(np.asarray(c)) for c in (v.c if isinstance(v, Vec) else v)]
        return ShimTransform4f(self.m @ t)

    def translation(self):
        return LVec(*[np.float32(self.m[k, 3]) for k in range(3)])

    def __matmul__(self, v):                       # vectors only (directions): 3x3 block, float32, no contraction
        m = self.m.astype(np.float32)
        return LVec(*[(m[k, 0] * v.x + m[k, 1] * v.y) + m[k, 2] * v.z for k in range(3)])


class ShapeShim:
    def __init__(self, scene, index): self.scene, self.index = scene, index
    def __eq__(self, o): return isinstance(o, ShapeShim) and o.index == self.index
    __hash__ = None
    def surface_area(self): return [float(self.scene.osv.lib.osv_shape_area(C.c_void_p(self.scene.osv.h), C.c_uint32(self.index)))]
    def sample_position(self, time, sample, active=True): return ShapePtrLanes(self.scene, self.index).sample_position(time, sample, active)


class ShapePtrLanes:
    def __init__(self, scene, index): self.scene, self.index = scene, index

    def sample_position(self, time, sample, active=True):
        n = width(sample, self.index, active)
        p, nr, pdf = _f(n, 3), _f(n, 3), _f(n)
        self.scene.osv.call("osv_shape_sample_position", _i32(self.index, n), sample.arr(n), _mask(active, n), p, nr, pdf, n=n)
        return types.SimpleNamespace(p=_out_vec(p, n), n=_out_vec(nr, n), pdf=_out(pdf, n))


class ProjectorShim:
    def __init__(self, scene):
        self.scene = scene
        from mitransient_b200.scene import _to_transform
        self._tw = ShimTransform4f(_to_transform(scene.our_scene.emitters()[0]["spec"].get("to_world")).matrix)

    def world_transform(self): return self._tw

    def sample_direction(self, it, sample, active=True):
        n = LANES
        dp, dd, dist, pdf, w = _f(n, 3), _f(n, 3), _f(n), _f(n), _f(n, 3)
        self.scene.osv.call("osv_projector_sample_direction", it.p.arr(), _mask(active), dp, dd, dist, pdf, w)
        ds = DirectionSample3f()
        ds.p, ds.d, ds.dist, ds.pdf, ds.delta, ds.emitter = LVec.from_arr(dp), LVec.from_arr(dd), dist.view(DrArray), pdf.view(DrArray), LBool(True), None
        ds.n = LVec(0.0)
        return ds, LVec.from_arr(w)


class DiscreteDistribution:
    def __init__(self, pmf): self.pmf = np.asarray(pmf, dtype=np.float32)

    def sample_reuse_pmf(self, u, active=True):
        n = width(u)
        idx, un, pm = np.zeros(n, np.int32), _f(n), _f(n)
        lib = _CURRENT["osv"].lib
        lib.osv_discrete_sample_reuse_pmf.restype = None
        lib.osv_discrete_sample_reuse_pmf(C.c_void_p(_CURRENT["osv"].h), C.c_uint64(n), C.c_void_p(self.pmf.ctypes.data), C.c_uint32(len(self.pmf)),
                                          C.c_void_p(_flt(u, n).ctypes.data), C.c_void_p(idx.ctypes.data), C.c_void_p(un.ctypes.data),
                                          C.c_void_p(pm.ctypes.data))
        return idx, un.view(DrArray), pm.view(DrArray)


class ShimSensorBase(Object):
    """mi.Sensor: film(), sampler(), get_shape() come from the properties the driver passes in."""
    def __init__(self, props):
        self._film, self._sampler, self._shape = props.get("_film"), props.get("_sampler"), props.get("_shape")
    def film(self): return self._film
    def sampler(self): return self._sampler
    def get_shape(self): return self._shape
    def shape(self): return self._shape
    def traverse(self, cb): pass


class _SymbolicLoopState(ast.NodeTransformer):
    """Dr.Jit records the body of a `while` over arrays ONCE (symbolic mode, the default): only the loop's state
    variables -- the names the body assigns -- are carried from one iteration to the next.  An array that lives outside
    the loop, is not assigned in it, but is MUTATED in place by something the body calls (`active_e &= ...` inside
    emitter_laser_targets_sample, called from the exhaustive-capture loop of emitter_laser_sample,
    transientnlospath.py:612-628) enters every recorded iteration with its pre-loop value.  This stand-in runs loops
    iteration by iteration, so it restores such arrays at the top of every iteration to keep that meaning."""
    counter = 0

    def visit_While(self, node):
        self.generic_visit(node)
        def root(t):
            while isinstance(t, (ast.Attribute, ast.Subscript, ast.Starred)):
                t = t.value
            return [t.id] if isinstance(t, ast.Name) else ([r for e in t.elts for r in root(e)] if isinstance(t, (ast.Tuple, ast.List)) else [])
        stored = {n.id for b in node.body for n in ast.walk(b) if isinstance(n, ast.Name) and isinstance(n.ctx, (ast.Store, ast.Del))}
        for b in node.body:                                   # `x.f = ...`, `x[mask] = ...`, `x[mask] += ...` modify x as well
            for n in ast.walk(b):
                for t in (n.targets if isinstance(n, ast.Assign) else [n.target] if isinstance(n, (ast.AugAssign, ast.AnnAssign)) else []):
                    stored.update(root(t))
        loaded = {n.id for b in node.body for n in ast.walk(b) if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load)}
        names = sorted(loaded - stored)
        _SymbolicLoopState.counter += 1
        snap = "__shim_snap%d" % _SymbolicLoopState.counter
        take = ast.Assign(targets=[ast.Name(id=snap, ctx=ast.Store())],
                          value=ast.Call(func=ast.Name(id="__shim_snapshot", ctx=ast.Load()),
                                         args=[ast.Call(func=ast.Name(id="locals", ctx=ast.Load()), args=[], keywords=[]),
                                               ast.Constant(value=tuple(names))], keywords=[]))
        put = ast.Expr(value=ast.Call(func=ast.Name(id="__shim_restore", ctx=ast.Load()), args=[ast.Name(id=snap, ctx=ast.Load())], keywords=[]))
        node.body = [put] + node.body
        return [take, node]


def _shim_snapshot(frame_locals, names):
    saved = []
    for n in names:
        v = frame_locals.get(n)
        for a in (v.c if isinstance(v, Vec) else (v,)):
            if isinstance(a, np.ndarray) and a.ndim > 0 and a.flags.writeable:
                saved.append((a, a.copy()))
    return saved


def _shim_restore(saved):
    for a, c in saved:
        a[...] = c


def _rewrite_array_ifs(fn):
    """dr.syntax for this stand-in: `if <array condition>: <assignments>` becomes lane-wise selects; everything else
    (scalar conditions, loops driven by dr.hint) runs as ordinary Python."""
    src = textwrap.dedent(inspect.getsource(fn))
    tree = ast.parse(src)
    fdef = tree.body[0]
    fdef.decorator_list = []

    class T(ast.NodeTransformer):
        def visit_If(self, node):
            self.generic_visit(node)
            simple = not node.orelse and all(isinstance(b, (ast.Assign, ast.AugAssign)) and
                                             isinstance(b.targets[0] if isinstance(b, ast.Assign) else b.target, ast.Name) for b in node.body)
            if not simple:
                return node
            masked = []
            for b in node.body:
                tgt = b.targets[0] if isinstance(b, ast.Assign) else b.target
                new = b.value if isinstance(b, ast.Assign) else ast.BinOp(left=ast.Name(id=tgt.id, ctx=ast.Load()), op=b.op, right=b.value)
                sel = ast.Call(func=ast.Name(id="__shim_select", ctx=ast.Load()),
                               args=[ast.Name(id="__shim_c", ctx=ast.Load()), new, ast.Name(id=tgt.id, ctx=ast.Load())], keywords=[])
                masked.append(ast.Assign(targets=[ast.Name(id=tgt.id, ctx=ast.Store())], value=sel))
            test = ast.Call(func=ast.Name(id="__shim_is_array", ctx=ast.Load()), args=[ast.Name(id="__shim_c", ctx=ast.Load())], keywords=[])
            return [ast.Assign(targets=[ast.Name(id="__shim_c", ctx=ast.Store())], value=node.test),
                    ast.If(test=test, body=masked, orelse=[ast.If(test=ast.Name(id="__shim_c", ctx=ast.Load()), body=node.body, orelse=[])])]
    tree = ast.fix_missing_locations(_SymbolicLoopState().visit(T().visit(tree)))
    g = fn.__globals__
    g["__shim_snapshot"], g["__shim_restore"] = _shim_snapshot, _shim_restore
    g["__shim_is_array"] = lambda c: isinstance(c, np.ndarray) and c.ndim > 0
    g["__shim_select"] = lambda c, a, b: sys.modules["drjit"].select(c, a, b)
    loc = {}
    exec(compile(tree, inspect.getsourcefile(fn), "exec"), g, loc)
    return loc[fdef.name]


def install_nlos_services(monochromatic=False, transform_cls=None, color_fn=None):
    mi, dr = install_scene_services(monochromatic, transform_cls, color_fn)
    mi.Transform4f, mi.Sensor = ShimTransform4f, ShimSensorBase
    mi.RayDifferential3f = Ray3f
    mi.DiscreteDistribution, mi.ShapePtr, mi.EmitterPtr = DiscreteDistribution, ShapePtrLanes, EmitterPtr
    mi.ScalarVector2f = Vec2F
    mi.Interaction3f = mi.PositionSample3f = object
    mi.BSDFSample3f = "BSDFSample3f"
    mi.BSDFFlags.Reflection = BSDFPtr.Reflection
    mi.Frame3f = types.SimpleNamespace(cos_theta=lambda v: v.z)
    sys.modules["drjit.scalar"].Array3f = VecF
    class _ScanProjector:
        """mi.load_dict({'type': 'projector', 'fov': ..., 'to_world': ...}) of TransientNLOSPath.prepare (:352-356): only its
        sample_ray is used (:366-367); the rays are the oracle's (same widened sample -> camera matrix as the host computes)."""
        def __init__(self, d):
            from mitransient_b200.transforms import parse_fov, perspective_projection
            fov_x = float(parse_fov(float(d["fov"]), "x", 1.0))
            self.s2c = np.ascontiguousarray(perspective_projection((1, 1), (1, 1), (0, 0), fov_x, 1e-4, 1e4).inverse().matrix.astype(np.float32))

        def sample_ray(self, time, sample1, sample2, sample3, active=True):
            n = width(sample3)
            o, dd = _f(n, 3), _f(n, 3)
            _CURRENT["osv"].call("osv_projector_sample_ray", self.s2c, sample3.arr(n), o, dd, n=n)
            return Ray3f(LVec.from_arr(o), LVec.from_arr(dd)), LVec(1.0)
    prev_load = mi.load_dict
    mi.load_dict = lambda d: _ScanProjector(d) if d.get("type") == "projector" else prev_load(d)
    mi.ScalarTransform4f = transform_cls
    mi.traverse = lambda obj: {"film.size": Vec2I(int(obj.film().size().x), int(obj.film().size().y)),
                               "film.laser_scan_width": getattr(obj.film(), "laser_scan_width", 0),
                               "film.laser_scan_height": getattr(obj.film(), "laser_scan_height", 0)}
    dr.syntax = lambda f=None, **kw: (_rewrite_array_ifs(f) if f is not None else (lambda g: _rewrite_array_ifs(g)))
    dr.make_opaque = lambda *a: None
    dr.linspace = lambda t, a, b, n, endpoint=True: (np.arange(int(n), dtype=np.float32) * np.float32((b - a) / (n if not endpoint else n - 1))
                                                      + np.float32(a)).view(DrArray)
    dr.meshgrid = lambda x, y: (np.tile(x, len(y)).view(DrArray), np.repeat(y, len(x)).view(DrArray))
    osv_dot = lambda a, b: _dot(a, b)
    dr.dot = osv_dot
    dr.norm = lambda v: np.sqrt(_dot(v, v)).view(DrArray)
    dr.normalize = lambda v: v * (np.float32(1.0) / np.sqrt(_dot(v, v)))
    dr.square = dr.sqr = lambda x: (np.asarray(x, dtype=np.float32) * np.asarray(x, dtype=np.float32)).view(DrArray)
    dr.abs = lambda x: np.abs(x).view(DrArray)
    dr.copy = lambda x: _copy.copy(x)
    base_zeros = dr.zeros
    dr.zeros = lambda t, n=1: (types.SimpleNamespace(wo=LVec(0.0), pdf=LFloat(0.0), eta=LFloat(1.0), sampled_type=LUInt(0), sampled_component=0)
                               if t == "BSDFSample3f" else base_zeros(t, n))
    base_select = dr.select

    def select(m, a, b):
        if isinstance(a, types.SimpleNamespace) and isinstance(b, types.SimpleNamespace):
            return types.SimpleNamespace(**{k: (select(m, getattr(a, k), getattr(b, k)) if k != "sampled_component" else 0)
                                            for k in ("wo", "pdf", "eta", "sampled_type", "sampled_component")})
        if isinstance(a, np.ndarray) and a.dtype.kind in "iu" and not isinstance(b, Vec):
            return np.where(m, a, b).view(DrArray)
        return base_select(m, a, b)
    dr.select = select

    def gather(t, src, idx, active=True):
        idx = np.asarray(idx, dtype=np.int64)
        if t is ShapePtrLanes:
            return ShapePtrLanes(src, idx.astype(np.int32))
        if isinstance(src, Vec):
            out = LVec.__new__(LVec); out.c = tuple(np.asarray(c)[idx].view(DrArray) for c in src.c); return out
        return np.asarray(src)[idx].view(DrArray)
    dr.gather = gather
    return mi, dr


def _dot(a, b):
    """dr.dot with the arithmetic contract's fused chain (the oracle's dot, through osv_dot)."""
    n = width(a, b)
    out = _f(n)
    _CURRENT["osv"].call("osv_dot", a.arr(n), b.arr(n), out, n=n)
    return _out(out, n)


def reference_render_nlos(our_scene, props, spp, seed=0):
    """Run the REFERENCE'S render() + TransientNLOSPath (prepare, sample, emitter_*_sample, hidden_geometry_sample)
    and its NLOSCaptureMeter for `our_scene`; returns (steady, raw film, deposits)."""
    import mitransient_b200 as mitr
    global LANES
    LANES = None
    mi, dr = install_nlos_services(False, mitr.ScalarTransform4f, mitr.ScalarColor3d)
    thf = ref_module("mitransient.films.transient_hdr_film")
    mod = ref_module("mitransient.integrators.transientnlospath")
    meter = ref_module("mitransient.sensors.nloscapturemeter")
    sensor = our_scene.sensors()[0]
    f = sensor.film()
    scene = None
    film = thf.TransientHDRFilm(mi.Properties(width=f.size()[0], height=f.size()[1], temporal_bins=f.temporal_bins,
                                              start_opl=f.start_opl, bin_width_opl=f.bin_width_opl,
                                              exhaustive_scan=f.exhaustive_scan, laser_scan_width=f.laser_scan_width,
                                              laser_scan_height=f.laser_scan_height))
    log = []
    inner = film.add_transient_data

    def logged(pos, distance, wavelengths, spec, ray_weight, active, **kw):
        a = np.broadcast_to(np.asarray(active, dtype=bool), (LANES,))
        val = spec * ray_weight
        cols = [np.broadcast_to(np.asarray(c, dtype=np.float32), (LANES,)) for c in (pos.x, pos.y, distance, val.x, val.y, val.z)]
        rec = np.stack(cols, axis=1).astype(np.float32)
        nz = a & ((rec[:, 3] != 0) | (rec[:, 4] != 0) | (rec[:, 5] != 0))
        log.append(rec[nz])
        return inner(pos, distance, wavelengths, spec, ray_weight, active, **kw)
    film.add_transient_data = logged
    _real_load = mi.load_dict
    mi.load_dict = lambda d: _steady_film_from_dict(d) if d.get("type") == "hdrfilm" else _real_load(d)
    scene = ShimScene(our_scene, sensor)
    _CURRENT["osv"] = scene.osv
    wall_index = int(scene.osv.lib.osv_sensor_shape(C.c_void_p(scene.osv.h)))
    sampler = ShimSampler(spp); sampler.base_seed = sensor.sampler().base_seed
    if not hasattr(sensor, "is_confocal"):                    # a perspective camera drives transient_nlos_path (nlos-z-simple.xml)
        class _PerspectiveSensor(ShimSensor):
            """mi.Sensor `perspective` [upstream]: sample_ray at a film position = the oracle's camera ray (pixel corner)."""
            def get_shape(self): return None

            def sample_ray(self, time, sample1, sample2, sample3, active=True):
                w, h = int(self._film.size().x), int(self._film.size().y)
                n = width(sample2)
                px = np.rint(np.asarray(sample2.x, dtype=np.float64) * w).astype(np.uint32)
                py = np.rint(np.asarray(sample2.y, dtype=np.float64) * h).astype(np.uint32)
                o, d, maxt = _f(n, 3), _f(n, 3), _f(n)
                scene.osv.call("osv_camera_rays", np.ascontiguousarray(np.broadcast_to(px, (n,))), np.ascontiguousarray(np.broadcast_to(py, (n,))),
                               _f(n, 2), o, d, maxt, n=n)
                return Ray3f(LVec.from_arr(o), LVec.from_arr(d), maxt.view(DrArray)), LVec(1.0)
        ref_sensor = _PerspectiveSensor(film, spp)
        ref_sensor._sampler.base_seed = sensor.sampler().base_seed
        scene.sensors = lambda: [ref_sensor]
        LANES = int(f.size()[0]) * int(f.size()[1])
        integ = mod.TransientNLOSPath(mi.Properties(**props))
        steady, transient = integ.render(scene, ref_sensor, seed=seed, spp=spp)
        LANES = None
        return np.asarray(steady), film.develop_transient_(raw=True).numpy(), np.concatenate(log, axis=0) if log else np.zeros((0, 6), np.float32)
    extra = {}
    if sensor.is_confocal:                                    # 1x1 film aimed at laser_target (nloscapturemeter.py:113-125)
        extra = {"original_film_width": sensor.original_film_width, "original_film_height": sensor.original_film_height}
    ref_sensor = meter.NLOSCaptureMeter(mi.Properties(sensor_origin=VecF(*[float(x) for x in sensor._world_transform.translation()]), _film=film,
                                                      _sampler=sampler, _shape=ShapeShim(scene, wall_index), **extra))
    if sensor.is_confocal:                                    # what mitransient.nlos.focus_emitter_at_relay_wall_* sets (nlos.py:28-34)
        ref_sensor.laser_target = LVec(*[np.float32(x) for x in np.asarray(sensor.laser_target, dtype=np.float64)])
    scene.sensors = lambda: [ref_sensor]
    LANES = int(f.size()[0]) * int(f.size()[1])            # width of the prepare() phase: one lane per film pixel
    integ = mod.TransientNLOSPath(mi.Properties(**props))
    steady, transient = integ.render(scene, ref_sensor, seed=seed, spp=spp)
    LANES = None
    return np.asarray(steady), film.develop_transient_(raw=True).numpy(), np.concatenate(log, axis=0) if log else np.zeros((0, 6), np.float32)


# ---------------------------------------------------------------------------------------------------------
#  Part 4 -- mitransient's own `angulararea` emitter (emitters/angulararea.py) executed inside the scene services:
#  Scene::sample_emitter_direction / pdf_emitter_direction [upstream: emitter choice, visibility test] around the
#  REFERENCE class's sample_direction / pdf_direction / eval; the shape underneath is the oracle's.
# ---------------------------------------------------------------------------------------------------------
class ShimEmitterBase(Object):
    def __init__(self, props): self._shape = props.get("_shape")
    def get_shape(self): return self._shape
    def traverse(self, cb): pass
    def parameters_changed(self, keys=None): pass


class _ConstTexture:
    def __init__(self, rgb): self.rgb = [np.float32(x) for x in rgb]
    def eval(self, si, active=True): return LVec(*self.rgb)


class _EmitterShape:
    """mi.Shape of the emitter: sample_direction / pdf_direction through the oracle."""
    def __init__(self, scene): self.scene = scene

    def sample_direction(self, ref, sample, active=True):
        n = LANES
        dp, dn, dd, dist, pdf = _f(n, 3), _f(n, 3), _f(n, 3), _f(n), _f(n)
        self.scene.osv.call("osv_shape_sample_direction", ref.p.arr(), sample.arr(), _mask(active), dp, dn, dd, dist, pdf)
        ds = DirectionSample3f()
        ds.p, ds.n, ds.d, ds.dist, ds.pdf = LVec.from_arr(dp), LVec.from_arr(dn), LVec.from_arr(dd), dist.view(DrArray), pdf.view(DrArray)
        ds.delta, ds.emitter = LBool(False), None
        return ds

    def pdf_direction(self, ref, ds, active=True):
        pdf = _f(LANES)
        self.scene.osv.call("osv_shape_pdf_direction", ds.d.arr(), ds.n.arr(), _flt(ds.dist), _mask(active), pdf)
        return pdf.view(DrArray)


class _Frame3f:
    def __init__(self, n): self.n = n

    def to_local(self, v):
        out = _f(LANES, 3); _CURRENT["osv"].call("osv_frame_to_local", self.n.arr(), v.arr(), out); return LVec.from_arr(out)
    cos_theta = staticmethod(lambda v: v.z)


def install_angular_services():
    """install_scene_services() + what emitters/angulararea.py needs; returns (mi, dr, AngularAreaLight)."""
    import mitransient_b200 as mitr
    mi, dr = install_scene_services(False, mitr.ScalarTransform4f, mitr.ScalarColor3d)
    mi.Emitter = ShimEmitterBase
    mi.Frame3f = _Frame3f
    mi.EmitterPtr = lambda e: e
    mi.SurfaceInteraction3f = _SiFromDs
    mi.Interaction3f = mi.TraversalCallback = mi.PositionSample3f = object
    dr.deg2rad = lambda x: x * (np.pi / 180.0)                   # Python floats stay Python floats (double), like Dr.Jit
    dr.cos = lambda x: float(np.cos(x)) if isinstance(x, float) else np.cos(np.asarray(x, dtype=np.float32))

    def acos(x):
        out = _f(LANES); _CURRENT["osv"].call("osv_acos", _flt(x), out); return out.view(DrArray)
    dr.acos = acos
    dr.normalize = lambda v: v * (np.float32(1.0) / np.sqrt(_dot(v, v)))
    dr.dot = _dot
    dr.square = lambda x: (np.asarray(x, dtype=np.float32) * np.asarray(x, dtype=np.float32)).view(DrArray)
    mod = ref_module("mitransient.emitters.angulararea")
    return mi, dr, mod.AngularAreaLight


class _SiFromDs(SurfaceInteraction3f):
    """mi.SurfaceInteraction3f(ds, wavelengths): only handed to Texture.eval by the emitter."""
    def __init__(self, ds=None, wavelengths=None): pass


class AngularScene(ShimScene):
    """A scene whose single emitter is the reference's AngularAreaLight."""
    def attach(self, emitter):
        self.ref_emitter = emitter

    def sample_emitter_direction(self, ref, sample, test_visibility=True, active=True):
        ds, spec = self.ref_emitter.sample_direction(ref, sample, LBool(active))          # Scene: one emitter, pick pdf 1
        if test_visibility:
            ray = ref.spawn_ray_to(ds.p)
            nz = (np.asarray(spec.x) != 0) | (np.asarray(spec.y) != 0) | (np.asarray(spec.z) != 0)
            occluded = self.ray_test(ray, nz)
            spec = sys.modules["drjit"].select(occluded, 0.0, spec)
        return ds, spec

    def pdf_emitter_direction(self, ref, ds, active=True):
        has = np.asarray(ds._shape) == self._emitter_shape
        return self.ref_emitter.pdf_direction(ref, ds, LBool(np.asarray(active, dtype=bool) & has))


class _RefEmitterPtr:
    def __init__(self, scene, shape): self.scene, self.shape = scene, shape

    def eval(self, si, active=True):
        has = np.asarray(self.shape) == self.scene._emitter_shape
        return self.scene.ref_emitter.eval(si, LBool(np.broadcast_to(np.asarray(active, dtype=bool), (LANES,)) & has))


def reference_render_angular(our_scene, props, spp, seed=0):
    """render() + TransientPath.sample() + AngularAreaLight.{sample_direction, pdf_direction, eval}, all reference code."""
    global LANES
    LANES = None
    mi, dr, AngularAreaLight = install_angular_services()
    thf = ref_module("mitransient.films.transient_hdr_film")
    mod = ref_module("mitransient.integrators.transientpath")
    sensor = our_scene.sensors()[0]
    f = sensor.film()
    film = thf.TransientHDRFilm(mi.Properties(width=f.size()[0], height=f.size()[1], temporal_bins=f.temporal_bins,
                                              start_opl=f.start_opl, bin_width_opl=f.bin_width_opl))
    log = []
    inner = film.add_transient_data

    def logged(pos, distance, wavelengths, spec, ray_weight, active, **kw):
        a = np.broadcast_to(np.asarray(active, dtype=bool), (LANES,))
        val = spec * ray_weight
        cols = [np.broadcast_to(np.asarray(c, dtype=np.float32), (LANES,)) for c in (pos.x, pos.y, distance, val.x, val.y, val.z)]
        rec = np.stack(cols, axis=1).astype(np.float32)
        log.append(rec[a & ((rec[:, 3] != 0) | (rec[:, 4] != 0) | (rec[:, 5] != 0))])
        return inner(pos, distance, wavelengths, spec, ray_weight, active, **kw)
    film.add_transient_data = logged
    _real_load = mi.load_dict
    mi.load_dict = lambda d: _steady_film_from_dict(d) if d.get("type") == "hdrfilm" else _real_load(d)
    scene = AngularScene(our_scene, sensor)
    _CURRENT["osv"] = scene.osv
    em = our_scene.emitters()[0]
    scene._emitter_shape = em["shape"]
    deg = 180.0 / np.pi
    ref_em = AngularAreaLight(mi.Properties(radiance=_ConstTexture(em["radiance"]), cutoff_angle=em["cutoff_angle"] * deg,
                                            beam_width=em["beam_width"] * deg, _shape=_EmitterShape(scene)))
    scene.attach(ref_em)
    SurfaceInteraction3f.emitter = lambda self, sc=None: _RefEmitterPtr(scene, self.shape)
    try:
        shim_sensor = ShimSensor(film, spp)
        shim_sensor._sampler.base_seed = sensor.sampler().base_seed
        scene.sensors = lambda: [shim_sensor]
        integ = mod.TransientPath(mi.Properties(**props))
        steady, transient = integ.render(scene, shim_sensor, seed=seed, spp=spp)
    finally:
        SurfaceInteraction3f.emitter = lambda self, sc=None: EmitterPtr(self.osv, self.shape)
        LANES = None
    return np.asarray(steady), film.develop_transient_(raw=True).numpy(), np.concatenate(log, axis=0) if log else np.zeros((0, 6), np.float32)


# ---------------------------------------------------------------------------------------------------------
#  Part 5 -- transient_prbvolpath (integrators/transient_prbvolpath.py): media, phase functions, masked struct
#  assignment (si[mask] = ..., ray[mask] = ..., medium[mask] = ...), lane-wise `while <array>`.
# ---------------------------------------------------------------------------------------------------------
class Struct:
    """A Dr.Jit struct of lane arrays: s[mask] is a copy, s[mask] = other assigns lane-wise, field by field."""
    _fields = ()

    def __getitem__(self, k):
        return _copy.copy(self)

    def __setitem__(self, k, other):
        k = np.asarray(k, dtype=bool)
        for f in self._fields:
            if not hasattr(other, f):
                continue
            a, b = getattr(self, f), getattr(other, f)
            if isinstance(a, Vec) or isinstance(b, Vec):
                av = a if isinstance(a, Vec) else LVec(a)
                bv = b if isinstance(b, Vec) else LVec(b)
                out = LVec.__new__(LVec)
                out.c = tuple(np.where(k, np.asarray(y, dtype=np.float32), np.asarray(x, dtype=np.float32)).view(DrArray) for x, y in zip(av.c, bv.c))
                setattr(self, f, out)
            else:
                setattr(self, f, np.where(k, np.asarray(b), np.asarray(a)).view(DrArray))


class VSurfaceInteraction(SurfaceInteraction3f, Struct):
    _fields = ("valid", "t", "p", "n", "s", "tt", "wi", "shape", "time")

    def emitter(self, scene=None): return VEmitterPtr(self.osv, _i32(self.shape))

    def _medium_info(self, d):
        tgt, tr = np.zeros(LANES, np.int32), np.zeros(LANES, np.uint8)
        self.osv.call("osv_target_medium", _i32(self.shape), self.n.arr(), d.arr(), tgt, tr)
        return tgt, tr

    def target_medium(self, d): return MediumPtr(self.osv, self._medium_info(d)[0])
    def is_medium_transition(self): return (self._medium_info(LVec(0.0))[1] != 0).view(DrArray)

    def spawn_ray_to(self, target): return VRay(SurfaceInteraction3f.spawn_ray_to(self, target))
    def spawn_ray(self, d): return VRay(SurfaceInteraction3f.spawn_ray(self, d))


class VInteraction(Struct):
    """mi.Interaction3f: a position and a normal (zero for medium interactions: spawn_ray_to then applies no offset)."""
    _fields = ("p", "n", "time")

    def __init__(self, osv): self.osv, self.p, self.n, self.time = osv, LVec(0.0), LVec(0.0), LFloat(0.0)

    def spawn_ray_to(self, target):
        o, d, maxt = _f(LANES, 3), _f(LANES, 3), _f(LANES)
        self.osv.call("osv_spawn_ray_to", self.p.arr(), self.n.arr(), target.arr(), o, d, maxt)
        return VRay(Ray3f(LVec.from_arr(o), LVec.from_arr(d), maxt.view(DrArray)))


class VRay(Ray3f, Struct):
    _fields = ("o", "d", "maxt")

    def __init__(self, o=None, d=None, time=None, wavelengths=None):
        if isinstance(o, Ray3f):
            self.o, self.d = LVec(o.o), LVec(o.d)
            self.maxt = np.broadcast_to(np.asarray(o.maxt, dtype=np.float32), (LANES,)).copy().view(DrArray)
        else:                                               # Ray3f(o, d, time, wavelengths): maxt = largest float
            self.o, self.d = LVec(o), LVec(d)
            self.maxt = LFloat(np.finfo(np.float32).max)
        self.time, self.wavelengths = LFloat(0.0), None


class VEmitterPtr(EmitterPtr):
    def __ne__(self, other):
        em = np.zeros(LANES, np.int32); self.osv.call("osv_shape_emitter", _i32(self.shape), em)
        return (em >= 0).view(DrArray)
    __hash__ = None


class MediumInteraction(Struct):
    _fields = ("t", "p", "n", "sigma_t", "sigma_s", "sigma_n", "combined_extinction", "wi", "time", "_valid", "_medium")

    def is_valid(self): return (np.asarray(self.t) < np.inf).view(DrArray)        # MediumInteraction::is_valid(): t < inf
    @property
    def medium(self): return MediumPtr(self.osv, self._medium)

    def spawn_ray(self, wo):
        return VRay(Ray3f(self.p, wo, LFloat(np.finfo(np.float32).max)))


class PhasePtr:
    def __init__(self, osv, idx): self.osv, self.idx = osv, np.asarray(idx, dtype=np.int32)
    def __getitem__(self, k): return PhasePtr(self.osv, self.idx.copy())
    def __setitem__(self, k, other): self.idx = np.where(np.asarray(k, dtype=bool), other.idx, self.idx).astype(np.int32)

    def eval_pdf(self, ctx, mei, wo, active=True):
        val = _f(LANES)
        self.osv.call("osv_phase_eval", _i32(self.idx), mei.wi.arr(), wo.arr(), _mask(active), val)
        return val.view(DrArray), val.copy().view(DrArray)

    def sample(self, ctx, mei, s1, s2, active=True):
        wo, pdf = _f(LANES, 3), _f(LANES)
        self.osv.call("osv_phase_sample", _i32(self.idx), mei.wi.arr(), s2.arr(), _mask(active), wo, pdf)
        return LVec.from_arr(wo), LFloat(1.0), pdf.view(DrArray)                      # HG: sampled exactly, weight 1


class MediumPtr:
    def __init__(self, osv, idx): self.osv, self.idx = osv, np.broadcast_to(np.asarray(idx, dtype=np.int32), (LANES,)).copy()
    def __ne__(self, other): return (self.idx >= 0).view(DrArray)
    def __eq__(self, other): return (self.idx < 0).view(DrArray)
    __hash__ = None
    def __getitem__(self, k): return MediumPtr(self.osv, self.idx.copy())
    def __setitem__(self, k, other): self.idx = np.where(np.asarray(k, dtype=bool), other.idx, self.idx).astype(np.int32)
    def is_homogeneous(self): return LBool(True)
    def use_emitter_sampling(self): return LBool(True)
    def phase_function(self): return PhasePtr(self.osv, self.idx.copy())

    def sample_interaction(self, ray, u, channel, active=True):
        n = LANES
        valid, t, p, st, ss = np.zeros(n, np.uint8), _f(n), _f(n, 3), _f(n, 3), _f(n, 3)
        self.osv.call("osv_medium_sample_interaction", _i32(self.idx), ray.o.arr(), ray.d.arr(), _flt(ray.maxt), _flt(u),
                      _i32(channel), _mask(active), valid, t, p, st, ss)
        mei = MediumInteraction(); mei.osv = self.osv
        mei.t, mei.p, mei.n = t.view(DrArray), LVec.from_arr(p), LVec(0.0)
        mei.sigma_t, mei.sigma_s, mei.sigma_n, mei.combined_extinction = LVec.from_arr(st), LVec.from_arr(ss), LVec(0.0), LVec.from_arr(st)
        mei.wi, mei.time, mei._valid, mei._medium = -ray.d, LFloat(0.0), valid != 0, self.idx.copy()
        return mei

    def transmittance_eval_pdf(self, mei, si, active=True):
        tr, pdf = _f(LANES, 3), _f(LANES, 3)
        self.osv.call("osv_medium_transmittance", _flt(mei.t), mei.sigma_t.arr(), _flt(si.t), _mask(active), tr, pdf)
        return LVec.from_arr(tr), LVec.from_arr(pdf)


class VScene(ShimScene):
    def ray_intersect(self, ray, active=True, **kw):
        si0 = ShimScene.ray_intersect(self, ray)
        si = VSurfaceInteraction(); si.__dict__.update(si0.__dict__)
        si.shape = np.asarray(si.shape, dtype=np.int32).view(DrArray)
        if isinstance(active, np.ndarray):                        # scene.ray_intersect(ray, active): inactive lanes = no hit
            a = np.asarray(active, dtype=bool)
            si.valid = (np.asarray(si.valid) & a).view(DrArray)
            si.t = np.where(a, si.t, np.float32(np.inf)).view(DrArray)
            si.shape = np.where(a, si.shape, -1).astype(np.int32).view(DrArray)
        return si

    def shapes(self):
        media = [(int(self.our_scene.shapes()[i].interior_medium), int(self.our_scene.shapes()[i].exterior_medium))
                 for i in range(len(self.our_scene.shapes()))]
        mk = lambda m: (types.SimpleNamespace(use_emitter_sampling=lambda: True, is_homogeneous=lambda: True) if m >= 0 else None)
        return [types.SimpleNamespace(interior_medium=(lambda a=a: mk(a)), exterior_medium=(lambda b=b: mk(b))) for a, b in media]


def _rewrite_syntax_vol(fn):
    """dr.syntax for the volumetric integrator: array-`if`s (as above) and `while <array>:` -> `while any(array):`."""
    src = textwrap.dedent(inspect.getsource(fn))
    tree = ast.parse(src)
    fdef = tree.body[0]
    fdef.decorator_list = []

    class W(ast.NodeTransformer):
        def visit_While(self, node):
            self.generic_visit(node)
            node.test = ast.Call(func=ast.Name(id="__shim_any", ctx=ast.Load()), args=[node.test], keywords=[])
            return node
    tree = ast.fix_missing_locations(_SymbolicLoopState().visit(W().visit(tree)))
    g = fn.__globals__
    g["__shim_snapshot"], g["__shim_restore"] = _shim_snapshot, _shim_restore
    g["__shim_any"] = lambda c: bool(np.any(c))
    loc = {}
    exec(compile(tree, inspect.getsourcefile(fn), "exec"), g, loc)
    return loc[fdef.name]


def reference_render_vol(our_scene, props, spp, seed=0):
    """Run the REFERENCE'S render() + TransientPRBVolpathIntegrator.sample / sample_emitter / first_surface."""
    import mitransient_b200 as mitr
    global LANES
    LANES = None
    mi, dr = install_scene_services(False, mitr.ScalarTransform4f, mitr.ScalarColor3d)
    mi.is_rgb = True
    mi.MediumPtr, mi.PhaseFunctionPtr, mi.Interaction3f = "MediumPtr", "PhaseFunctionPtr", "Interaction3f"
    mi.PhaseFunctionContext = lambda *a, **k: None
    mi.SurfaceInteraction3f = VSurfaceInteraction
    mi.Ray3f = mi.RayDifferential3f = VRay
    mi.Mask = LBool
    mi.math = types.SimpleNamespace(RayEpsilon=np.float32(8.940696716308594e-05))
    mi.BSDFFlags.Null = 0x8
    dr.size_v = lambda t: 3
    dr.syntax = lambda f=None, **kw: (_rewrite_syntax_vol(f) if f is not None else (lambda g: _rewrite_syntax_vol(g)))
    dr.square = lambda x: (np.asarray(x, dtype=np.float32) * np.asarray(x, dtype=np.float32)).view(DrArray)
    osv_holder = {}
    base_zeros = dr.zeros

    def zeros(t, n=1):
        if t == "MediumPtr": return MediumPtr(osv_holder["osv"], -1)
        if t == "PhaseFunctionPtr": return PhasePtr(osv_holder["osv"], np.full(LANES, -1, np.int32))
        if t == "Interaction3f": return VInteraction(osv_holder["osv"])
        if isinstance(t, type) and issubclass(t, SurfaceInteraction3f):
            si = VSurfaceInteraction(); si.__dict__.update(_zero_si(osv_holder["osv"]).__dict__)
            si.shape = np.full(LANES, -1, np.int32).view(DrArray); si.time = LFloat(0.0)
            si.p = si.n = si.s = si.tt = si.wi = LVec(LFloat(0.0), LFloat(0.0), LFloat(0.0))
            return si
        return base_zeros(t, n)
    dr.zeros = zeros
    base_select = dr.select

    def select(m, a, b):
        if isinstance(a, MediumPtr) and isinstance(b, MediumPtr):
            return MediumPtr(a.osv, np.where(np.asarray(m, dtype=bool), a.idx, b.idx))
        if isinstance(a, np.ndarray) and a.dtype.kind in "iu" and not isinstance(b, Vec):
            return np.where(m, a, b).view(DrArray)
        return base_select(m, a, b)
    dr.select = select
    dr.any = lambda v: np.logical_or.reduce([np.broadcast_to(np.asarray(x), (LANES,)) for x in v.c]).view(DrArray) if isinstance(v, Vec) else np.any(v)
    Vec.__ne__ = lambda self, o: self._zip(o, lambda a, b: a != b)

    # BSDFPtr additions: Null flag, eval_null_transmission
    def flags(self):
        sm, nl = np.zeros(LANES, np.uint8), np.zeros(LANES, np.uint8)
        self.osv.call("osv_bsdf_smooth", _i32(self.shape), sm, nl)
        return np.where(sm != 0, BSDFPtr.Smooth, np.where(nl != 0, 0x8 | BSDFPtr.Delta, BSDFPtr.Delta)).astype(np.int64).view(DrArray)

    def sample(self, ctx, si, sample1, sample2, active=True):
        wo, pdf, eta, delta, w = _f(LANES, 3), _f(LANES), _f(LANES), np.zeros(LANES, np.uint8), _f(LANES, 3)
        self.osv.call("osv_bsdf_sample", _i32(self.shape), si.wi.arr(), _flt(sample1), sample2.arr(), _mask(active), wo, pdf, eta, delta, w)
        fl = self.flags()
        st = np.where(np.asarray(fl) & 0x8, 0x8 | BSDFPtr.Delta, np.where(delta != 0, BSDFPtr.Delta, BSDFPtr.Smooth)).astype(np.int64).view(DrArray)
        return types.SimpleNamespace(wo=LVec.from_arr(wo), pdf=pdf.view(DrArray), eta=eta.view(DrArray), sampled_type=st), LVec.from_arr(w)

    def eval_null_transmission(self, si, active=True):
        sm, nl = np.zeros(LANES, np.uint8), np.zeros(LANES, np.uint8)
        self.osv.call("osv_bsdf_smooth", _i32(self.shape), sm, nl)
        v = np.where((nl != 0) & np.broadcast_to(np.asarray(active, dtype=bool), (LANES,)), np.float32(1), np.float32(0))
        return LVec(v, v, v)
    saved = (BSDFPtr.flags, BSDFPtr.sample)
    BSDFPtr.flags, BSDFPtr.sample, BSDFPtr.eval_null_transmission = flags, sample, eval_null_transmission
    try:
        thf = ref_module("mitransient.films.transient_hdr_film")
        mod = ref_module("mitransient.integrators.transient_prbvolpath")
        sensor = our_scene.sensors()[0]
        f = sensor.film()
        film = thf.TransientHDRFilm(mi.Properties(width=f.size()[0], height=f.size()[1], temporal_bins=f.temporal_bins,
                                                  start_opl=f.start_opl, bin_width_opl=f.bin_width_opl))
        log = []
        inner = film.add_transient_data

        def logged(pos, distance, wavelengths, spec, ray_weight, active, **kw):
            a = np.broadcast_to(np.asarray(active, dtype=bool), (LANES,))
            val = spec * ray_weight
            cols = [np.broadcast_to(np.asarray(c, dtype=np.float32), (LANES,)) for c in (pos.x, pos.y, distance, val.x, val.y, val.z)]
            rec = np.stack(cols, axis=1).astype(np.float32)
            log.append(rec[a & ((rec[:, 3] != 0) | (rec[:, 4] != 0) | (rec[:, 5] != 0))])
            return inner(pos, distance, wavelengths, spec, ray_weight, active, **kw)
        film.add_transient_data = logged
        _real_load = mi.load_dict
        mi.load_dict = lambda d: _steady_film_from_dict(d) if d.get("type") == "hdrfilm" else _real_load(d)
        scene = VScene(our_scene, sensor)
        _CURRENT["osv"] = osv_holder["osv"] = scene.osv
        shim_sensor = ShimSensor(film, spp)
        shim_sensor._sampler.base_seed = sensor.sampler().base_seed
        scene.sensors = lambda: [shim_sensor]
        integ = mod.TransientPRBVolpathIntegrator(mi.Properties(**props))
        steady, transient = integ.render(scene, shim_sensor, seed=seed, spp=spp)
    finally:
        BSDFPtr.flags, BSDFPtr.sample = saved
        del BSDFPtr.eval_null_transmission
        del Vec.__ne__
        LANES = None
    return np.asarray(steady), film.develop_transient_(raw=True).numpy(), np.concatenate(log, axis=0) if log else np.zeros((0, 6), np.float32)
