"""ctypes wrapper of the CPU oracle (oracle/transient_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Nothing under mitransient_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from mitransient_b200 import _abi   # struct definitions of include/mitransient_b200.h (data only)

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libtransient_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "transient_oracle.cpp")
    hdr = os.path.join(HERE, "..", "include", "mitransient_b200.h")
    stale = (not os.path.exists(LIB)) or any(os.path.getmtime(p) > os.path.getmtime(LIB) for p in (src, hdr))
    if force or stale:
        r = subprocess.run(["make", "-C", HERE] + (["-B"] if force else []), stdout=subprocess.PIPE,
                           stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + r.stdout)
    return LIB


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        build()          # no-op unless the sources are newer than the library
        lib = C.CDLL(LIB)
        lib.oracle_last_error.restype = C.c_char_p
        lib.oracle_render.restype = C.c_int
        lib.oracle_render.argtypes = [C.POINTER(_abi.TrSceneDesc), C.POINTER(_abi.TrRenderParams), C.c_void_p,
                                      C.c_void_p, C.POINTER(_abi.TrCounters), C.c_int, C.c_void_p, C.c_uint64,
                                      C.POINTER(C.c_uint64)]
        lib.oracle_film_scatter_add.restype = C.c_int
        lib.oracle_film_scatter_add.argtypes = [C.c_void_p, C.POINTER(_abi.TrFilmDesc), C.c_int, C.c_void_p, C.c_uint64]
        lib.oracle_film_scatter_add_laser.restype = C.c_int
        lib.oracle_film_scatter_add_laser.argtypes = [C.c_void_p, C.POINTER(_abi.TrFilmDesc), C.c_int, C.c_void_p, C.c_void_p,
                                                      C.c_uint64]
        lib.oracle_film_develop.restype = C.c_int
        lib.oracle_film_develop.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int]
        lib.oracle_pcg32.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p, C.c_int]
        lib.oracle_tea32.argtypes = [C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        lib.oracle_sampler_floats.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p, C.c_int]
        lib.oracle_sincos.argtypes = [C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        lib.oracle_fresnel_conductor.restype = C.c_float
        lib.oracle_fresnel_conductor.argtypes = [C.c_float, C.c_float, C.c_float]
        lib.oracle_fresnel_dielectric.argtypes = [C.c_float, C.c_float, C.c_void_p]
        lib.oracle_exp.restype = C.c_float; lib.oracle_exp.argtypes = [C.c_float]
        lib.oracle_log.restype = C.c_float; lib.oracle_log.argtypes = [C.c_float]
        lib.oracle_cosine_hemisphere.argtypes = [C.c_float, C.c_float, C.c_void_p]
        lib.oracle_camera_ray.restype = C.c_int
        lib.oracle_camera_ray.argtypes = [C.POINTER(_abi.TrSceneDesc), C.c_uint32, C.c_uint32, C.c_float, C.c_float,
                                          C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]
        lib.oracle_ray_intersect.restype = C.c_int
        lib.oracle_ray_intersect.argtypes = [C.POINTER(_abi.TrSceneDesc), C.c_void_p, C.c_void_p, C.c_float,
                                             C.POINTER(C.c_float), C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32),
                                             C.POINTER(C.c_uint32)]
        _lib = lib
    return _lib


def _check(rc):
    if rc != 0:
        raise RuntimeError(f"oracle error {rc}: {load().oracle_last_error().decode()}")


def render(scene, integrator, sensor=0, seed=0, spp=0, raw=True, n_threads=0, rank=0, world=1, log_deposits=0):
    """Render ``scene`` (a mitransient_b200 Scene) with the oracle.
    Returns dict(steady (H,W,3), transient (H,W,T,C), counters, deposits (N,6) or None)."""
    lib = load()
    if isinstance(sensor, int):
        sensor = scene.sensors()[sensor]
    film = sensor.film()
    s = sensor.sampler().clone()
    if spp:
        s.set_sample_count(spp)
    spp = s.sample_count()
    desc = scene.describe(sensor)
    w, h = film.crop_size()
    if desc.film.type == _abi.TR_FILM_PHASOR_HDR:      # (H, W, 2F + 1), always raw
        C_ = film.raw_channels()
        shape = (h, w, C_)
    elif desc.film.laser_scan_width:                   # exhaustive_scan: (H, W, Lh, Lw, T, C)
        C_ = 4 if raw else 3
        shape = (h, w, film.laser_scan_height, film.laser_scan_width, film.temporal_bins, C_)
    else:
        C_ = 4 if raw else 3
        shape = (h, w, film.temporal_bins, C_)
    params = integrator.make_params(spp, spp, (int(seed) + s.base_seed) & 0xFFFFFFFF, rank=rank, world=world,
                                    film_channels=C_)
    integrator.fill_params_(params, scene, sensor)
    out = np.zeros(shape, dtype=np.float32)
    steady = np.zeros((h, w, 3), dtype=np.float32)
    cnt = _abi.TrCounters()
    log = np.zeros((max(1, log_deposits), 6), dtype=np.float32)
    n_logged = C.c_uint64(0)
    _check(lib.oracle_render(C.byref(desc), C.byref(params), out.ctypes.data, steady.ctypes.data, C.byref(cnt),
                             int(n_threads), log.ctypes.data if log_deposits else None, int(log_deposits),
                             C.byref(n_logged)))
    if getattr(film, "is_cropped", lambda: False)():
        # the reference's transient block keeps the full film's shape with the window in its top-left corner
        # (transient_hdr_film.py:192-195)
        fw, fh = film.size()
        full = np.zeros((fh, fw) + out.shape[2:], dtype=np.float32)
        full[:h, :w] = out
        out = full
    return {"steady": steady, "transient": out, "counters": cnt.as_dict(),
            "deposits": log[: n_logged.value] if log_deposits else None}


def film_scatter_add(film: np.ndarray, film_desc: _abi.TrFilmDesc, deposits: np.ndarray):
    assert film.dtype == np.float32 and film.flags.c_contiguous
    dep = np.ascontiguousarray(deposits, dtype=np.float32)
    _check(load().oracle_film_scatter_add(film.ctypes.data, C.byref(film_desc), film.shape[-1], dep.ctypes.data,
                                          dep.shape[0]))


def film_scatter_add_laser(film: np.ndarray, film_desc: _abi.TrFilmDesc, deposits: np.ndarray, laser_xy: np.ndarray):
    assert film.dtype == np.float32 and film.flags.c_contiguous
    dep = np.ascontiguousarray(deposits, dtype=np.float32)
    lxy = np.ascontiguousarray(laser_xy, dtype=np.uint32)
    _check(load().oracle_film_scatter_add_laser(film.ctypes.data, C.byref(film_desc), film.shape[-1], dep.ctypes.data,
                                                lxy.ctypes.data, dep.shape[0]))


def film_develop(raw: np.ndarray) -> np.ndarray:
    raw = np.ascontiguousarray(raw, dtype=np.float32)
    out = np.empty(raw.shape[:-1] + (raw.shape[-1] - 1,), dtype=np.float32)
    _check(load().oracle_film_develop(raw.ctypes.data, out.ctypes.data, raw.size // raw.shape[-1], raw.shape[-1]))
    return out


def pcg32(initstate, initseq, n):
    out = np.zeros(n, dtype=np.uint32)
    load().oracle_pcg32(initstate, initseq, out.ctypes.data, n)
    return out


def tea32(v0, v1):
    a, b = C.c_uint32(), C.c_uint32()
    load().oracle_tea32(v0, v1, C.byref(a), C.byref(b))
    return a.value, b.value


def sampler_floats(seed, lane, n):
    out = np.zeros(n, dtype=np.float32)
    load().oracle_sampler_floats(seed, lane, out.ctypes.data, n)
    return out


def sincos(x):
    s, c = C.c_float(), C.c_float()
    load().oracle_sincos(x, C.byref(s), C.byref(c))
    return s.value, c.value


def fresnel_conductor(cos_theta_i, eta, k):
    return load().oracle_fresnel_conductor(cos_theta_i, eta, k)


def fresnel_dielectric(cos_theta_i, eta):
    """-> (r, cos_theta_t, eta_it, eta_ti) like mitsuba's fresnel()"""
    out = np.zeros(4, dtype=np.float32)
    load().oracle_fresnel_dielectric(cos_theta_i, eta, out.ctypes.data)
    return tuple(float(x) for x in out)


def cosine_hemisphere(u0, u1):
    out = np.zeros(3, dtype=np.float32)
    load().oracle_cosine_hemisphere(u0, u1, out.ctypes.data)
    return out


def camera_ray(scene, sensor, px, py, jx, jy):
    desc = scene.describe(sensor)
    o, d, maxt = np.zeros(3, np.float32), np.zeros(3, np.float32), C.c_float()
    _check(load().oracle_camera_ray(C.byref(desc), px, py, jx, jy, o.ctypes.data, d.ctypes.data, C.byref(maxt)))
    return o, d, maxt.value


def ray_intersect(scene, sensor, o, d, maxt=3.4e38):
    desc = scene.describe(sensor)
    o = np.ascontiguousarray(o, np.float32); d = np.ascontiguousarray(d, np.float32)
    t = C.c_float(); p = np.zeros(3, np.float32); n = np.zeros(3, np.float32)
    shape, prim = C.c_uint32(), C.c_uint32()
    hit = load().oracle_ray_intersect(C.byref(desc), o.ctypes.data, d.ctypes.data, maxt, C.byref(t), p.ctypes.data,
                                      n.ctypes.data, C.byref(shape), C.byref(prim))
    if hit < 0:
        _check(hit)
    if hit == 0:
        return None
    return {"t": t.value, "p": p, "n": n, "shape": shape.value, "prim": prim.value}


def exp(x):
    return load().oracle_exp(x)


def log(x):
    return load().oracle_log(x)
