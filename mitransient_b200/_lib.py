"""Loader for the CUDA C-ABI library (include/mitransient_b200.h).

There is NO CPU fallback: if the shared object is missing or no B200 is visible, every compute call
raises.  (The CPU oracle under oracle/ is test infrastructure and is never imported from here.)
"""
from __future__ import annotations

import ctypes as C
import os
import threading

from . import _abi

_LIB_PATH = os.environ.get("MITRANSIENT_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib",
                                                                   "libmitransient_b200.so")
_lock = threading.Lock()
_lib = None


class TransientError(RuntimeError):
    """A C-ABI call returned a negative TrStatus."""

    def __init__(self, code: int, message: str):
        super().__init__(f"[TrStatus {code}] {message}")
        self.code = code


def lib_path() -> str:
    return _LIB_PATH


def load() -> C.CDLL:
    """Load (once) and type the library.  Raises ImportError with build instructions if it is absent."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(_LIB_PATH):
            raise ImportError(
                f"{_LIB_PATH} not found: the CUDA extension has not been built.  Run "
                "`python -m mitransient_b200.build` (needs nvcc; cross-compiles for sm_100a).  "
                "mitransient_b200 has no CPU fallback.")
        lib = C.CDLL(_LIB_PATH)
        vp = C.c_void_p
        lib.tr_last_error.restype = C.c_char_p
        lib.tr_last_error.argtypes = []
        lib.tr_abi_version.restype = C.c_int
        lib.tr_device_check.restype = C.c_int
        lib.tr_scene_create.restype = C.c_int
        lib.tr_scene_create.argtypes = [C.POINTER(_abi.TrSceneDesc), C.POINTER(vp)]
        lib.tr_scene_destroy.restype = None
        lib.tr_scene_destroy.argtypes = [vp]
        lib.tr_scene_refresh.restype = C.c_int
        lib.tr_scene_refresh.argtypes = [vp, vp]
        lib.tr_scene_info.restype = C.c_int
        lib.tr_scene_info.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        lib.tr_render.restype = C.c_int
        lib.tr_render.argtypes = [vp, C.POINTER(_abi.TrRenderParams), vp, vp, vp, vp]
        lib.tr_render_to_host.restype = C.c_int
        lib.tr_render_to_host.argtypes = [vp, C.POINTER(_abi.TrRenderParams), vp, vp, C.POINTER(_abi.TrCounters), vp]
        lib.tr_owned_pixels.restype = C.c_int
        lib.tr_owned_pixels.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, vp, C.POINTER(C.c_uint64)]
        lib.tr_film_scatter_add.restype = C.c_int
        lib.tr_film_scatter_add.argtypes = [vp, C.POINTER(_abi.TrFilmDesc), C.c_int32, vp, C.c_uint64, vp]
        lib.tr_film_develop.restype = C.c_int
        lib.tr_film_develop.argtypes = [vp, vp, C.c_uint64, C.c_int32, vp]
        lib.tr_last_render_info.restype = None
        lib.tr_last_render_info.argtypes = [C.POINTER(_abi.TrRenderInfo)]
        lib.tr_selftest_rcp.restype = C.c_int
        lib.tr_selftest_rcp.argtypes = [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), vp]
        lib.tr_peer_alloc.restype = C.c_int
        lib.tr_peer_alloc.argtypes = [C.c_uint64, C.POINTER(vp), C.c_char_p]
        lib.tr_peer_open.restype = C.c_int
        lib.tr_peer_open.argtypes = [C.c_char_p, C.POINTER(vp)]
        lib.tr_peer_close.restype = C.c_int
        lib.tr_peer_close.argtypes = [vp]
        lib.tr_peer_free.restype = C.c_int
        lib.tr_peer_free.argtypes = [vp]
        if lib.tr_abi_version() != _abi.TR_ABI_VERSION:
            raise ImportError(f"{_LIB_PATH} was built for ABI {lib.tr_abi_version()}, "
                              f"the Python host expects {_abi.TR_ABI_VERSION}: rebuild it")
        _lib = lib
        return lib


def check(code: int) -> None:
    if code != _abi.TR_OK:
        msg = load().tr_last_error()
        raise TransientError(code, msg.decode("utf-8", "replace") if msg else "unknown error")


def current_stream_ptr() -> int:
    import torch
    return int(torch.cuda.current_stream().cuda_stream)


class DeviceScene:
    """Owns an opaque ``TrScene*`` (device copy of the scene + BVH)."""

    def __init__(self, desc: _abi.TrSceneDesc):
        self._lib = load()
        handle = C.c_void_p()
        check(self._lib.tr_scene_create(C.byref(desc), C.byref(handle)))
        self.handle = handle

    def info(self):
        b, n, t = C.c_uint64(), C.c_uint32(), C.c_uint32()
        check(self._lib.tr_scene_info(self.handle, C.byref(b), C.byref(n), C.byref(t)))
        return {"blob_bytes": b.value, "bvh_nodes": n.value, "triangles": t.value, "primitives": t.value}

    def close(self):
        if getattr(self, "handle", None):
            self._lib.tr_scene_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
