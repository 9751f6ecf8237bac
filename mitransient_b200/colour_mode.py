"""Colour mode of the host layer (``mi.is_monochromatic`` in the reference, e.g. transient_hdr_film.py:174,
phasor_hdr_film.py:147).  The kernels always carry three channels; in monochromatic mode every scene colour is
replaced by its luminance at flattening time, so the three channels are equal and channel 0 is "L"."""
from __future__ import annotations

_mono = False


def set_monochromatic(flag: bool) -> None:
    global _mono
    _mono = bool(flag)


def is_monochromatic() -> bool:
    return _mono


def luminance(rgb):
    """Mitsuba's ``luminance(Color3f)`` (Rec. 709 weights), evaluated in float32 like the reference would."""
    import numpy as np
    r, g, b = (np.float32(x) for x in rgb)
    l = np.float32(0.212671) * r + np.float32(0.715160) * g + np.float32(0.072169) * b
    return np.array([l, l, l], dtype=np.float64)


def colour(rgb):
    """A scene colour as the kernels see it in the active variant."""
    return luminance(rgb) if _mono else rgb
