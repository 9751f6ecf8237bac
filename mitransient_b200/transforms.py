"""ScalarTransform4f stand-in.

The reference builds its scenes with ``mi.ScalarTransform4f`` (mitransient/utils.py:82,101-105,
149,166-216).  Mitsuba is not available to this build, so the host side carries a small float64
4x4 transform with the same chaining semantics (Mitsuba >= 3.6: ``T().translate(a).rotate(b)``
composes on the right, i.e. equals ``translate(a) @ rotate(b)``).
"""
from __future__ import annotations

import math
from typing import Sequence

import numpy as np


def _vec3(v) -> np.ndarray:
    if np.isscalar(v):
        return np.array([v, v, v], dtype=np.float64)
    a = np.asarray(v, dtype=np.float64).reshape(-1)
    if a.size != 3:
        raise ValueError(f"expected a 3-vector, got {v!r}")
    return a


class ScalarTransform4f:
    """Affine/projective 4x4 transform in float64 (cast to fp32 when the scene is flattened)."""

    def __init__(self, matrix=None):
        if matrix is None:
            self.matrix = np.eye(4, dtype=np.float64)
        elif isinstance(matrix, ScalarTransform4f):
            self.matrix = matrix.matrix.copy()
        else:
            m = np.asarray(matrix, dtype=np.float64)
            if m.shape != (4, 4):
                raise ValueError("ScalarTransform4f needs a 4x4 matrix")
            self.matrix = m.copy()

    # -- composition ---------------------------------------------------------------------
    def __matmul__(self, other):
        if isinstance(other, ScalarTransform4f):
            return ScalarTransform4f(self.matrix @ other.matrix)
        v = np.asarray(other, dtype=np.float64).reshape(-1)
        if v.size == 3:  # treated as a point
            return self.transform_affine(v)
        raise TypeError("unsupported operand for @")

    def _chain(self, m: np.ndarray) -> "ScalarTransform4f":
        return ScalarTransform4f(self.matrix @ m)

    def translate(self, v) -> "ScalarTransform4f":
        m = np.eye(4)
        m[:3, 3] = _vec3(v)
        return self._chain(m)

    def scale(self, v) -> "ScalarTransform4f":
        m = np.eye(4)
        s = _vec3(v)
        m[0, 0], m[1, 1], m[2, 2] = s
        return self._chain(m)

    def rotate(self, axis, angle: float) -> "ScalarTransform4f":
        """Rotation by ``angle`` DEGREES about ``axis`` (Mitsuba convention)."""
        a = _vec3(axis)
        a = a / np.linalg.norm(a)
        th = math.radians(float(angle))
        s, c = math.sin(th), math.cos(th)
        x, y, z = a
        m = np.eye(4)
        m[:3, :3] = np.array([
            [x * x + (1 - x * x) * c, x * y * (1 - c) - z * s, x * z * (1 - c) + y * s],
            [x * y * (1 - c) + z * s, y * y + (1 - y * y) * c, y * z * (1 - c) - x * s],
            [x * z * (1 - c) - y * s, y * z * (1 - c) + x * s, z * z + (1 - z * z) * c],
        ])
        return self._chain(m)

    def look_at(self, origin, target, up) -> "ScalarTransform4f":
        o, t, u = _vec3(origin), _vec3(target), _vec3(up)
        d = t - o
        d = d / np.linalg.norm(d)
        left = np.cross(u, d)
        n = np.linalg.norm(left)
        if n == 0:
            raise ValueError("look_at(): 'up' is parallel to the viewing direction")
        left = left / n
        new_up = np.cross(d, left)
        m = np.eye(4)
        m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = left, new_up, d, o
        return self._chain(m)

    @staticmethod
    def perspective(fov: float, near: float, far: float) -> "ScalarTransform4f":
        recip = 1.0 / (far - near)
        cot = 1.0 / math.tan(math.radians(fov * 0.5))
        m = np.zeros((4, 4))
        m[0, 0] = cot
        m[1, 1] = cot
        m[2, 2] = far * recip
        m[2, 3] = -near * far * recip
        m[3, 2] = 1.0
        return ScalarTransform4f(m)

    # -- queries -------------------------------------------------------------------------
    def inverse(self) -> "ScalarTransform4f":
        return ScalarTransform4f(np.linalg.inv(self.matrix))

    def translation(self) -> np.ndarray:
        return self.matrix[:3, 3].copy()

    def transform_affine(self, p) -> np.ndarray:
        p = _vec3(p)
        return self.matrix[:3, :3] @ p + self.matrix[:3, 3]

    def transform_vector(self, v) -> np.ndarray:
        return self.matrix[:3, :3] @ _vec3(v)

    def affine3x4(self) -> np.ndarray:
        """Row-major 3x4 fp32 block handed to the C ABI."""
        return np.ascontiguousarray(self.matrix[:3, :4], dtype=np.float32)

    def __repr__(self):
        return f"ScalarTransform4f(\n{self.matrix}\n)"


def perspective_projection(film_size: Sequence[int], crop_size: Sequence[int], crop_offset: Sequence[int],
                           fov_x: float, near_clip: float, far_clip: float) -> ScalarTransform4f:
    """Mitsuba's ``perspective_projection`` (camera space -> [0,1]^2 sample space)."""
    fw, fh = float(film_size[0]), float(film_size[1])
    rel_size = (crop_size[0] / fw, crop_size[1] / fh)
    rel_off = (crop_offset[0] / fw, crop_offset[1] / fh)
    aspect = fw / fh
    T = ScalarTransform4f
    return (T().scale([1.0 / rel_size[0], 1.0 / rel_size[1], 1.0])
            .translate([-rel_off[0], -rel_off[1], 0.0])
            .scale([-0.5, -0.5 * aspect, 1.0])
            .translate([-1.0, -1.0 / aspect, 0.0])
            @ T.perspective(fov_x, near_clip, far_clip))


def parse_fov(fov: float, fov_axis: str, aspect: float) -> float:
    """Mitsuba's ``parse_fov``: horizontal field of view in degrees."""
    axis = fov_axis
    if axis == "smaller":
        axis = "y" if aspect > 1 else "x"
    elif axis == "larger":
        axis = "x" if aspect > 1 else "y"
    if axis == "x":
        return fov
    if axis == "y":
        return math.degrees(2.0 * math.atan(math.tan(0.5 * math.radians(fov)) * aspect))
    if axis == "diagonal":
        diagonal = 2.0 * math.tan(0.5 * math.radians(fov))
        width = diagonal / math.sqrt(1.0 + 1.0 / (aspect * aspect))
        return math.degrees(2.0 * math.atan(width * 0.5))
    raise ValueError(f"parse_fov(): invalid fov_axis {fov_axis!r}")
