"""Laser aiming helpers -- counterpart of mitransient/nlos.py (used by *single* captures)."""
from __future__ import annotations

import numpy as np

from .transforms import ScalarTransform4f


def focus_emitter_at_relay_wall_3dpoint(target, relay_wall, emitter):
    """Aim the laser (a `projector` PluginSpec) at ``target``; updates the sensor's ``laser_target`` /
    ``laser_bounce_opl`` like the reference (nlos.py:5-34)."""
    sensor = relay_wall.sensor()
    target = np.asarray(target, dtype=np.float64).reshape(3)
    origin = emitter.world_transform().translation()
    emitter["to_world"] = ScalarTransform4f().look_at(origin=origin, target=target, up=[0, 1, 0])
    sensor.laser_bounce_opl = float(np.linalg.norm(target - origin))
    sensor.laser_target = target


def focus_emitter_at_relay_wall_uv(uv, relay_wall, emitter):      # nlos.py:37-50
    target = relay_wall.sample_position(0.0, uv, True).p
    return focus_emitter_at_relay_wall_3dpoint(target, relay_wall, emitter)


def focus_emitter_at_relay_wall_pixel(pixel, relay_wall, emitter):  # nlos.py:53-70
    sensor = relay_wall.sensor()
    fw, fh = sensor.film_size
    return focus_emitter_at_relay_wall_uv((float(pixel[0]) / fw, float(pixel[1]) / fh), relay_wall, emitter)
