"""``nlos_capture_meter`` -- host mirror of mitransient/sensors/nloscapturemeter.py (parameters only;
ray generation, nloscapturemeter.py:136-202, runs on the device)."""
from __future__ import annotations

import numpy as np

from .. import _abi
from ..scene import IndependentSampler, Properties
from ..transforms import ScalarTransform4f


class NLOSSensor:          # mitransient/sensors/nlossensor.py:5-10
    pass


class NLOSCaptureMeter(NLOSSensor):
    plugin_name = "nlos_capture_meter"

    def __init__(self, props: Properties):
        from .. import plugins
        self.needs_sample_3 = False
        if "account_first_and_last_bounces" in props:
            import warnings
            warnings.warn("account_first_and_last_bounces in nlos_capture_meter has been deprecated. "
                          "Please, use account_first_and_last_bounces in transient_nlos_path instead.")
        props.get("account_first_and_last_bounces", True)
        origin = np.asarray(props.get("sensor_origin", (0.0, 0.0, 0.0)), dtype=np.float64).reshape(3)
        self._world_transform = ScalarTransform4f().translate(origin)
        self.laser_bounce_opl = 0.0
        self.laser_target = np.zeros(3)
        # nested plugins are found by TYPE, under any key (the notebooks use "transient_film", the XML "film")
        from ..scene import _find_child
        _, fspec = _find_child(props, tuple(plugins.FILMS))
        if fspec is None:
            fspec = {"type": "transient_hdr_film"}
        self._film = fspec if not isinstance(fspec, dict) else plugins.instantiate(fspec)
        _, sspec = _find_child(props, ("independent",))
        self._sampler = IndependentSampler(Properties(sspec or {}))
        if hasattr(self._film, "crop_size") and (tuple(self._film.crop_size()) != tuple(self._film.size()) or tuple(self._film.crop_offset()) != (0, 0)):
            # the reference's sensor divides crop-relative film samples by the FULL film size (nloscapturemeter.py:133-150)
            raise NotImplementedError("nlos_capture_meter: films with a crop window are not supported")
        self.original_film_width = props.get("original_film_width", None)
        self.original_film_height = props.get("original_film_height", None)
        if self.original_film_width is None or self.original_film_height is None:
            self.film_size = tuple(float(v) for v in self._film.size())
            self.is_confocal = False
        else:
            self.film_size = (float(self.original_film_width), float(self.original_film_height))
            self.is_confocal = True
            if self._film.size() != (1, 1):
                raise ValueError(f"Confocal configuration requires a film with size [1,1] instead of {self._film.size()}")
        self._shape_index, self._shape = -1, None

    def set_shape(self, index, shape):
        self._shape_index, self._shape = index, shape

    def get_shape(self):
        return self._shape

    def film(self):
        return self._film

    def sampler(self):
        return self._sampler

    def world_transform(self):
        return self._world_transform

    def state_key(self):
        return ("nlos_capture_meter", self._world_transform.matrix.tobytes(), self._shape_index, self.is_confocal,
                tuple(self.laser_target), self.film_size)

    def describe(self, scene) -> _abi.TrSensorDesc:
        d = _abi.TrSensorDesc()
        d.type, d.shape = _abi.TR_SENSOR_NLOS_CAPTURE_METER, self._shape_index
        d.origin[:] = [float(x) for x in self._world_transform.translation()]
        d.medium = -1
        d.is_confocal = int(self.is_confocal)
        d.laser_target[:] = [float(x) for x in np.asarray(self.laser_target, dtype=np.float64).reshape(3)]
        return d

    def traverse(self):       # nloscapturemeter.py:219-228
        return {"needs_sample_3": self.needs_sample_3, "is_confocal": self.is_confocal,
                "laser_bounce_opl": self.laser_bounce_opl, "laser_target": self.laser_target,
                "film.size": self._film.size()}
