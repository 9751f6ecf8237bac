"""``phasor_hdr_film`` -- host mirror of mitransient/films/phasor_hdr_film.py.

Stores the transient signal in the frequency domain for a sparse set of frequencies chosen by a Morlet-wavelet
band (``wl_mean``, ``wl_sigma``; phasor_hdr_film.py:112-137).  Storage is ``(H, W, 2F + 1)``: (Re, Im) per
frequency plus the weight channel (phasor_hdr_film.py:162-176); a deposit adds
``spec.x * (cos, sin)(fmod(-2 pi f (opl - start_opl), 2 pi))`` for every frequency
(render/phasor_image_block.py:43-67) -- done inside the render kernel, 32 frequencies per warp step.
``develop()`` returns ``(steady, phasors)`` with phasors ``(H, W, F, 2)`` (phasor_hdr_film.py:210-239).
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np

from .. import _abi, _lib
from .. import colour_mode as _variant
from ..scene import Properties
from .transient_hdr_film import TransientHDRFilm


class PhasorHDRFilm(TransientHDRFilm):
    plugin_name = "phasor_hdr_film"

    def __init__(self, props: Properties):
        self.wl_mean = float(props.get("wl_mean", 100.0))                          # phasor_hdr_film.py:112-116
        self.wl_sigma = float(props.get("wl_sigma", 1000.0))
        if "temporal_bins" not in props:
            props = Properties({**props, "temporal_bins": 4096})
        super().__init__(props)
        if self.exhaustive_scan:
            raise ValueError("phasor_hdr_film has no exhaustive_scan layout")
        if self._crop_size != (self._width, self._height):                           # :118-123
            raise ValueError("PhasorHDRFilm: crop_size must match size")
        if self._crop_offset != (0, 0):
            raise ValueError("PhasorHDRFilm: crop_offset must be (0, 0)")
        if self._sample_border:
            raise ValueError("PhasorHDRFilm: sample_border must be False")
        nt = self.temporal_bins                                                      # :125-137
        mean_idx = (nt * self.bin_width_opl) / self.wl_mean
        sigma_idx = (nt * self.bin_width_opl) / (self.wl_sigma * 6)
        freq_min_idx = np.maximum(0, int(np.floor(mean_idx - 3 * sigma_idx)))
        freq_max_idx = np.minimum(nt // 2, int(np.ceil(mean_idx + 3 * sigma_idx)))
        self.frequencies = np.ascontiguousarray(
            np.fft.fftfreq(nt, d=self.bin_width_opl)[freq_min_idx:freq_max_idx + 1].astype(np.float32))
        if len(self.frequencies) == 0:
            raise ValueError("PhasorHDRFilm: the (wl_mean, wl_sigma) band selects no frequency")
        self.channels = [f"L_fq{j:03d}_{p}" for j in range(len(self.frequencies)) for p in ("Re", "Im")] + ["W"]
        self.phasor_storage = None

    def describe(self) -> _abi.TrFilmDesc:
        d = super().describe()
        d.type = _abi.TR_FILM_PHASOR_HDR
        d.n_frequencies = len(self.frequencies)
        d.frequencies = self.frequencies.ctypes.data_as(C.POINTER(C.c_float))
        return d

    def state_key(self):
        return ("phasor_hdr_film", self.frequencies.tobytes())

    def raw_channels(self) -> int:
        return 2 * len(self.frequencies) + 1

    def needs_zeroed_accumulate(self) -> bool:
        return False

    def prepare(self, aovs: Sequence[str] = ()):                                     # :146-188
        import torch
        if not _variant.is_monochromatic():
            raise ValueError("PhasorHDRFilm: Only monochromatic rendering supported "
                             "(call mitransient_b200.set_variant('cuda_mono') before loading the scene)")
        if len(aovs):
            raise ValueError("PhasorHDRFilm: AOVs not supported")
        self.check_supported()
        _lib.check(_lib.load().tr_device_check())
        w, h = self.size()
        self.phasor_storage = torch.zeros((h, w, self.raw_channels()), dtype=torch.float32, device="cuda")
        self.transient_storage = self.phasor_storage       # what the integrator renders into
        self.steady_storage = torch.zeros((h, w, 3), dtype=torch.float32, device="cuda")
        return len(self.channels)

    def create_block(self):
        raise NotImplementedError("Not implemented for phasor_hdr_film")           # :140-141

    def gather_derivatives_at_distance(self, pos, dL, distance):
        raise NotImplementedError("Not implemented for phasor_hdr_film")           # :143-144

    def add_transient_data(self, *args, **kwargs):
        raise NotImplementedError("phasor deposits are accumulated inside the render kernel (tr_render)")

    def develop(self, raw: bool = False):                                            # :204-208
        return self.steady_storage[..., :1], self.develop_phasors_(raw=raw)

    def develop_phasors_(self, raw: bool = False):                                   # :210-239
        import torch
        if self.phasor_storage is None:
            raise RuntimeError("No phasor storage allocated, was prepare() called first?")
        data = self.phasor_storage
        if raw:
            return data
        n_pix = data.numel() // data.shape[-1]
        out = torch.empty(tuple(data.shape[:-1]) + (data.shape[-1] - 1,), dtype=torch.float32, device=data.device)
        _lib.check(_lib.load().tr_film_develop(data.data_ptr(), out.data_ptr(), n_pix, data.shape[-1],
                                               _lib.current_stream_ptr()))
        return out.reshape(tuple(data.shape[:-1]) + (len(self.frequencies), 2))

    def traverse(self):                                                              # :273-279
        return {"frequencies": self.frequencies, "start_opl": self.start_opl, "size": self.size()}

    def to_string(self):                                                             # :266-271
        return (f"PhasorHDRFilm[\n  size = {self.size()},\n  frequencies = {self.frequencies},\n"
                f"  start_opl = {self.start_opl},\n]")

    __repr__ = to_string
