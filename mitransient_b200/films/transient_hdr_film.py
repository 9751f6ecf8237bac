"""``transient_hdr_film`` -- host mirror of mitransient/films/transient_hdr_film.py.

Same parameter names and defaults (transient_hdr_film.py:114-121), same storage layout
``(H, W, T, C)`` with ``C = len("RGBW")`` (transient_hdr_film.py:173-182,
transient_image_block.py:63-68), same ``develop`` semantics (transient_hdr_film.py:210-248).
Storage lives in CUDA tensors allocated by PyTorch; every arithmetic step is a C-ABI kernel call.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

from .. import _abi, _lib
from .. import colour_mode as _variant
from ..scene import Properties


class TransientHDRFilm:
    plugin_name = "transient_hdr_film"

    def __init__(self, props: Properties):
        # base mi.Film parameters (transient_hdr_film.py:104-111)
        self._width = int(props.get("width", 768))
        self._height = int(props.get("height", 576))
        self._crop_offset = (int(props.get("crop_offset_x", 0)), int(props.get("crop_offset_y", 0)))
        self._crop_size = (int(props.get("crop_width", self._width)), int(props.get("crop_height", self._height)))
        self._sample_border = bool(props.get("sample_border", False))
        rf = props.get("rfilter", {"type": "box"})
        self._rfilter = rf.get("type", "box") if isinstance(rf, dict) else str(rf)
        # transient parameters (transient_hdr_film.py:116-121)
        self.temporal_bins = int(props.get("temporal_bins", 2048))
        self.bin_width_opl = float(props.get("bin_width_opl", 0.003))
        self.start_opl = float(props.get("start_opl", 0.0))
        self.exhaustive_scan = bool(props.get("exhaustive_scan", False))
        self.laser_scan_width = int(props.get("laser_scan_width", 0))
        self.laser_scan_height = int(props.get("laser_scan_height", 0))
        props.get("pixel_format"), props.get("component_format"), props.get("file_format")
        self.channels = list("RGBW")           # no FilmFlags.Alpha on a Python mi.Film subclass
        self.transient_storage = None          # torch.float32 CUDA tensor (H, W, T, C), like TransientImageBlock.tensor
        self._window_storage = None
        self.steady_storage = None             # torch.float32 CUDA tensor (H, W, 3), the developed steady image

    # ---- mi.Film accessors
    def size(self):
        return (self._width, self._height)

    def crop_size(self):
        return self._crop_size

    def crop_offset(self):
        return self._crop_offset

    def sample_border(self):
        return self._sample_border

    def rfilter(self):
        return self._rfilter

    def end_opl(self):                         # transient_hdr_film.py:123-124
        return self.start_opl + self.bin_width_opl * self.temporal_bins

    def base_channels_count(self):
        return 3

    def describe(self) -> _abi.TrFilmDesc:
        d = _abi.TrFilmDesc()
        # the block covers the crop window (transient_hdr_film.py:192-197); the whole film unless one is set
        d.width, d.height, d.temporal_bins = self._crop_size[0], self._crop_size[1], self.temporal_bins
        d.crop_offset_x, d.crop_offset_y = self._crop_offset
        d.start_opl, d.bin_width_opl = self.start_opl, self.bin_width_opl
        d.type = _abi.TR_FILM_TRANSIENT_HDR
        if self.exhaustive_scan:               # transient_image_block.py:59-66
            d.laser_scan_width, d.laser_scan_height = self.laser_scan_width, self.laser_scan_height
        return d

    def state_key(self):
        return ("transient_hdr_film", self.exhaustive_scan, self.laser_scan_width, self.laser_scan_height)

    def raw_channels(self) -> int:
        return len(self.channels)

    def needs_zeroed_accumulate(self) -> bool:
        """exhaustive_scan films are accumulated with global vector reductions (a pixel owns Lw*Lh histograms):
        the kernel adds into the zero-filled storage instead of overwriting it."""
        return self.exhaustive_scan

    def check_supported(self):
        if self._rfilter != "box":
            # the reference's TransientImageBlock rejects everything else (transient_image_block.py:150-151)
            raise ValueError("TransientImageBlock::put_(): using a rfilter but it is not supported "
                             "(only the box reconstruction filter)")
        if self.exhaustive_scan and (self.laser_scan_width <= 0 or self.laser_scan_height <= 0):
            raise ValueError("exhaustive_scan needs laser_scan_width and laser_scan_height > 0")
        cropped = self._crop_offset != (0, 0) or self._crop_size != (self._width, self._height)
        if cropped and self.exhaustive_scan:   # "this type of scanning does not support cropping" (transient_hdr_film.py:87)
            raise ValueError("exhaustive_scan films do not support crop windows")
        if (min(self._crop_offset) < 0 or min(self._crop_size) <= 0 or self._crop_offset[0] + self._crop_size[0] > self._width
                or self._crop_offset[1] + self._crop_size[1] > self._height):
            raise ValueError("Film: the crop window must lie inside the film")

    # ---- lifecycle (transient_hdr_film.py:129-148, 173-202)
    def prepare(self, aovs: Sequence[str] = ()):
        """Re-create zeroed storage, like the reference does on every render (common.py:49)."""
        import torch
        self.check_supported()
        if len(aovs):
            raise NotImplementedError("AOVs are not supported")
        _lib.check(_lib.load().tr_device_check())
        w, h = self.crop_size()
        if self.exhaustive_scan:               # (H, W, Lh, Lw, T, C)   transient_image_block.py:63-66
            shape = (h, w, self.laser_scan_height, self.laser_scan_width, self.temporal_bins, len(self.channels))
        else:
            shape = (h, w, self.temporal_bins, len(self.channels))
        self._window_storage = None
        if self.is_cropped():
            # The reference sizes the transient block from size(), not crop_size(), while giving it the crop OFFSET
            # (transient_hdr_film.py:192-195): the tensor keeps the full film's shape and the rendered window sits in
            # its top-left corner.  The kernel renders the window into a dense tensor of its own; finish_render()
            # pastes it into that corner.  (The steady hdrfilm is sized by the crop window, :131-144.)
            fw, fh = self.size()
            self._window_storage = torch.zeros(shape, dtype=torch.float32, device="cuda")
            self.transient_storage = torch.zeros((fh, fw) + shape[2:], dtype=torch.float32, device="cuda")
        else:
            self.transient_storage = torch.zeros(shape, dtype=torch.float32, device="cuda")
        self.steady_storage = torch.zeros((h, w, 3), dtype=torch.float32, device="cuda")
        return len(self.channels)

    def is_cropped(self) -> bool:
        return self._crop_offset != (0, 0) or self._crop_size != (self._width, self._height)

    def render_target(self):
        """The dense (crop_h, crop_w, T, C) tensor the integrator's kernel writes."""
        return self._window_storage if self._window_storage is not None else self.transient_storage

    def finish_render(self):
        if self._window_storage is not None:
            w, h = self.crop_size()
            self.transient_storage[:h, :w] = self._window_storage

    def clear(self):
        if self.transient_storage is not None:
            self.transient_storage.zero_()
        if getattr(self, "_window_storage", None) is not None:
            self._window_storage.zero_()
        if self.steady_storage is not None:
            self.steady_storage.zero_()

    # ---- K4 on its own (transient_hdr_film.py:250-276)
    def add_transient_data(self, pos, distance, wavelengths, spec, ray_weight=1.0, active=None,
                           laser_x=0, laser_y=0):
        """Scatter a batch of deposits.  ``pos`` (N,2), ``distance`` (N,), ``spec`` (N,3): CUDA float32 tensors."""
        import torch
        if self.transient_storage is None:
            raise RuntimeError("No transient storage allocated, was prepare() called first?")
        if self.exhaustive_scan:
            raise NotImplementedError("add_transient_data on an exhaustive_scan film is only available through render()")
        n = int(distance.shape[0])
        spec = spec * ray_weight
        if active is not None:
            spec = torch.where(active.reshape(-1, 1), spec, torch.zeros_like(spec))
        dep = torch.cat([pos.reshape(n, 2), distance.reshape(n, 1), spec.reshape(n, 3)], dim=1)
        dep = dep.to(device="cuda", dtype=torch.float32).contiguous()
        fd = self.describe()
        fd.width, fd.height = self.size()      # the block is film-sized even under a crop window (see prepare())
        _lib.check(_lib.load().tr_film_scatter_add(self.transient_storage.data_ptr(), C.byref(fd),
                                                   len(self.channels), dep.data_ptr(), n,
                                                   _lib.current_stream_ptr()))

    # ---- develop (transient_hdr_film.py:210-248)
    def develop(self, raw: bool = False):
        steady, transient = self.steady_storage, self.develop_transient_(raw=raw)
        if self.exhaustive_scan:               # the reference's quirk: steady = mean over the CHANNEL axis (:213-214)
            steady = transient.mean(dim=-1)
        if _variant.is_monochromatic():        # "L" instead of "RGB" (transient_hdr_film.py:174, hdrfilm luminance)
            steady = steady[..., :1]
            if not raw:
                transient = transient[..., :1]
        return steady, transient

    def develop_transient_(self, raw: bool = False):
        import torch
        if self.transient_storage is None:
            raise RuntimeError("No transient storage allocated, was prepare_transient_() called first?")
        if raw:
            return self.transient_storage
        data = self.transient_storage
        n_bins = data.numel() // data.shape[-1]
        out = torch.empty(tuple(data.shape[:-1]) + (data.shape[-1] - 1,), dtype=torch.float32, device=data.device)
        _lib.check(_lib.load().tr_film_develop(data.data_ptr(), out.data_ptr(), n_bins, data.shape[-1],
                                               _lib.current_stream_ptr()))
        return out

    def traverse(self):                        # transient_hdr_film.py:295-308
        return {"temporal_bins": self.temporal_bins, "bin_width_opl": self.bin_width_opl,
                "start_opl": self.start_opl, "exhaustive_scan": self.exhaustive_scan,
                "laser_scan_width": self.laser_scan_width, "laser_scan_height": self.laser_scan_height,
                "size": self.size()}

    def to_string(self):
        return ("TransientHDRFilm[\n"
                f"  exhaustive_scan = {self.exhaustive_scan},\n  size = {self.size()},\n"
                f"  crop_size = {self.crop_size()},\n  crop_offset = {self.crop_offset()},\n"
                f"  sample_border = {self.sample_border()},\n  filter = {self.rfilter()},\n"
                f"  temporal_bins = {self.temporal_bins},\n  bin_width_opl = {self.bin_width_opl},\n"
                f"  start_opl = {self.start_opl},\n]")

    __repr__ = to_string
