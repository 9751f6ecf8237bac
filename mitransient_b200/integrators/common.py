"""Host driver of the transient integrators -- mirror of mitransient/integrators/common.py.

``TransientADIntegrator.render`` keeps the reference's signature (common.py:122-129) and control flow
(prepare -> per-pass sample -> develop, common.py:141-213); the per-pass body -- sample_rays, the bounce
loop, both film deposits and the steady splat -- is ONE call into the CUDA library (``tr_render``).
Only the primal render is in scope; render_forward / render_backward (common.py:215-409) are not.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Tuple

from .. import _abi, _lib
from ..sampler import independent_sampler_first_float
from ..scene import Properties, Scene


class TransientADIntegrator:
    integrator_id = -1

    def __init__(self, props: Properties):
        # mitsuba ADIntegrator defaults [upstream]: max_depth 6 (-1 = unbounded), rr_depth 5
        max_depth = int(props.get("max_depth", 6))
        if max_depth < 0 and max_depth != -1:
            raise ValueError('"max_depth" must be set to -1 (infinite) or a value >= 0')
        self.max_depth = max_depth if max_depth != -1 else 0xFFFFFFFF
        self.rr_depth = int(props.get("rr_depth", 5))
        if self.rr_depth <= 0:
            raise ValueError('"rr_depth" must be set to a value greater than zero!')
        self.hide_emitters = bool(props.get("hide_emitters", False))
        self.camera_unwarp = bool(props.get("camera_unwarp", False))              # common.py:25
        self.discard_direct_light = bool(props.get("discard_direct_light", False))  # common.py:27
        props.get("gaussian_stddev", 0.5)      # read and ignored, like the reference (common.py:28-30)
        props.get("temporal_filter", "")
        props.get("block_size", 0)             # documented, never read (transientpath.py:68-71)
        self.last_counters: Optional[dict] = None
        self.collect_counters = False
        # where the kernel accumulates a pixel's histogram: "auto" | "shared" | "global" (TrRenderParams.histogram_path;
        # no counterpart in the reference, not a plugin property -- the film is the same either way)
        self.histogram_path = os.environ.get("MITRANSIENT_B200_HISTOGRAM_PATH", "auto")

    # ---- common.py:32-85
    def prepare(self, scene: Scene, sensor, seed: int, spp: int, aovs=()) -> List[Tuple[int, int]]:
        """Returns ``[(seed_i, spp_i)]`` -- one entry per pass.  One pass up to 2^32 samples, otherwise
        passes of ``int((2^26 - 1) / (W*H))`` spp seeded from the sampler itself (common.py:51-85)."""
        film = sensor.film()
        sampler = sensor.sampler().clone()
        if spp != 0:
            sampler.set_sample_count(spp)
        spp = sampler.sample_count()
        film_size = film.crop_size()
        n_pix = film_size[0] * film_size[1]
        wavefront_size = n_pix * spp
        film.prepare(aovs)
        seed = (int(seed) + sampler.base_seed) & 0xFFFFFFFF
        if wavefront_size <= 2 ** 32:
            # lane indices are 32-bit: exactly 2^32 would overflow; the C ABI checks
            return [(seed, spp)]
        spp_per_pass = int((2 ** 26 - 1) / n_pix)
        if spp_per_pass == 0:
            raise Exception("Your film is too big. Please make it smaller.")
        needs_remainder = spp % spp_per_pass != 0
        num_passes = spp // spp_per_pass + 1 * needs_remainder
        passes = []
        for i in range(num_passes):
            u = independent_sampler_first_float(seed, i)
            # common.py:72 draws the seed, :80 hands it to a fresh clone of the sensor's sampler, whose seed() adds the
            # sampler's own `seed` property again (independent.cpp) -- like the single-pass seed above
            seed_i = (int(u * 2.0 ** 32) + sampler.base_seed) & 0xFFFFFFFF
            spp_i = spp % spp_per_pass if (needs_remainder and i == num_passes - 1) else spp_per_pass
            passes.append((seed_i, spp_i))
        return passes

    def check_transient_(self, scene: Scene, sensor):                               # common.py:424-447
        from ..films.transient_hdr_film import TransientHDRFilm
        from ..sensors.nloscapturemeter import NLOSCaptureMeter
        if isinstance(sensor, NLOSCaptureMeter) and self.camera_unwarp:
            raise AssertionError("camera_unwarp is not supported for NLOSCaptureMeter. "
                                 "Use account_first_and_last_bounces in the NLOSCaptureMeter plugin instead.")
        if not isinstance(sensor.film(), TransientHDRFilm):
            raise AssertionError("The film of the sensor must be of type transient_hdr_film or phasor_hdr_film")

    # ---- flags / params handed to the kernel
    def render_flags(self) -> int:
        f = 0
        if self.camera_unwarp:
            f |= _abi.TR_FLAG_CAMERA_UNWARP
        if self.discard_direct_light:
            f |= _abi.TR_FLAG_DISCARD_DIRECT_LIGHT
        if self.hide_emitters:
            f |= _abi.TR_FLAG_HIDE_EMITTERS
        return f

    def make_params(self, spp_i: int, total_spp: int, seed_i: int, rank: int = 0, world: int = 1,
                    film_channels: int = 4, accumulate: bool = False, tile_size: int = 16) -> _abi.TrRenderParams:
        p = _abi.TrRenderParams()
        p.integrator = self.integrator_id
        p.max_depth, p.rr_depth, p.flags = self.max_depth, self.rr_depth, self.render_flags()
        p.spp, p.total_spp, p.seed = spp_i, total_spp, seed_i
        p.tile_size, p.rank, p.world = tile_size, rank, world
        p.film_channels, p.accumulate = film_channels, int(accumulate)
        p.capture_type, p.filter_depth, p.force_equal_grids = 0, -1, 1
        if self.histogram_path not in _abi.HIST_NAMES:
            raise ValueError(f'histogram_path must be one of {sorted(_abi.HIST_NAMES)}, not {self.histogram_path!r}')
        p.histogram_path = _abi.HIST_NAMES[self.histogram_path]
        return p

    @staticmethod
    def last_render_info() -> dict:
        """What the last kernel launch of this thread used (tr_last_render_info): histogram path, slots, threads."""
        info = _abi.TrRenderInfo()
        _lib.load().tr_last_render_info(info)
        return {"histogram_path": {0: None, _abi.TR_HIST_SHARED: "shared", _abi.TR_HIST_GLOBAL: "global"}[info.histogram_path],
                "slots": int(info.slots), "threads": int(info.threads), "smem_bytes": int(info.smem_bytes)}

    def device_scene(self, scene: Scene, sensor) -> _lib.DeviceScene:
        key = scene.description_key(sensor)
        if scene._handle is None or scene._handle_key != key:
            if scene._handle is not None:
                scene._handle.close()
            scene._handle = _lib.DeviceScene(scene.describe(sensor))
            scene._handle_key = key
        return scene._handle

    # ---- common.py:122-213
    def render(self, scene: Scene, sensor=0, seed: int = 0, spp: int = 0, develop: bool = True,
               evaluate: bool = True, progress_callback=None, rank: int = 0, world: int = 1):
        """Returns ``(steady (H,W,3), transient (H,W,T,3))`` CUDA float32 tensors.

        ``rank``/``world`` select this process's pixel tiles; the caller reduces the film."""
        import torch
        if not develop:
            raise Exception("develop=True must be specified when invoking AD integrators")
        if isinstance(sensor, int):
            sensor = scene.sensors()[sensor]
        film = sensor.film()
        self.check_transient_(scene, sensor)
        self.check_scene_(scene, sensor)
        passes = self.prepare(scene, sensor, seed, spp, aovs=())
        total_spp = sum(s for _, s in passes)
        dev = self.device_scene(scene, sensor)
        lib = _lib.load()
        counters = torch.zeros(8, dtype=torch.int64, device="cuda") if self.collect_counters else None
        for i, (seed_i, spp_i) in enumerate(passes):
            params = self.make_params(spp_i, total_spp, seed_i, rank=rank, world=world,
                                      film_channels=film.raw_channels(),
                                      accumulate=(i > 0) or film.needs_zeroed_accumulate())
            self.fill_params_(params, scene, sensor)
            _lib.check(lib.tr_render(dev.handle, C.byref(params), film.render_target().data_ptr(),
                                     film.steady_storage.data_ptr(),
                                     counters.data_ptr() if counters is not None else None,
                                     _lib.current_stream_ptr()))
            if progress_callback:
                progress_callback((i + 1) / len(passes))
        if counters is not None:
            c = counters.cpu().tolist()
            self.last_counters = dict(zip(("paths", "vertices", "deposits", "deposits_in_range", "bin_checksum",
                                           "shadow_rays"), c[:6]))
        film.finish_render()
        steady, transient = film.develop()
        if evaluate:
            torch.cuda.current_stream().synchronize()
        return steady, transient

    def render_to_host(self, scene: Scene, sensor=0, seed: int = 0, spp: int = 0, raw: bool = False,
                       out_transient=None, out_steady=None, rank: int = 0, world: int = 1, tile_size: int = 16):
        """Render straight into HOST (ideally pinned) tensors with the device->host copy of the film
        overlapped with rendering (``tr_render_to_host``).  Returns ``(steady, transient)`` CPU tensors:
        transient is (H,W,T,3) developed, or (H,W,T,4) raw.

        ``world > 1``: this rank renders only the pixels it owns and the outputs are the COMPACT arrays
        ``(n_owned, T, C)`` / ``(n_owned, 3)`` in ``distributed.owned_pixels(...)`` order, so that every GPU of a
        node moves only its share of the film over its own PCIe link."""
        import torch
        if isinstance(sensor, int):
            sensor = scene.sensors()[sensor]
        film = sensor.film()
        self.check_transient_(scene, sensor)
        self.check_scene_(scene, sensor)
        film.check_supported()
        if film.raw_channels() != 4 or film.needs_zeroed_accumulate() or film.is_cropped():
            raise NotImplementedError("render_to_host supports plain transient_hdr_film films without a crop window")
        s = sensor.sampler().clone()
        if spp != 0:
            s.set_sample_count(spp)
        spp = s.sample_count()
        w, h = film.crop_size()
        ch = 4 if raw else 3
        if world > 1:
            from ..distributed import owned_pixels
            n_owned = len(owned_pixels(w, h, tile_size, rank, world))
            shape_t, shape_s = (n_owned, film.temporal_bins, ch), (n_owned, 3)
        else:
            shape_t, shape_s = (h, w, film.temporal_bins, ch), (h, w, 3)
        if out_transient is None:
            out_transient = torch.empty(shape_t, dtype=torch.float32, pin_memory=True)
        if out_steady is None:
            out_steady = torch.empty(shape_s, dtype=torch.float32, pin_memory=True)
        if tuple(out_transient.shape) != shape_t or tuple(out_steady.shape) != shape_s:
            raise ValueError(f"output tensors must have shapes {shape_t} and {shape_s}")
        dev = self.device_scene(scene, sensor)
        params = self.make_params(spp, spp, (int(seed) + s.base_seed) & 0xFFFFFFFF, film_channels=ch, rank=rank,
                                  world=world, tile_size=tile_size)
        self.fill_params_(params, scene, sensor)
        cnt = _abi.TrCounters() if self.collect_counters else None
        _lib.check(_lib.load().tr_render_to_host(dev.handle, C.byref(params), out_transient.data_ptr(),
                                                 out_steady.data_ptr(), C.byref(cnt) if cnt is not None else None,
                                                 _lib.current_stream_ptr()))
        self.last_counters = cnt.as_dict() if cnt is not None else None
        return out_steady, out_transient

    # hooks for subclasses
    def check_scene_(self, scene: Scene, sensor):
        pass

    def fill_params_(self, params: _abi.TrRenderParams, scene: Scene, sensor):
        pass

    def aov_names(self):
        return []

    def to_string(self):                                                            # common.py:449-454
        md = -1 if self.max_depth == 0xFFFFFFFF else self.max_depth
        return f"{type(self).__name__}[\n  max_depth = {md}, \n  rr_depth = {self.rr_depth}\n]"

    __repr__ = to_string
