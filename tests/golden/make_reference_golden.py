"""Generate tests/golden/reference_host_golden.npz / .json by EXECUTING the reference's own host-side Python
(/root/reference/mitransient, unmodified) over the NumPy stand-in of refshim.py.

    python tests/golden/make_reference_golden.py          (build container only: needs /root/reference)

These vectors pin the film half of the hot path to the reference's code rather than to our reading of it:
rows a5-a8 and a10 of SURVEY section 8 (add_transient_data, TransientImageBlock.put_/accum incl. the exhaustive_scan
index, develop), row f4 (phasor film: frequency selection, PhasorImageBlock.put, develop_phasors_), the benchmark
scene dictionary (utils.cornell_box), the integrators' parameter parsing and pass split, and vis.tonemap_transient.
The ray-tracing half (Mitsuba C++) cannot be executed here and stays unpinned.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import refshim  # noqa: E402

import mitransient_b200 as mitr  # noqa: E402  (only ScalarTransform4f / ScalarColor3d, handed to the stand-in)

out, meta = {}, {}


def deposits(rng, n, W, H, lo, hi, start, bw, T):
    """x, y integer-valued; opl spans below / inside / above the window and includes the exact edges, inf, nan."""
    x = rng.integers(0, W, n).astype(np.float32)
    y = rng.integers(0, H, n).astype(np.float32)
    opl = (lo + (hi - lo) * rng.random(n)).astype(np.float32)
    end = np.float32(start) + np.float32(bw) * np.float32(T)
    special = [np.float32(start), np.nextafter(np.float32(start), np.float32(-1e9)), np.nextafter(end, np.float32(-1e9)), end,
               np.float32(np.inf), np.float32(np.nan), np.float32(start) + np.float32(bw), np.float32(0.0)]
    opl[: len(special)] = special
    rgb = rng.random((n, 3)).astype(np.float32)
    rgb[rng.random(n) < 0.1] = 0.0
    return np.concatenate([x[:, None], y[:, None], opl[:, None], rgb], axis=1)


# ------------------------------------------------------------------ A / B: transient_hdr_film (plain + exhaustive_scan)
mi, dr = refshim.install(False, mitr.ScalarTransform4f, mitr.ScalarColor3d)
thf = refshim.ref_module("mitransient.films.transient_hdr_film")
rng = np.random.default_rng(20261017)
for tag, props, lo, hi in (("film_a", dict(width=7, height=5, temporal_bins=32, start_opl=1.25, bin_width_opl=0.37), 0.0, 15.0),
                           ("film_b", dict(width=3, height=4, temporal_bins=1000, start_opl=0.0, bin_width_opl=0.003), 0.0, 3.3),
                           ("film_c", dict(width=1, height=1, temporal_bins=1, start_opl=2.0, bin_width_opl=5.0), 0.0, 9.0)):
    film = thf.TransientHDRFilm(mi.Properties(**props))
    film.prepare([])
    assert film.channels == ["R", "G", "B", "W"]
    dep = deposits(rng, 6000, props["width"], props["height"], lo, hi, props["start_opl"], props["bin_width_opl"],
                   props["temporal_bins"])
    film.add_transient_data(mi.Vector2f(dep[:, 0], dep[:, 1]), mi.Float(dep[:, 2]), None,
                            mi.Spectrum(dep[:, 3], dep[:, 4], dep[:, 5]), mi.Float(1.0), mi.Bool(np.ones(len(dep), dtype=bool)))
    out[tag + "_deposits"] = dep
    out[tag + "_raw"] = film.develop_transient_(raw=True).numpy()
    out[tag + "_developed"] = film.develop_transient_(raw=False).numpy()
    meta[tag] = props

props = dict(width=4, height=3, temporal_bins=16, start_opl=0.5, bin_width_opl=0.25, exhaustive_scan=True,
             laser_scan_width=3, laser_scan_height=2)            # non-square laser grid: pins the index formula as written
film = thf.TransientHDRFilm(mi.Properties(**props))
film.prepare([])
dep = deposits(rng, 4000, 4, 3, 0.0, 5.0, 0.5, 0.25, 16)
lx = rng.integers(0, 3, len(dep)); ly = rng.integers(0, 2, len(dep))
film.add_transient_data(mi.Vector2f(dep[:, 0], dep[:, 1]), mi.Float(dep[:, 2]), None, mi.Spectrum(dep[:, 3], dep[:, 4], dep[:, 5]),
                        mi.Float(1.0), mi.Bool(np.ones(len(dep), dtype=bool)), laser_x=mi.UInt32(lx), laser_y=mi.UInt32(ly))
raw = film.develop_transient_(raw=True)
assert raw.shape == (3, 4, 2, 3, 16, 4)                          # (H, W, Lh, Lw, T, C)   transient_image_block.py:63-66
out["film_exh_deposits"], out["film_exh_laser"] = dep, np.stack([lx, ly], axis=1).astype(np.uint32)
out["film_exh_raw"] = raw.numpy()
meta["film_exh"] = props

# a crop window: the transient block keeps the FULL film's size but takes the crop offset (transient_hdr_film.py:192-195),
# so deposits land at (pos - offset) inside a film-sized tensor and positions left of / above the window are dropped
props = dict(width=9, height=7, temporal_bins=16, start_opl=0.5, bin_width_opl=0.4, crop_width=4, crop_height=3,
             crop_offset_x=3, crop_offset_y=2)
film = thf.TransientHDRFilm(mi.Properties(**props))
film.prepare([])
dep = deposits(np.random.default_rng(77), 5000, 9, 7, 0.0, 8.0, 0.5, 0.4, 16)
film.add_transient_data(mi.Vector2f(dep[:, 0], dep[:, 1]), mi.Float(dep[:, 2]), None, mi.Spectrum(dep[:, 3], dep[:, 4], dep[:, 5]),
                        mi.Float(1.0), mi.Bool(np.ones(len(dep), dtype=bool)))
out["film_d_deposits"], out["film_d_raw"] = dep, film.develop_transient_(raw=True).numpy()
out["film_d_developed"] = film.develop_transient_(raw=False).numpy()
assert out["film_d_raw"].shape == (7, 9, 16, 4) and np.all(out["film_d_raw"][:, 6:] == 0) and np.all(out["film_d_raw"][5:] == 0)
meta["film_d"] = props

# ------------------------------------------------------------------ C: phasor_hdr_film (monochromatic variant)
mi, dr = refshim.install(True, mitr.ScalarTransform4f, mitr.ScalarColor3d)
phf = refshim.ref_module("mitransient.films.phasor_hdr_film")
meta["phasor"] = {}
for tag, props in (("phasor_a", dict(width=5, height=4, temporal_bins=400, start_opl=1000.0, bin_width_opl=6.5, wl_mean=100.0, wl_sigma=100.0)),
                   ("phasor_b", dict(width=3, height=3, temporal_bins=256, start_opl=3.5, bin_width_opl=6.0 / 256, wl_mean=0.5, wl_sigma=0.5)),
                   ("phasor_c", dict(width=2, height=2, temporal_bins=4096, start_opl=0.0, bin_width_opl=0.003))):
    film = phf.PhasorHDRFilm(mi.Properties(**props))
    film.prepare([])
    freqs = np.array([float(f) for f in film.frequencies], dtype=np.float32)
    lo, hi = props["start_opl"], props["start_opl"] + props["bin_width_opl"] * props["temporal_bins"]
    dep = deposits(rng, 1500, props["width"], props["height"], lo, hi, lo, props["bin_width_opl"], props["temporal_bins"])
    dep[:, 4] = dep[:, 3]; dep[:, 5] = dep[:, 3]               # monochromatic: one value per deposit
    film.add_transient_data(mi.Vector2f(dep[:, 0], dep[:, 1]), mi.Float(dep[:, 2]), None,
                            mi.Spectrum(dep[:, 3], dep[:, 4], dep[:, 5]), mi.Float(1.0), mi.Bool(np.ones(len(dep), dtype=bool)))
    out[tag + "_frequencies"], out[tag + "_deposits"] = freqs, dep
    out[tag + "_raw"] = film.develop_phasors_(raw=True).numpy()
    out[tag + "_developed"] = film.develop_phasors_(raw=False).numpy()
    meta["phasor"][tag] = {**props, "channels": film.channels[:2] + film.channels[-1:], "n_channels": len(film.channels)}

# ------------------------------------------------------------------ D: the benchmark scene dictionary
mi, dr = refshim.install(False, mitr.ScalarTransform4f, mitr.ScalarColor3d)
utils = refshim.ref_module("mitransient.utils")


def jsonable(v):
    if isinstance(v, dict):
        return {k: jsonable(x) for k, x in v.items()}
    if isinstance(v, mitr.ScalarTransform4f):
        return {"__transform__": np.asarray(v.matrix, dtype=np.float64).tolist()}
    if isinstance(v, np.ndarray):
        return {"__array__": v.astype(np.float64).tolist()}
    if isinstance(v, (list, tuple)):
        return [jsonable(x) for x in v]
    return v


meta["cornell_box"] = jsonable(utils.cornell_box())
meta["speed_of_light"] = utils.speed_of_light

# ------------------------------------------------------------------ E: integrator parameter parsing
nlos = refshim.ref_module("mitransient.integrators.transientnlospath")
tpath = refshim.ref_module("mitransient.integrators.transientpath")
cases = [{}, {"capture_type": 2}, {"capture_type": "confocal"}, {"capture_type": "exhaustive", "force_equal_illumination_scanning": False},
         {"filter_bounces": 2}, {"filter_depth": 3, "max_depth": 12}, {"nlos_hidden_geometry_sampling_do_rroulette": True},
         {"nlos_hidden_geometry_sampling": True, "nlos_hidden_geometry_sampling_do_rroulette": True,
          "nlos_hidden_geometry_sampling_includes_relay_wall": True, "nlos_laser_sampling": True},
         {"max_depth": -1, "rr_depth": 3, "discard_direct_paths": True, "account_first_and_last_bounces": True,
          "illumination_scan_fov": 35.0}]
attrs = ["max_depth", "rr_depth", "camera_unwarp", "discard_direct_light", "filter_depth", "filter_bounces", "discard_direct_paths",
         "force_equal_grids", "illumination_scan_fov", "laser_sampling", "hg_sampling", "hg_sampling_do_rroulette",
         "hg_sampling_includes_relay_wall", "account_first_and_last_bounces"]
meta["nlos_params"] = []
for c in cases:
    it = nlos.TransientNLOSPath(mi.Properties(**c))
    rec = {a: (float(getattr(it, a)) if isinstance(getattr(it, a), float) else int(getattr(it, a))) for a in attrs}
    rec["capture_type"] = int(it.capture_type)
    meta["nlos_params"].append({"props": c, "attrs": rec})
try:
    nlos.TransientNLOSPath(mi.Properties(filter_depth=2, filter_bounces=1))
    meta["nlos_both_filters_error"] = None
except AssertionError as e:
    meta["nlos_both_filters_error"] = str(e)
it = tpath.TransientPath(mi.Properties(max_depth=8, camera_unwarp=True, gaussian_stddev=2.0, temporal_filter="box"))
meta["path_params"] = {"max_depth": int(it.max_depth), "rr_depth": int(it.rr_depth), "camera_unwarp": bool(it.camera_unwarp),
                       "discard_direct_light": bool(it.discard_direct_light)}

# ------------------------------------------------------------------ G: TransientADIntegrator.prepare pass split


class _Sampler:
    def __init__(self, spp): self.spp, self.log = spp, []
    def clone(self): s = _Sampler(self.spp); s.log = self.log; return s
    def set_sample_count(self, n): self.spp = int(n)
    def sample_count(self): return self.spp
    def set_samples_per_wavefront(self, n): pass
    def seed(self, seed, wavefront): self.log.append(int(wavefront))
    def next_1d(self): return np.linspace(0.1, 0.9, self.spp).astype(np.float32)      # one value per pass


class _Sensor:
    def __init__(self, film, spp): self._f, self._s = film, _Sampler(spp)
    def film(self): return self._f
    def sampler(self): return self._s


meta["prepare_split"] = []
for (w, h, spp_film, spp_arg) in ((64, 64, 16, 0), (512, 512, 256, 0), (512, 512, 4, 4096), (512, 512, 4, 2 ** 15), (1024, 1024, 1, 5000),
                                 (300, 200, 1, 80000)):
    film = thf.TransientHDRFilm(mi.Properties(width=w, height=h, temporal_bins=2))
    sensor = _Sensor(film, spp_film)
    passes = tpath.TransientPath(mi.Properties()).prepare(None, sensor, 0, spp_arg, [])
    meta["prepare_split"].append({"width": w, "height": h, "film_spp": spp_film, "spp": spp_arg,
                                  "passes": [int(s) for _, s in passes]})

# ------------------------------------------------------------------ F: vis.tonemap_transient
vis = refshim.ref_module("mitransient.unpolarized_visualization")
x = (np.random.default_rng(5).standard_normal((6, 5, 7, 3)) ** 3).astype(np.float32)
out["tonemap_in"], out["tonemap_out"] = x, vis.tonemap_transient(x, scaling=2.0).astype(np.float32)

# ------------------------------------------------------------------ H: the integrator loop itself
# The reference's render() (integrators/common.py:122-213) and TransientPath.sample() (transientpath.py:88-326),
# UNMODIFIED, executed lane-wise over refshim's scene services (whose Mitsuba primitives are the oracle's osv_*
# exports).  Stored: every non-zero deposit (x, y, opl, r, g, b), sorted; the raw film; the steady image.
from path_cases import (ANGULAR_CASES, NLOS_CASES, PATH_CASES, PHASOR_CASES, VOL_CASES, angular_case_scene,  # noqa: E402
                        nlos_case_scene, path_case_scene, phasor_case_scene, sort_rows, vol_case_scene)

if __name__ == "__main__":
    path_out = {}
    for name in PATH_CASES:
        scene, cfg, props = path_case_scene(name)
        steady, raw, dep = refshim.reference_render(scene, "mitransient.integrators.transientpath", "TransientPath", props,
                                                    spp=cfg[3], seed=3)
        path_out[name + "_deposits"], path_out[name + "_raw"], path_out[name + "_steady"] = sort_rows(dep), raw, steady.astype(np.float32)
        print(name, cfg, props, "deposits:", len(dep))
    # ---- I: transient_nlos_path: render() + TransientNLOSPath.prepare / sample / emitter_laser_sample /
    # emitter_laser_targets_sample / emitter_nee_sample / hidden_geometry_sample (transientnlospath.py:251-927) and the
    # NLOSCaptureMeter sensor (nloscapturemeter.py:136-202), all executed unmodified.
    for name in NLOS_CASES:
        scene, cfg, props = nlos_case_scene(name)
        steady, raw, dep = refshim.reference_render_nlos(scene, props, spp=cfg[3], seed=5)
        path_out[name + "_deposits"], path_out[name + "_raw"] = sort_rows(dep), raw
        print(name, cfg, props.get("capture_type"), "deposits:", len(dep))
    # ---- J: mitransient's own emitter: AngularAreaLight.sample_direction / pdf_direction / eval (emitters/angulararea.py)
    # executed inside the reference's transient_path loop
    for name in ANGULAR_CASES:
        scene, cfg, props = angular_case_scene(name)
        steady, raw, dep = refshim.reference_render_angular(scene, props, spp=cfg[3], seed=2)
        path_out[name + "_deposits"], path_out[name + "_raw"], path_out[name + "_steady"] = sort_rows(dep), raw, steady.astype(np.float32)
        print(name, cfg, "deposits:", len(dep))
    # ---- K: transient_prbvolpath: render() + TransientPRBVolpathIntegrator.sample / sample_emitter / first_surface
    # (transient_prbvolpath.py:114-528) over homogeneous media with HG phase functions behind null boundaries
    for name in VOL_CASES:
        scene, cfg, props = vol_case_scene(name)
        steady, raw, dep = refshim.reference_render_vol(scene, props, spp=cfg[3], seed=7)
        path_out[name + "_deposits"], path_out[name + "_raw"], path_out[name + "_steady"] = sort_rows(dep), raw, steady.astype(np.float32)
        print(name, cfg, props, "deposits:", len(dep))
    # ---- L: the reference's PhasorHDRFilm filled by the reference's transient_path loop (monochromatic variant)
    mitr.set_variant("llvm_ad_mono")
    for name in PHASOR_CASES:
        scene, cfg, props = phasor_case_scene(name)
        steady, raw, dep = refshim.reference_render(scene, "mitransient.integrators.transientpath", "TransientPath", props, spp=cfg[3],
                                                    seed=3, monochromatic=True)
        path_out[name + "_raw"], path_out[name + "_steady"] = raw, steady.astype(np.float32)
        print(name, cfg, "phasor film", raw.shape, "deposits:", len(dep))
    mitr.set_variant("cuda_rgb")
    np.savez_compressed(os.path.join(HERE, "reference_path_golden.npz"), **path_out)

np.savez_compressed(os.path.join(HERE, "reference_host_golden.npz"), **out)
with open(os.path.join(HERE, "reference_host_golden.json"), "w") as f:
    json.dump(meta, f, indent=1, sort_keys=True)
print({k: v.shape for k, v in out.items()})
print("passes:", [(c["width"], c["spp"] or c["film_spp"], len(c["passes"])) for c in meta["prepare_split"]])
