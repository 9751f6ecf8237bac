"""GPU parity of transient_nlos_path + nlos_capture_meter against the CPU oracle (same tolerances and the same
exact-work-counter requirement as tests/test_gpu_parity.py)."""
import numpy as np
import pytest

from parity_utils import film_l1

pytestmark = pytest.mark.gpu
WHOLE_TOL, PER_BIN_TOL = 1e-5, 1e-4


def _both(mitr, oracle, d, spp, seed=0, focus_pixel=None):
    scene = mitr.load_dict(d)
    if focus_pixel is not None:     # single capture: aim the laser (mitransient/nlos.py)
        em = scene.emitters()[0]["spec"]
        wall = [v for v in d.values() if isinstance(v, dict) and v.get("type") == "rectangle"][0]
        mitr.nlos.focus_emitter_at_relay_wall_pixel(focus_pixel, mitr.PluginSpec(wall), em)
    integ = scene.integrator()
    integ.collect_counters = True
    steady, transient = integ.render(scene, seed=seed, spp=spp)
    raw = scene.sensors()[0].film().develop(raw=True)[1].cpu().numpy()
    ref = oracle.render(scene, integ, seed=seed, spp=spp)
    assert integ.last_counters == ref["counters"], (integ.last_counters, ref["counters"])
    assert ref["counters"]["deposits_in_range"] > 0
    err = film_l1(raw, ref["transient"])
    assert err["zero_bins_mismatch"] == 0 and err["whole"] <= WHOLE_TOL and err["max_t"] <= PER_BIN_TOL, err
    np.testing.assert_allclose(steady.cpu().numpy(), ref["steady"], rtol=1e-4, atol=1e-6)
    return raw, ref


def test_confocal_laser_hg_sampling(mitr, oracle):
    """BASELINE config C3 at test size: confocal capture, laser + hidden-geometry sampling."""
    raw, ref = _both(mitr, oracle, mitr.nlos_z_scene(32, 32, 300, 64), 64)
    # KAT-6 time of flight: hidden plane at distance 1 -> the first light returns at OPL = 2 (first and last
    # segments are not accounted): first non-zero bin = floor((2 - 1.85) / 0.006) = 25
    prof = raw[..., 0].sum(axis=(0, 1))
    assert np.nonzero(prof)[0][0] in (24, 25)


@pytest.mark.parametrize("integ,spp,seed", [
    ({"account_first_and_last_bounces": True}, 32, 1),
    ({"nlos_hidden_geometry_sampling_do_rroulette": True, "max_depth": 4}, 48, 2),
    ({"nlos_hidden_geometry_sampling": False, "max_depth": 5}, 64, 3),
    ({"nlos_hidden_geometry_sampling_includes_relay_wall": True, "max_depth": 5}, 32, 4),
    ({"filter_depth": 3}, 32, 5),
    ({"discard_direct_paths": True, "max_depth": 6}, 32, 6),
])
def test_confocal_variants(mitr, oracle, integ, spp, seed):
    d = mitr.nlos_z_scene(16, 24, 512, spp, start_opl=0.0, window=6.0, **integ)
    _both(mitr, oracle, d, spp, seed=seed)


def test_single_capture_with_focused_laser(mitr, oracle):
    d = mitr.nlos_z_scene(24, 24, 300, 64, capture_type="single")
    _both(mitr, oracle, d, 64, focus_pixel=(12, 12))


def test_single_capture_plain_nee(mitr, oracle):
    """No laser sampling: the projector is sampled directly from every vertex (emitter_nee_sample)."""
    d = mitr.nlos_z_scene(16, 16, 400, 256, start_opl=0.0, window=8.0, capture_type="single",
                          nlos_laser_sampling=False, account_first_and_last_bounces=True)
    d["emitter"]["fov"] = 60.0          # a wide laser so that plain NEE actually finds it
    _both(mitr, oracle, d, 256, focus_pixel=(8, 8))


def test_nlos_rejections(mitr):
    d = mitr.nlos_z_scene(8, 8, 16, 1, camera_unwarp=True)
    scene = mitr.load_dict(d)
    with pytest.raises(AssertionError, match="camera_unwarp"):
        scene.integrator().render(scene, spp=1)
    with pytest.raises(AssertionError, match="Only use one"):
        mitr.load_dict({"type": "transient_nlos_path", "filter_depth": 2, "filter_bounces": 1})
    d = mitr.nlos_z_scene(8, 8, 16, 1, capture_type="exhaustive")     # Exhaustive needs an exhaustive_scan film
    scene = mitr.load_dict(d)
    with pytest.raises(ValueError, match="exhaustive_scan"):
        scene.integrator().render(scene, spp=1)


@pytest.mark.parametrize("capture", ["single", "confocal"])
def test_nlos_path_with_a_perspective_camera(mitr, oracle, capture):
    """transient_nlos_path looking at the relay wall through an ordinary perspective camera (the setup of
    examples/transient-nlos/nlos-z-simple.xml) instead of an nlos_capture_meter."""
    T = mitr.ScalarTransform4f
    v, f = mitr.letter_z_mesh()
    cam = T().look_at(origin=[-1.2, 0.0, 1.6], target=[0.0, 0.0, 0.0], up=[0, 1, 0])
    d = {"type": "scene",
         "integrator": {"type": "transient_nlos_path", "nlos_laser_sampling": True, "nlos_hidden_geometry_sampling": True,
                        "capture_type": capture, "account_first_and_last_bounces": capture == "single", "max_depth": 5},
         "sensor": {"type": "perspective", "fov": 40.0, "near_clip": 0.1, "far_clip": 100.0, "to_world": cam,
                    "sampler": {"type": "independent", "sample_count": 32},
                    "film": {"type": "transient_hdr_film", "width": 16, "height": 12, "temporal_bins": 256,
                             "start_opl": 0.0, "bin_width_opl": 12.0 / 256, "rfilter": {"type": "box"}}},
         "emitter": {"type": "projector", "irradiance": 100.0, "fov": 0.2, "to_world": cam},
         "relay_wall": {"type": "rectangle", "to_world": T().scale([2.0, 2.0, 1.0]),
                        "bsdf": {"type": "twosided", "bsdf": {"type": "diffuse", "reflectance": 0.8}}},
         "hidden": {"type": "trimesh", "vertices": v, "faces": f, "to_world": T().translate([0.0, 0.0, 1.0]),
                    "bsdf": {"type": "twosided", "bsdf": {"type": "diffuse", "reflectance": 1.0}}}}
    raw, ref = _both(mitr, oracle, d, 32, seed=4)
    assert ref["counters"]["shadow_rays"] > 0
