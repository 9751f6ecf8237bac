"""GPU parity of the two other film layouts against the CPU oracle: the `exhaustive_scan` transient film driven by
an Exhaustive NLOS capture (transient_image_block.py:59-66,134-140, transientnlospath.py:597-627) and
`phasor_hdr_film` (phasor_hdr_film.py, phasor_image_block.py).  Same bar as the other parity tests: identical
integer work counters, film relative L1 <= 1e-5 (only the order of the fp32 additions differs)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-5


def rel_l1(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).sum() / max(np.abs(b).sum(), 1e-300))


@pytest.fixture
def mono(mitr):
    mitr.set_variant("cuda_mono")
    yield
    mitr.set_variant("cuda_rgb")


def exhaustive_scene(mitr, res, bins, spp, **integ):
    d = mitr.nlos_z_scene(res, res, bins, spp, start_opl=0.0, window=8.0, capture_type="exhaustive", **integ)
    f = d["relay_wall"]["nlos_sensor"]["transient_film"]
    f["exhaustive_scan"], f["laser_scan_width"], f["laser_scan_height"] = True, res, res
    return d


@pytest.mark.parametrize("integ", [{}, {"nlos_hidden_geometry_sampling": False, "account_first_and_last_bounces": True},
                                   {"nlos_laser_sampling": False, "account_first_and_last_bounces": True}])
def test_exhaustive_capture_parity(mitr, oracle, integ):
    res, bins, spp = 4, 64, 16
    d = exhaustive_scene(mitr, res, bins, spp, **integ)
    if integ.get("nlos_laser_sampling") is False:
        d["emitter"]["fov"] = 60.0
    scene = mitr.load_dict(d)
    it = scene.integrator()
    it.collect_counters = True
    steady, transient = it.render(scene, spp=spp, seed=3)
    assert tuple(transient.shape) == (res, res, res, res, bins, 3)
    raw = scene.sensors()[0].film().develop(raw=True)[1].cpu().numpy()
    ref = oracle.render(scene, it, spp=spp, seed=3)
    assert raw.shape == ref["transient"].shape == (res, res, res, res, bins, 4)
    assert it.last_counters == ref["counters"], (it.last_counters, ref["counters"])
    assert ref["counters"]["deposits_in_range"] > 0
    assert rel_l1(raw, ref["transient"]) <= TOL
    assert ((raw != 0) == (ref["transient"] != 0)).all()
    # the reference's develop() of an exhaustive film returns the CHANNEL mean of the transient image as "steady"
    # (transient_hdr_film.py:213-214); the per-pixel radiance image stays available as film.steady_storage
    assert tuple(steady.shape) == (res, res, res, res, bins)
    np.testing.assert_allclose(steady.cpu().numpy(), transient.cpu().numpy().mean(axis=-1), rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(scene.sensors()[0].film().steady_storage.cpu().numpy(), ref["steady"], rtol=1e-4, atol=1e-6)
    if integ.get("nlos_laser_sampling") is False:       # plain NEE: every deposit lands at laser (0, 0)  (:500-505)
        assert np.abs(raw[:, :, 1:]).sum() == 0 and np.abs(raw[:, :, 0, 1:]).sum() == 0


def test_exhaustive_rejections(mitr):
    d = mitr.nlos_z_scene(4, 4, 16, 1, capture_type="exhaustive")          # plain film
    scene = mitr.load_dict(d)
    with pytest.raises(ValueError, match="exhaustive_scan"):
        scene.integrator().render(scene, spp=1)
    d = exhaustive_scene(mitr, 4, 16, 1)
    d["relay_wall"]["nlos_sensor"]["transient_film"]["laser_scan_width"] = 2
    scene = mitr.load_dict(d)
    with pytest.raises(AssertionError, match="must be equal"):
        scene.integrator().render(scene, spp=1)


def phasor_film(w, h, bins, window=(3.5, 9.5), wl_mean=0.5, wl_sigma=0.5):
    return {"type": "phasor_hdr_film", "width": w, "height": h, "temporal_bins": bins, "start_opl": window[0],
            "bin_width_opl": (window[1] - window[0]) / bins, "wl_mean": wl_mean, "wl_sigma": wl_sigma,
            "rfilter": {"type": "box"}}


def test_phasor_film_needs_mono(mitr):
    d = mitr.cornell_box_config(8, 8, 64, 2)
    d["sensor"]["film"] = phasor_film(8, 8, 64)
    scene = mitr.load_dict(d)
    with pytest.raises(ValueError, match="monochromatic"):
        scene.integrator().render(scene, spp=2)


@pytest.mark.parametrize("w,h,bins,spp,wl", [(24, 16, 256, 16, (0.5, 0.5)), (16, 16, 1024, 8, (0.2, 0.1))])
def test_phasor_film_cornell_parity(mitr, oracle, mono, w, h, bins, spp, wl):
    d = mitr.cornell_box_config(w, h, bins, spp)
    d["sensor"]["film"] = phasor_film(w, h, bins, wl_mean=wl[0], wl_sigma=wl[1])
    scene = mitr.load_dict(d)
    film = scene.sensors()[0].film()
    F = len(film.frequencies)
    assert F >= 8
    it = scene.integrator()
    it.collect_counters = True
    steady, phasors = it.render(scene, spp=spp, seed=1)
    assert tuple(phasors.shape) == (h, w, F, 2) and tuple(steady.shape) == (h, w, 1)
    raw = film.develop(raw=True)[1].cpu().numpy()
    ref = oracle.render(scene, it, spp=spp, seed=1)
    assert raw.shape == ref["transient"].shape == (h, w, 2 * F + 1)
    assert it.last_counters == ref["counters"], (it.last_counters, ref["counters"])
    assert rel_l1(raw, ref["transient"]) <= TOL
    assert np.abs(raw[..., -1]).sum() == 0                                   # weight channel stays 0 (:262)
    np.testing.assert_allclose(phasors.cpu().numpy().reshape(h, w, 2 * F), ref["transient"][..., :-1], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(steady.cpu().numpy()[..., 0], ref["steady"][..., 0], rtol=1e-4, atol=1e-6)


def test_phasor_film_nlos_parity(mitr, oracle, mono):
    d = mitr.nlos_z_scene(16, 16, 300, 64)
    d["relay_wall"]["nlos_sensor"]["transient_film"] = phasor_film(16, 16, 300, window=(1.85, 3.65), wl_mean=0.06, wl_sigma=0.06)
    scene = mitr.load_dict(d)
    film = scene.sensors()[0].film()
    it = scene.integrator()
    it.collect_counters = True
    steady, phasors = it.render(scene, spp=64, seed=0)
    raw = film.develop(raw=True)[1].cpu().numpy()
    ref = oracle.render(scene, it, spp=64, seed=0)
    assert it.last_counters == ref["counters"], (it.last_counters, ref["counters"])
    assert ref["counters"]["deposits_in_range"] > 0
    assert rel_l1(raw, ref["transient"]) <= TOL


def test_mono_transient_film_is_luminance(mitr, oracle, mono):
    d = mitr.cornell_box_config(16, 16, 128, 8)
    scene = mitr.load_dict(d)
    it = scene.integrator()
    steady, transient = it.render(scene, spp=8)
    assert tuple(transient.shape) == (16, 16, 128, 1) and tuple(steady.shape) == (16, 16, 1)
    raw = scene.sensors()[0].film().develop(raw=True)[1].cpu().numpy()
    assert (raw[..., 0] == raw[..., 1]).all() and (raw[..., 1] == raw[..., 2]).all()
    ref = oracle.render(scene, it, spp=8)
    assert rel_l1(raw, ref["transient"]) <= TOL


def test_exhaustive_own_laser_grid_that_misses_the_scene_is_refused(mitr, oracle):
    """transientnlospath.py:370-377: with laser sampling, an Exhaustive capture's own laser grid must hit the scene
    somewhere ("The emitter did not intersect any geometry in the scene").  The oracle and tr_render refuse alike -- on
    the first render AND on a second one that finds the cached laser targets (the check is cached with them)."""
    d = exhaustive_scene(mitr, 3, 32, 2)
    f = d["relay_wall"]["nlos_sensor"]["transient_film"]
    f["laser_scan_width"], f["laser_scan_height"] = 2, 2
    d["integrator"].update(force_equal_illumination_scanning=False, illumination_scan_fov=10.0)
    d["emitter"]["to_world"] = mitr.ScalarTransform4f().look_at(origin=[-0.5, 0.0, 0.25], target=[-10.0, 0.0, 0.25], up=[0, 1, 0])
    scene = mitr.load_dict(d)
    it = scene.integrator()
    with pytest.raises(RuntimeError, match="did not intersect any geometry"):
        oracle.render(scene, it, spp=2, seed=0)
    for _ in range(2):
        with pytest.raises(mitr.TransientError, match="did not intersect any geometry"):
            it.render(scene, spp=2, seed=0)
