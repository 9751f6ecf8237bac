"""Pins against the REFERENCE'S OWN CODE: tests/golden/reference_host_golden.* were produced by executing the
unmodified host-side Python of /root/reference/mitransient (film, image blocks, phasor film, scene dictionary,
integrator parameters, pass split, tonemapping) over the NumPy stand-in tests/golden/refshim.py
(tests/golden/make_reference_golden.py).  The oracle, the host mirror and -- on the GPU -- the film kernels must
reproduce them.  Deposits go in sequentially on both sides, so the film tensors match BIT FOR BIT on the CPU."""
import ctypes as C
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "reference_host_golden.npz"))
META = json.load(open(os.path.join(HERE, "golden", "reference_host_golden.json")))


def _film_desc(mitr, props, frequencies=None):
    fd = mitr._abi.TrFilmDesc()
    fd.width, fd.height, fd.temporal_bins = props["width"], props["height"], props["temporal_bins"]
    fd.start_opl, fd.bin_width_opl = props["start_opl"], props["bin_width_opl"]
    if props.get("exhaustive_scan"):
        fd.laser_scan_width, fd.laser_scan_height = props["laser_scan_width"], props["laser_scan_height"]
    fd.crop_offset_x, fd.crop_offset_y = props.get("crop_offset_x", 0), props.get("crop_offset_y", 0)   # block: film-sized + offset
    if frequencies is not None:
        fd.type, fd.n_frequencies = mitr._abi.TR_FILM_PHASOR_HDR, len(frequencies)
        fd.frequencies = frequencies.ctypes.data_as(C.POINTER(C.c_float))
    return fd


@pytest.mark.parametrize("tag", ["film_a", "film_b", "film_c", "film_d"])
def test_oracle_film_equals_reference_film_bit_for_bit(mitr, oracle, tag):
    """add_transient_data + TransientImageBlock.put_ + accum + develop_transient_ (transient_hdr_film.py:250-276,
    220-248; transient_image_block.py:83-151): window edges, inf / nan distances, zero deposits, W channel; film_d has
    a crop window (film-sized block addressed at pos - crop_offset, :131 and transient_hdr_film.py:192-195)."""
    props = META[tag]
    film = np.zeros(G[tag + "_raw"].shape, dtype=np.float32)
    oracle.film_scatter_add(film, _film_desc(mitr, props), G[tag + "_deposits"])
    assert np.array_equal(film, G[tag + "_raw"])
    assert np.all(film[..., 3] == 0)
    assert np.array_equal(oracle.film_develop(film), G[tag + "_developed"])


def test_oracle_exhaustive_index_equals_reference(mitr, oracle):
    """exhaustive_scan with a NON-square laser grid: the flat index of transient_image_block.py:134-140 exactly as
    written (laser_scan_width multiplies before laser_x, laser_scan_height before laser_y) inside the
    (H, W, Lh, Lw, T, C) tensor of :63-66."""
    props = META["film_exh"]
    film = np.zeros(G["film_exh_raw"].shape, dtype=np.float32)
    assert film.shape == (props["height"], props["width"], props["laser_scan_height"], props["laser_scan_width"],
                          props["temporal_bins"], 4)
    oracle.film_scatter_add_laser(film, _film_desc(mitr, props), G["film_exh_deposits"], G["film_exh_laser"])
    assert np.array_equal(film, G["film_exh_raw"])


@pytest.mark.parametrize("tag", ["phasor_a", "phasor_b", "phasor_c"])
def test_phasor_film_equals_reference(mitr, oracle, tag):
    """PhasorHDRFilm frequency selection (phasor_hdr_film.py:125-137) bit for bit; PhasorImageBlock.put
    (phasor_image_block.py:43-67) within the accuracy of the cos / sin implementations (NumPy there, Cephes
    polynomial here; the phase itself is computed with the reference's fp32 fmod formula)."""
    props = META["phasor"][tag]
    want_f = G[tag + "_frequencies"]
    film_plugin = mitr.load_dict({"type": "phasor_hdr_film", **{k: v for k, v in props.items() if k not in ("channels", "n_channels")}})
    assert np.array_equal(film_plugin.frequencies, want_f)
    assert film_plugin.raw_channels() == props["n_channels"] == 2 * len(want_f) + 1
    assert film_plugin.channels[:2] + film_plugin.channels[-1:] == props["channels"]
    film = np.zeros(G[tag + "_raw"].shape, dtype=np.float32)
    oracle.film_scatter_add(film, _film_desc(mitr, props, np.ascontiguousarray(want_f)), G[tag + "_deposits"])
    ref = G[tag + "_raw"]
    scale = np.abs(G[tag + "_deposits"][:, 3]).sum() / film[..., 0].size
    assert np.abs(film - ref).max() <= 2e-5 * max(scale, 1.0), np.abs(film - ref).max()
    assert np.all(film[..., -1] == 0)
    dev = oracle.film_develop(film).reshape(ref.shape[:-1] + (len(want_f), 2))
    assert np.abs(dev - G[tag + "_developed"]).max() <= 2e-5 * max(scale, 1.0)


def _same(a, b, path=""):
    if isinstance(b, dict) and "__transform__" in b:
        np.testing.assert_allclose(np.asarray(a.matrix, dtype=np.float64), np.array(b["__transform__"]), atol=1e-12, err_msg=path)
    elif isinstance(b, dict) and "__array__" in b:
        np.testing.assert_allclose(np.asarray(a, dtype=np.float64), np.array(b["__array__"]), atol=0, err_msg=path)
    elif isinstance(b, dict):
        assert isinstance(a, dict) and set(a) == set(b), (path, sorted(a), sorted(b))
        for k in b:
            _same(a[k], b[k], path + "/" + k)
    elif isinstance(b, list):
        assert len(a) == len(b), path
        for i, (x, y) in enumerate(zip(a, b)):
            _same(x, y, f"{path}[{i}]")
    else:
        assert a == b, (path, a, b)


def test_cornell_box_dictionary_equals_reference(mitr):
    """mitr.cornell_box() == mitransient.utils.cornell_box() (utils.py:78-220): every key, value, colour and
    transform matrix of the benchmark scene."""
    _same(mitr.cornell_box(), META["cornell_box"])
    assert mitr.speed_of_light == META["speed_of_light"]


def test_nlos_integrator_parameters_equal_reference(mitr):
    """TransientNLOSPath.__init__ (transientnlospath.py:201-249) + TransientADIntegrator.__init__ (common.py:22-30)."""
    for case in META["nlos_params"]:
        it = mitr.load_dict({"type": "transient_nlos_path", **case["props"]})
        for k, v in case["attrs"].items():
            got = getattr(it, k)
            assert (int(got) if not isinstance(v, float) else float(got)) == v, (case["props"], k, got, v)
    with pytest.raises(AssertionError) as e:
        mitr.load_dict({"type": "transient_nlos_path", "filter_depth": 2, "filter_bounces": 1})
    assert str(e.value) == META["nlos_both_filters_error"]
    it = mitr.load_dict({"type": "transient_path", "max_depth": 8, "camera_unwarp": True, "gaussian_stddev": 2.0,
                         "temporal_filter": "box"})
    for k, v in META["path_params"].items():
        assert getattr(it, k) == v


def test_pass_split_equals_reference(mitr, monkeypatch):
    """TransientADIntegrator.prepare (common.py:32-85): one pass up to 2^32 samples, else int((2^26-1)/(W*H)) spp
    per pass plus a remainder pass."""
    from mitransient_b200.films.transient_hdr_film import TransientHDRFilm
    monkeypatch.setattr(TransientHDRFilm, "prepare", lambda self, aovs=(): 4)        # no device allocation here
    for case in META["prepare_split"]:
        d = mitr.cornell_box_config(case["width"], case["height"], 2, case["film_spp"])
        scene = mitr.load_dict(d)
        passes = scene.integrator().prepare(scene, scene.sensors()[0], 0, case["spp"])
        assert [s for _, s in passes] == case["passes"], case


def test_tonemap_equals_reference(mitr):
    np.testing.assert_allclose(mitr.vis.tonemap_transient(G["tonemap_in"], scaling=2.0), G["tonemap_out"], rtol=1e-6)


# ------------------------------------------------------------------------------------------ GPU kernels
@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["film_a", "film_b", "film_c", "film_d"])
def test_gpu_film_kernels_equal_reference(mitr, tag):
    """tr_film_scatter_add + tr_film_develop through the host mirror's add_transient_data / develop against the
    reference film (only the order of the fp32 additions differs)."""
    import torch
    props = META[tag]
    film = mitr.load_dict({"type": "transient_hdr_film", **props})
    film.prepare()
    dep = torch.from_numpy(G[tag + "_deposits"]).cuda()
    film.add_transient_data(dep[:, 0:2], dep[:, 2], None, dep[:, 3:6])
    raw = film.develop(raw=True)[1].cpu().numpy()
    ref = G[tag + "_raw"]
    assert np.array_equal(raw != 0, ref != 0)
    assert np.abs(raw - ref).sum() <= 1e-5 * np.abs(ref).sum()       # film_c: 5000 deposits into ONE bin, atomics in any order
    dev = film.develop()[1].cpu().numpy()
    assert np.abs(dev - G[tag + "_developed"]).sum() <= 1e-5 * np.abs(ref).sum()


# ------------------------------------------------------------------------------------------ the integrator loop
import sys  # noqa: E402

sys.path.insert(0, os.path.join(HERE, "golden"))
from path_cases import PATH_CASES, path_case_scene, sort_rows  # noqa: E402

GP = np.load(os.path.join(HERE, "golden", "reference_path_golden.npz"))


@pytest.mark.parametrize("name", sorted(PATH_CASES))
def test_oracle_deposits_equal_the_reference_integrator(mitr, oracle, name):
    """The reference's render() (common.py:122-213) and TransientPath.sample() (transientpath.py:88-326) were EXECUTED
    over the oracle's scene services; every deposit (pixel, optical path length, RGB value) the oracle produces on
    its own must be one of theirs, bit for bit -- draw order, masks, MIS, roulette, depth / eta / OPL bookkeeping."""
    scene, cfg, _ = path_case_scene(name)
    r = oracle.render(scene, scene.integrator(), spp=cfg[3], seed=3, log_deposits=1 << 20)
    assert np.array_equal(sort_rows(r["deposits"]), GP[name + "_deposits"])
    # hdrfilm divides the sum by the weight (= spp); the oracle multiplies by 1/spp: one ulp apart when spp is not 2^k
    np.testing.assert_allclose(r["steady"], GP[name + "_steady"], rtol=3e-7, atol=0)
    ref = GP[name + "_raw"]                                   # same deposits, different summation order
    assert np.abs(r["transient"] - ref).max() <= 1e-6 * max(1.0, np.abs(ref).max())


@pytest.mark.skipif(not os.path.isdir("/root/reference/mitransient"), reason="the reference tree is not available here")
def test_reference_integrator_runs_live(mitr, oracle):
    """Re-execute the reference integrator here (build container only) on a configuration that is NOT in the golden
    file and compare with the oracle."""
    import warnings
    import refshim
    d = mitr.cornell_box_config(7, 5, 40, 6, max_depth=7)
    d["integrator"]["rr_depth"] = 3
    d["large-box"]["bsdf"] = {"type": "conductor"}
    scene = mitr.load_dict(d)
    props = {k: v for k, v in d["integrator"].items() if k != "type"}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        steady, raw, dep = refshim.reference_render(scene, "mitransient.integrators.transientpath", "TransientPath", props,
                                                    spp=6, seed=11)
    r = oracle.render(scene, scene.integrator(), spp=6, seed=11, log_deposits=1 << 20)
    assert len(dep) > 100 and np.array_equal(sort_rows(dep), sort_rows(r["deposits"]))
    np.testing.assert_allclose(steady.astype(np.float32), r["steady"], rtol=3e-7, atol=0)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(PATH_CASES))
def test_gpu_render_equals_the_reference_integrator(mitr, name):
    """The CUDA path against the film / steady image the reference integrator produced (same deposits, other order)."""
    scene, cfg, _ = path_case_scene(name)
    integ = scene.integrator()
    integ.collect_counters = True
    steady, _ = integ.render(scene, spp=cfg[3], seed=3)
    raw = scene.sensors()[0].film().develop(raw=True)[1].cpu().numpy()
    ref = GP[name + "_raw"]
    assert integ.last_counters["deposits"] == len(GP[name + "_deposits"])
    assert np.array_equal(raw != 0, ref != 0)
    assert np.abs(raw - ref).sum() <= 1e-5 * np.abs(ref).sum()
    np.testing.assert_allclose(steady.cpu().numpy(), GP[name + "_steady"], rtol=1e-4, atol=1e-6)


# ------------------------------------------------------------------------------------------ transient_nlos_path
from path_cases import NLOS_CASES, nlos_case_scene  # noqa: E402


@pytest.mark.parametrize("name", sorted(NLOS_CASES))
def test_oracle_nlos_deposits_equal_the_reference_integrator(mitr, oracle, name):
    """TransientNLOSPath.prepare / sample / emitter_laser_sample / emitter_laser_targets_sample / emitter_nee_sample /
    hidden_geometry_sample (transientnlospath.py:251-927) and NLOSCaptureMeter.sample_ray_differential
    (nloscapturemeter.py:136-202) were EXECUTED over the oracle's scene services: confocal, single and exhaustive
    captures, laser and plain emitter sampling, hidden-geometry sampling with and without roulette / relay wall,
    filter_depth, discard_direct_paths, account_first_and_last_bounces.  Deposits bit for bit; the exhaustive films
    (laser-indexed) element for element."""
    scene, cfg, _ = nlos_case_scene(name)
    r = oracle.render(scene, scene.integrator(), spp=cfg[3], seed=5, log_deposits=1 << 20)
    assert len(GP[name + "_deposits"]) > 0
    assert np.array_equal(sort_rows(r["deposits"]), GP[name + "_deposits"])
    ref = GP[name + "_raw"]
    assert r["transient"].shape == ref.shape and np.array_equal(r["transient"] != 0, ref != 0)
    assert np.abs(r["transient"] - ref).max() <= 1e-6 * max(1.0, np.abs(ref).max())


@pytest.mark.skipif(not os.path.isdir("/root/reference/mitransient"), reason="the reference tree is not available here")
def test_reference_nlos_integrator_runs_live(mitr, oracle):
    import warnings
    import refshim
    d = mitr.nlos_z_scene(4, 5, 48, 6, start_opl=0.0, window=8.0, max_depth=5)
    scene = mitr.load_dict(d)
    props = {k: v for k, v in d["integrator"].items() if k != "type"}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        steady, raw, dep = refshim.reference_render_nlos(scene, props, spp=6, seed=9)
    r = oracle.render(scene, scene.integrator(), spp=6, seed=9, log_deposits=1 << 20)
    assert len(dep) > 50 and np.array_equal(sort_rows(dep), sort_rows(r["deposits"]))


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(NLOS_CASES))
def test_gpu_nlos_render_equals_the_reference_integrator(mitr, name):
    scene, cfg, _ = nlos_case_scene(name)
    integ = scene.integrator()
    integ.collect_counters = True
    integ.render(scene, spp=cfg[3], seed=5)
    raw = scene.sensors()[0].film().develop(raw=True)[1].cpu().numpy()
    ref = GP[name + "_raw"]
    assert integ.last_counters["deposits"] == len(GP[name + "_deposits"])
    assert raw.shape == ref.shape and np.array_equal(raw != 0, ref != 0)
    assert np.abs(raw - ref).sum() <= 1e-5 * np.abs(ref).sum()


# ------------------------------------------------------------------------------------------ angulararea emitter
from path_cases import ANGULAR_CASES, angular_case_scene  # noqa: E402


@pytest.mark.parametrize("name", sorted(ANGULAR_CASES))
def test_oracle_angulararea_equals_the_reference_emitter(mitr, oracle, name):
    """AngularAreaLight.sample_direction / pdf_direction / eval / _fallof_curve (emitters/angulararea.py:73-141) were
    EXECUTED inside the reference's transient_path loop; oracle deposits bit for bit."""
    scene, cfg, _ = angular_case_scene(name)
    r = oracle.render(scene, scene.integrator(), spp=cfg[3], seed=2, log_deposits=1 << 20)
    assert len(GP[name + "_deposits"]) > 100
    assert np.array_equal(sort_rows(r["deposits"]), GP[name + "_deposits"])
    np.testing.assert_allclose(r["steady"], GP[name + "_steady"], rtol=3e-7, atol=0)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(ANGULAR_CASES))
def test_gpu_angulararea_equals_the_reference_emitter(mitr, name):
    scene, cfg, _ = angular_case_scene(name)
    integ = scene.integrator()
    integ.collect_counters = True
    steady, _ = integ.render(scene, spp=cfg[3], seed=2)
    raw = scene.sensors()[0].film().develop(raw=True)[1].cpu().numpy()
    ref = GP[name + "_raw"]
    assert integ.last_counters["deposits"] == len(GP[name + "_deposits"])
    assert np.array_equal(raw != 0, ref != 0)
    assert np.abs(raw - ref).sum() <= 1e-5 * np.abs(ref).sum()
    np.testing.assert_allclose(steady.cpu().numpy(), GP[name + "_steady"], rtol=1e-4, atol=1e-6)


# ------------------------------------------------------------------------------------------ transient_prbvolpath
from path_cases import VOL_CASES, VOL_CASES_GPU, vol_case_scene  # noqa: E402


@pytest.mark.parametrize("name", sorted(VOL_CASES))
def test_oracle_volpath_deposits_equal_the_reference_integrator(mitr, oracle, name):
    """TransientPRBVolpathIntegrator.sample / sample_emitter / first_surface (transient_prbvolpath.py:114-528) were
    EXECUTED over the oracle's scene services (homogeneous media, HG phase functions, null boundaries, the
    transmittance loop of next-event estimation, camera_unwarp, roulette at depth > rr_depth): deposits bit for bit."""
    scene, cfg, _ = vol_case_scene(name)
    r = oracle.render(scene, scene.integrator(), spp=cfg[3], seed=7, log_deposits=1 << 20)
    assert len(GP[name + "_deposits"]) > 100
    assert np.array_equal(sort_rows(r["deposits"]), GP[name + "_deposits"])
    np.testing.assert_allclose(r["steady"], GP[name + "_steady"], rtol=3e-7, atol=0)


@pytest.mark.skipif(not os.path.isdir("/root/reference/mitransient"), reason="the reference tree is not available here")
def test_reference_volpath_integrator_runs_live(mitr, oracle):
    import warnings
    import refshim
    d = mitr.cornell_box_volumetric(5, 4, 40, 6, max_depth=7)
    scene = mitr.load_dict(d)
    props = {k: v for k, v in d["integrator"].items() if k != "type"}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        steady, raw, dep = refshim.reference_render_vol(scene, props, spp=6, seed=13)
    r = oracle.render(scene, scene.integrator(), spp=6, seed=13, log_deposits=1 << 20)
    assert len(dep) > 100 and np.array_equal(sort_rows(dep), sort_rows(r["deposits"]))


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(VOL_CASES_GPU))
def test_gpu_volpath_render_equals_the_reference_integrator(mitr, name):
    scene, cfg, _ = vol_case_scene(name)
    integ = scene.integrator()
    integ.collect_counters = True
    steady, _ = integ.render(scene, spp=cfg[3], seed=7)
    raw = scene.sensors()[0].film().develop(raw=True)[1].cpu().numpy()
    ref = GP[name + "_raw"]
    assert integ.last_counters["deposits"] == len(GP[name + "_deposits"])
    assert np.array_equal(raw != 0, ref != 0)
    assert np.abs(raw - ref).sum() <= 1e-5 * np.abs(ref).sum()
    np.testing.assert_allclose(steady.cpu().numpy(), GP[name + "_steady"], rtol=1e-4, atol=1e-6)


# ------------------------------------------------------------------------------------------ phasor film through the loop
from path_cases import PHASOR_CASES, phasor_case_scene  # noqa: E402


@pytest.fixture
def mono(mitr):
    mitr.set_variant("llvm_ad_mono")
    yield
    mitr.set_variant("cuda_rgb")


@pytest.mark.parametrize("name", sorted(PHASOR_CASES))
def test_oracle_phasor_render_equals_the_reference(mitr, oracle, mono, name):
    """The reference's PhasorHDRFilm + PhasorImageBlock collecting the deposits of the reference's transient_path loop
    (monochromatic variant): the oracle's phasor film within the cos / sin accuracy (NumPy there, Cephes here)."""
    scene, cfg, _ = phasor_case_scene(name)
    r = oracle.render(scene, scene.integrator(), spp=cfg[3], seed=3)
    ref = GP[name + "_raw"]
    assert r["transient"].shape == ref.shape
    assert np.abs(r["transient"] - ref).max() <= 2e-6 * max(1.0, np.abs(ref).max())
    np.testing.assert_allclose(r["steady"][..., 0], GP[name + "_steady"][..., 0], rtol=3e-7, atol=0)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(PHASOR_CASES))
def test_gpu_phasor_render_equals_the_reference(mitr, mono, name):
    scene, cfg, _ = phasor_case_scene(name)
    integ = scene.integrator()
    integ.render(scene, spp=cfg[3], seed=3)
    raw = scene.sensors()[0].film().develop(raw=True)[1].cpu().numpy()
    ref = GP[name + "_raw"]
    assert np.abs(raw - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max())


# ------------------------------------------------------------------------------------------ random option sweeps (live)
@pytest.mark.skipif(not os.path.isdir("/root/reference/mitransient"), reason="the reference tree is not available here")
@pytest.mark.parametrize("which,seed", [("path", 1300), ("path", 1307), ("path", 1311), ("nlos", 105), ("nlos", 1203), ("nlos", 1210),
                                        ("nlos", 1226), ("vol", 1400), ("vol", 1403), ("vol", 1409), ("nlos_scene", 2001),
                                        ("nlos_scene", 2005), ("angular", 600), ("angular", 603), ("phasor", 500), ("phasor", 502)])
def test_oracle_equals_the_executed_reference_on_random_configurations(mitr, oracle, which, seed):
    """A slice of tests/golden/fuzz_reference.py (about 7,000 configurations were swept clean when it was written): random integrator
    options, BSDFs, media parameters and occluders; the oracle's deposits must equal the executed reference's."""
    import fuzz_reference
    assert fuzz_reference.run_one(which, seed)
