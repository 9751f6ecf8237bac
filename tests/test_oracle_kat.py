"""Known-answer tests that pin the CPU oracle.

The reference ships no golden vectors for this path (its only test is stale and shape-only) and Mitsuba
cannot be installed here, so parity is *unpinned* against the reference itself; these analytic tests are what
validate the restatement."""
import math

import numpy as np
import pytest


def test_kat8_pcg32_known_answers(oracle):
    """pcg32 reference demo: pcg32_srandom(42, 54) -> first six outputs (public PCG test vectors)."""
    want = [0xA15C02B7, 0x7B47F409, 0xBA1D3330, 0x83D2F293, 0xBFA4784B, 0xCBED606E]
    assert oracle.pcg32(42, 54, 6).tolist() == want


def test_kat8_tea_and_host_sampler_agree(oracle):
    from mitransient_b200 import sampler as hs
    for v0, v1 in [(0, 0), (1, 2), (0xFFFFFFFF, 12345), (7, 0xDEADBEEF)]:
        assert oracle.tea32(v0, v1) == hs.sample_tea_32(v0, v1)
    assert oracle.tea32(0, 0) != (0, 0)
    for seed, lane in [(0, 0), (0, 1), (5, 77), (123456, 2 ** 31 + 3)]:
        f = oracle.sampler_floats(seed, lane, 4)
        assert 0.0 <= f.min() and f.max() < 1.0
        assert np.float32(hs.independent_sampler_first_float(seed, lane)) == f[0]
    # float conversion: 23 mantissa bits of the top of the word
    r = hs.PCG32(42, 54)
    u = r.next_uint32()
    assert u == 0xA15C02B7


def test_sincos_and_cosine_hemisphere(oracle):
    xs = np.linspace(-3.0, 3.0, 601, dtype=np.float32)
    for x in xs:
        s, c = oracle.sincos(float(x))
        assert abs(s - math.sin(float(x))) < 3e-7 and abs(c - math.cos(float(x))) < 3e-7
    rng = np.random.default_rng(0)
    u = rng.random((20000, 2), dtype=np.float32)
    w = np.array([oracle.cosine_hemisphere(float(a), float(b)) for a, b in u])
    assert np.all(w[:, 2] >= 0) and np.allclose(np.linalg.norm(w, axis=1), 1.0, atol=1e-5)
    assert abs(w[:, 2].mean() - 2.0 / 3.0) < 0.01          # E[cos] = 2/3 under a cosine-weighted pdf
    assert abs(w[:, 0].mean()) < 0.01 and abs(w[:, 1].mean()) < 0.01
    np.testing.assert_allclose(oracle.cosine_hemisphere(0.5, 0.5), [0, 0, 1], atol=1e-7)


def test_camera_rays(oracle, mitr):
    scene = mitr.load_dict(mitr.cornell_box_config(64, 64, 16, 1))
    sensor = scene.sensors()[0]
    np.testing.assert_allclose(sensor.to_world.matrix[:3, :3], [[-1, 0, 0], [0, 1, 0], [0, 0, -1]], atol=1e-12)
    o, d, maxt = oracle.camera_ray(scene, sensor, 31, 31, 1.0 - 2 ** -24, 1.0 - 2 ** -24)   # image centre
    np.testing.assert_allclose(d, [0, 0, -1], atol=1e-6)
    np.testing.assert_allclose(o, [0, 0, 3.9 - 0.001], atol=1e-6)      # origin pushed to the near plane
    assert abs(maxt - (100.0 - 0.001)) < 1e-3
    # field of view: the left edge of the image looks towards +x ... no: towards -x (red wall is on the left)
    o, d, _ = oracle.camera_ray(scene, sensor, 0, 32, 0.0, 0.0)
    half = math.radians(39.3077 / 2)
    assert d[0] < 0 and abs(math.atan2(-d[0], -d[2]) - half) < 1e-5
    o, d, _ = oracle.camera_ray(scene, sensor, 32, 0, 0.0, 0.0)        # top edge looks up
    assert d[1] > 0 and abs(math.atan2(d[1], -d[2]) - half) < 1e-5
    hit = oracle.ray_intersect(scene, sensor, [0, 0.8, 3.9], [0, 0, -1])      # above the boxes: back wall
    assert hit is not None and abs(hit["t"] - 4.9) < 1e-5 and np.allclose(hit["n"], [0, 0, 1])
    hit = oracle.ray_intersect(scene, sensor, [0, 0, 3.9], [0, 0, -1])        # centre ray: the large box
    assert hit is not None and hit["shape"] == 7 and 3.9 < hit["t"] < 4.1
    assert oracle.ray_intersect(scene, sensor, [0, 0, 3.9], [0, 0, 1]) is None   # out of the open front


def _plane_scene(mitr, light_y=1.0, light_half=0.01, width=9, height=9, bins=600, bw=0.01, spp=64, max_depth=2,
                 radiance=100.0, albedo=0.5):
    T = mitr.ScalarTransform4f
    return {
        "type": "scene",
        "integrator": {"type": "transient_path", "max_depth": max_depth},
        "sensor": {"type": "perspective", "fov": 20.0, "near_clip": 0.01, "far_clip": 100.0,
                   "to_world": T().look_at(origin=[0, 2.0, 2.0], target=[0, 0, 0], up=[0, 1, 0]),
                   "sampler": {"type": "independent", "sample_count": spp},
                   "film": {"type": "transient_hdr_film", "width": width, "height": height, "rfilter": {"type": "box"},
                            "temporal_bins": bins, "start_opl": 0.0, "bin_width_opl": bw}},
        "floor": {"type": "rectangle", "to_world": T().rotate([1, 0, 0], -90).scale(5.0),
                  "bsdf": {"type": "diffuse", "reflectance": {"type": "rgb", "value": [albedo] * 3}}},
        "light": {"type": "rectangle",
                  "to_world": T().translate([0, light_y, 0]).rotate([1, 0, 0], 90).scale(light_half),
                  "bsdf": {"type": "diffuse", "reflectance": {"type": "rgb", "value": [0.0] * 3}},
                  "emitter": {"type": "area", "radiance": {"type": "rgb", "value": [radiance] * 3}}},
    }


def test_kat2_direct_only_closed_form(oracle, mitr):
    """max_depth=2, a small area light over a diffuse plane: each pixel's radiance is
    rho/pi * Le * cos * cos' * A / r^2 and it lands in bin floor((|cam-x| + |x-light|)/bw)."""
    d = _plane_scene(mitr, spp=256)
    scene = mitr.load_dict(d)
    r = oracle.render(scene, scene.integrator(), spp=256, seed=0)
    tr, st = r["transient"][..., :3], r["steady"]
    light = np.array([0, 1.0, 0])
    A = (2 * 0.01) ** 2
    checked = 0
    for (px, py) in [(4, 4), (2, 6), (7, 1), (0, 8)]:
        o, dvec, _ = oracle.camera_ray(scene, scene.sensors()[0], px, py, 0.5, 0.5)
        t = -o[1] / dvec[1]
        x = o + t * dvec                       # point on the floor seen through the pixel centre
        wl = light - x
        r2 = float(wl @ wl)
        cos_s = wl[1] / math.sqrt(r2)          # at the floor (normal +y)
        cos_l = cos_s                          # at the light (normal -y), same angle
        want = 0.5 / math.pi * 100.0 * cos_s * cos_l * A / r2
        got = st[py, px]
        assert np.allclose(got, want, rtol=0.03), (px, py, got, want)     # pixel footprint + small-light limit
        # time of flight: the OPL is measured from the ray origin (on the near plane) to x to the light;
        # over the pixel footprint it spans the values at the four pixel corners
        def opl_at(jx, jy):
            oo, dd, _ = oracle.camera_ray(scene, scene.sensors()[0], px, py, jx, jy)
            xx = oo + (-oo[1] / dd[1]) * dd
            return float(np.linalg.norm(xx - oo) + np.linalg.norm(light - xx))
        corner = [opl_at(a, b) for a in (0.0, 1.0) for b in (0.0, 1.0)]
        light_extent = 2 * 0.01 * math.sqrt(2)
        lo_bin = int(math.floor((min(corner) - light_extent) / 0.01))
        hi_bin = int(math.floor((max(corner) + light_extent) / 0.01))
        nz = np.nonzero(tr[py, px, :, 0])[0]
        assert nz.min() >= lo_bin and nz.max() <= hi_bin, (nz, lo_bin, hi_bin)
        centre_bin = int(math.floor(opl_at(0.5, 0.5) / 0.01))
        mean_bin = float((tr[py, px, :, 0] * np.arange(tr.shape[2])).sum() / tr[py, px, :, 0].sum())
        assert abs(mean_bin - centre_bin) <= 1.5, (mean_bin, centre_bin)
        np.testing.assert_allclose(tr[py, px].sum(axis=0), got, rtol=1e-5)    # KAT-1 for this pixel
        checked += 1
    assert checked == 4


def test_kat1_energy_transient_sums_to_steady(oracle, mitr):
    """With a window covering every path length, transient.sum(time) == steady (the reference states this
    invariant in examples/transient/0-render_cbox_diffuse.ipynb cell 9)."""
    d = mitr.cornell_box_config(24, 24, 512, 32, window=(0.0, 64.0))
    scene = mitr.load_dict(d)
    r = oracle.render(scene, scene.integrator(), spp=32, seed=1)
    assert r["counters"]["deposits"] == r["counters"]["deposits_in_range"]
    np.testing.assert_allclose(r["transient"][..., :3].sum(axis=2), r["steady"], rtol=2e-5, atol=1e-7)
    assert np.all(r["transient"][..., 3] == 0)            # weight channel never written (appendix B.2)


def test_kat3_binning_edges(oracle, mitr):
    """Deposits at opl = start, just below the end, just below start and at the end: in / in / out / out
    (transient_hdr_film.py:263-265)."""
    from mitransient_b200 import _abi
    fd = _abi.TrFilmDesc()
    fd.width, fd.height, fd.temporal_bins, fd.start_opl, fd.bin_width_opl = 4, 2, 8, 1.0, 0.25
    end = np.float32(1.0 + 8 * 0.25)
    dep = np.array([
        [0, 0, 1.0, 1, 2, 3],
        [1, 0, np.nextafter(end, np.float32(0)), 1, 2, 3],
        [2, 0, np.nextafter(np.float32(1.0), np.float32(0)), 1, 2, 3],
        [3, 0, end, 1, 2, 3],
        [3, 1, np.inf, 1, 2, 3],            # a ray that missed: distance = +inf (appendix B.4)
        [4, 0, 1.5, 1, 2, 3],               # x outside the film
        [-1, 1, 1.5, 1, 2, 3],
    ], dtype=np.float32)
    film = np.zeros((2, 4, 8, 4), dtype=np.float32)
    oracle.film_scatter_add(film, fd, dep)
    assert film[0, 0, 0, :3].tolist() == [1, 2, 3]
    assert film[0, 1, 7, :3].tolist() == [1, 2, 3]
    assert film.sum() == 12.0


def test_kat4_layout_non_square(oracle, mitr):
    """Non-square film: the deposit of pixel (x, y) lands at film[y][x][t][c] (transient_image_block.py:142-144)."""
    from mitransient_b200 import _abi
    fd = _abi.TrFilmDesc()
    fd.width, fd.height, fd.temporal_bins, fd.start_opl, fd.bin_width_opl = 4, 2, 3, 0.0, 1.0
    film = np.zeros((2, 4, 3, 4), dtype=np.float32)
    oracle.film_scatter_add(film, fd, np.array([[3, 1, 2.5, 7, 8, 9]], dtype=np.float32))
    flat = film.reshape(-1)
    idx = ((1 * 4 + 3) * 3 + 2) * 4
    assert flat[idx:idx + 4].tolist() == [7, 8, 9, 0] and flat.sum() == 24
    dev = oracle.film_develop(film)
    assert dev.shape == (2, 4, 3, 3) and dev[1, 3, 2].tolist() == [7, 8, 9]
    film[1, 3, 2, 3] = 2.0                                  # a non-zero weight divides (transient_hdr_film.py:243-246)
    assert oracle.film_develop(film)[1, 3, 2].tolist() == [3.5, 4, 4.5]


def test_kat5_furnace(oracle, mitr):
    """Closed cube, every wall a diffuse emitter of radiance Le and albedo rho, camera inside:
    L = Le * sum_{k<max_depth} rho^k exactly (each path sees the emitter at every vertex)."""
    T = mitr.ScalarTransform4f
    rho, Le, depth = 0.5, 2.0, 5
    walls = {
        "floor": T().translate([0, -1, 0]).rotate([1, 0, 0], -90), "ceil": T().translate([0, 1, 0]).rotate([1, 0, 0], 90),
        "back": T().translate([0, 0, -1]), "front": T().translate([0, 0, 1]).rotate([0, 1, 0], 180),
        "right": T().translate([1, 0, 0]).rotate([0, 1, 0], -90), "left": T().translate([-1, 0, 0]).rotate([0, 1, 0], 90),
    }
    d = {"type": "scene", "integrator": {"type": "transient_path", "max_depth": depth, "rr_depth": 100},
         "sensor": {"type": "perspective", "fov": 60.0, "near_clip": 0.01, "far_clip": 100.0,
                    "to_world": T().look_at(origin=[0.1, 0.2, 0.3], target=[0, 0, -1], up=[0, 1, 0]),
                    "sampler": {"type": "independent", "sample_count": 64},
                    "film": {"type": "transient_hdr_film", "width": 8, "height": 8, "rfilter": {"type": "box"},
                             "temporal_bins": 64, "start_opl": 0.0, "bin_width_opl": 1.0}},
         "m": {"type": "diffuse", "reflectance": {"type": "rgb", "value": [rho] * 3}}}
    for name, tw in walls.items():
        d[name] = {"type": "rectangle", "to_world": tw, "bsdf": {"type": "ref", "id": "m"},
                   "emitter": {"type": "area", "radiance": {"type": "rgb", "value": [Le] * 3}}}
    scene = mitr.load_dict(d)
    r = oracle.render(scene, scene.integrator(), spp=256, seed=0)
    want = Le * sum(rho ** k for k in range(depth))
    # MIS between emitter sampling and BSDF sampling must add up to exactly one estimator per vertex
    assert abs(r["steady"].mean() - want) / want < 0.01, (r["steady"].mean(), want)
    np.testing.assert_allclose(r["transient"][..., :3].sum(axis=2), r["steady"], rtol=2e-5)


def test_box_primitive_equals_six_quads(oracle, mitr):
    """TR_PRIM_BOX (one slab test) gives the hits of the same cube described as six parallelograms."""
    d6 = mitr.cornell_box_config(16, 16, 32, 4)
    d6["small-box"]["box_primitive"] = False
    d6["large-box"]["box_primitive"] = False
    s1, s6 = mitr.load_dict(mitr.cornell_box_config(16, 16, 32, 4)), mitr.load_dict(d6)
    rng = np.random.default_rng(3)
    n_hit = 0
    for _ in range(1500):
        o = rng.uniform(-0.95, 0.95, 3)
        dd = rng.normal(size=3)
        dd /= np.linalg.norm(dd)
        a = oracle.ray_intersect(s1, s1.sensors()[0], o, dd)
        b = oracle.ray_intersect(s6, s6.sensors()[0], o, dd)
        assert (a is None) == (b is None)
        if a is None:
            continue
        n_hit += 1
        if abs(a["t"] - b["t"]) > 1e-4:          # grazing an edge: the two descriptions may pick different faces
            continue
        assert a["shape"] == b["shape"]
        if a["shape"] >= 6 and np.abs(a["n"] - b["n"]).max() > 1e-4:
            assert abs(np.dot(a["n"], b["n"])) < 0.1          # adjacent faces at an edge
    assert n_hit > 1000
    ra = oracle.render(s1, s1.integrator(), spp=64, seed=0)["steady"]
    rb = oracle.render(s6, s6.integrator(), spp=64, seed=0)["steady"]
    assert abs(ra.mean() - rb.mean()) / rb.mean() < 2e-3


def test_seed_changes_result_and_is_deterministic(oracle, mitr):
    scene = mitr.load_dict(mitr.cornell_box_config(16, 16, 32, 4))
    a = oracle.render(scene, scene.integrator(), spp=4, seed=0, n_threads=1)
    b = oracle.render(scene, scene.integrator(), spp=4, seed=0, n_threads=4)
    c = oracle.render(scene, scene.integrator(), spp=4, seed=1)
    assert np.array_equal(a["transient"], b["transient"]) and a["counters"] == b["counters"]
    assert not np.array_equal(a["transient"], c["transient"])


def test_oracle_matches_committed_golden(oracle, mitr):
    """Regression pin of the oracle itself: tests/golden/cornell_c1_oracle.npz (made by make_golden.py)."""
    import os
    path = os.path.join(os.path.dirname(__file__), "golden", "cornell_c1_oracle.npz")
    g = np.load(path)
    scene = mitr.load_dict(mitr.cornell_box_config(64, 64, 256, 16))
    r = oracle.render(scene, scene.integrator(), spp=16, seed=0)
    for k in ("paths", "vertices", "deposits", "deposits_in_range", "bin_checksum", "shadow_rays"):
        assert r["counters"][k] == int(g["counter_" + k]), k
    np.testing.assert_allclose(r["steady"], g["steady"], rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(r["transient"][..., :3].sum(axis=(0, 1)), g["temporal_profile"], rtol=1e-5, atol=1e-7)


# ---------------------------------------------------------------- phasor film / exhaustive scan
@pytest.fixture
def mono(mitr):
    mitr.set_variant("llvm_ad_mono")
    yield
    mitr.set_variant("cuda_rgb")


def test_phasor_film_is_the_fourier_transform_of_the_histogram(oracle, mitr, mono):
    """KAT: with a window covering every path length, the phasor at frequency f equals the DFT of the time
    histogram up to the bin-centre phase error pi*f*bin_width (phasor_image_block.py:52-58 vs
    transient_hdr_film.py:263)."""
    T, window, spp, n = 4096, (0.0, 60.0), 32, 12
    d = mitr.cornell_box_config(n, n, T, spp, window=window)
    s0 = mitr.load_dict(d)
    hist = oracle.render(s0, s0.integrator(), spp=spp)
    assert hist["counters"]["deposits"] == hist["counters"]["deposits_in_range"]      # nothing outside the window
    d = mitr.cornell_box_config(n, n, T, spp, window=window)
    d["sensor"]["film"] = {"type": "phasor_hdr_film", "width": n, "height": n, "temporal_bins": T, "start_opl": window[0],
                           "bin_width_opl": (window[1] - window[0]) / T, "wl_mean": 4.0, "wl_sigma": 2.0}
    s1 = mitr.load_dict(d)
    film = s1.sensors()[0].film()
    F = len(film.frequencies)
    ph = oracle.render(s1, s1.integrator(), spp=spp)
    assert ph["transient"].shape == (n, n, 2 * F + 1) and ph["counters"]["vertices"] == hist["counters"]["vertices"]
    assert ph["counters"]["deposits_in_range"] == hist["counters"]["deposits"]
    np.testing.assert_array_equal(ph["steady"], hist["steady"])
    t = (np.arange(T) + 0.5) * (window[1] - window[0]) / T
    h = hist["transient"][..., 0].astype(np.float64)
    got = ph["transient"][..., :2 * F].reshape(n, n, F, 2).astype(np.float64)
    for j in range(F):
        f = float(film.frequencies[j])
        ref = (h * np.exp(-2j * np.pi * f * t)).sum(-1)
        err = np.abs(got[..., j, 0] + 1j * got[..., j, 1] - ref).sum() / np.abs(ref).sum()
        assert err < 3.0 * np.pi * f * (window[1] - window[0]) / T + 1e-3, (j, f, err)
    assert np.abs(ph["transient"][..., -1]).sum() == 0


def test_exhaustive_capture_energy_and_layout(oracle, mitr):
    """KAT: an Exhaustive capture deposits Lr of EVERY illuminated point while the path's radiance is their mean
    (transientnlospath.py:621-627): sum(film) == n_laser * sum(steady) when the window holds everything; the
    confocal-like diagonal is populated and the film has the reference's (H, W, Lh, Lw, T, C) shape."""
    res, T, spp = 4, 64, 32
    d = mitr.nlos_z_scene(res, res, T, spp, start_opl=0.0, window=10.0, capture_type="exhaustive")
    f = d["relay_wall"]["nlos_sensor"]["transient_film"]
    f["exhaustive_scan"], f["laser_scan_width"], f["laser_scan_height"] = True, res, res
    scene = mitr.load_dict(d)
    r = oracle.render(scene, scene.integrator(), spp=spp)
    film, steady = r["transient"].astype(np.float64), r["steady"].astype(np.float64)
    assert film.shape == (res, res, res, res, T, 4)
    assert r["counters"]["deposits"] == r["counters"]["deposits_in_range"] > 0
    per_pixel = film[..., :3].sum(axis=(2, 3, 4))
    np.testing.assert_allclose(per_pixel, steady * res * res, rtol=2e-4, atol=1e-7)
    assert (film.sum(axis=(0, 1, 4, 5)) > 0).all()          # every illuminated point receives light from somewhere
    # the same scene under a confocal capture = only the laser point that coincides with the scanned point
    d2 = mitr.nlos_z_scene(res, res, T, spp, start_opl=0.0, window=10.0, capture_type="confocal")
    s2 = mitr.load_dict(d2)
    r2 = oracle.render(s2, s2.integrator(), spp=spp)
    assert r2["transient"].shape == (res, res, T, 4) and r2["counters"]["shadow_rays"] < r["counters"]["shadow_rays"]


# ---------------------------------------------------------------- conductor / dielectric (delta lobes, eta-scaled OPL)
def test_fresnel_known_answers(oracle):
    for c in (1.0, 0.7, 0.3, 0.05):
        assert abs(oracle.fresnel_conductor(c, 0.0, 1.0) - 1.0) < 1e-6               # material "none": perfect mirror
    n, k = 0.2, 3.5
    assert abs(oracle.fresnel_conductor(1.0, n, k) - ((n - 1) ** 2 + k * k) / ((n + 1) ** 2 + k * k)) < 1e-6
    r, cos_t, eta_it, eta_ti = oracle.fresnel_dielectric(1.0, 1.5)
    assert abs(r - 0.04) < 1e-6 and abs(cos_t + 1.0) < 1e-6 and abs(eta_it - 1.5) < 1e-7 and abs(eta_ti - 1 / 1.5) < 1e-7
    r, cos_t, eta_it, eta_ti = oracle.fresnel_dielectric(-1.0, 1.5)                  # from the inside
    assert abs(r - 0.04) < 1e-6 and abs(cos_t - 1.0) < 1e-6 and abs(eta_it - 1 / 1.5) < 1e-7
    r, cos_t, _, _ = oracle.fresnel_dielectric(-0.5, 1.5)                            # beyond the critical angle (41.8 deg)
    assert r == 1.0 and cos_t == 0.0
    c = 0.6                                                                          # Snell: sin_t = sin_i / eta
    r, cos_t, _, _ = oracle.fresnel_dielectric(c, 1.5)
    assert abs(-cos_t - math.sqrt(1 - (1 - c * c) / 2.25)) < 1e-6 and 0.04 < r < 1.0
    assert oracle.fresnel_dielectric(0.3, 1.0)[0] == 0.0                             # index matched


def _delta_scene(mitr, middle, T=512, window=(0.0, 16.0), spp=64, max_depth=8):
    """Pinhole at the origin looking down +z through a tiny field of view; `middle` = extra shapes; an area light
    (radiance 2) either behind the camera (mirror test) or far in front (glass test)."""
    TT = mitr.ScalarTransform4f
    d = {"type": "scene",
         "integrator": {"type": "transient_path", "max_depth": max_depth, "rr_depth": 1000},
         "sensor": {"type": "perspective", "fov": 0.5, "near_clip": 0.01, "far_clip": 100.0,
                    "to_world": TT().look_at(origin=[0, 0, 0], target=[0, 0, 1], up=[0, 1, 0]),
                    "sampler": {"type": "independent", "sample_count": spp},
                    "film": {"type": "transient_hdr_film", "width": 2, "height": 2, "temporal_bins": T,
                             "start_opl": window[0], "bin_width_opl": (window[1] - window[0]) / T,
                             "rfilter": {"type": "box"}}}}
    d.update(middle)
    return d


def test_mirror_time_of_flight(oracle, mitr):
    """KAT: a perfect mirror at z = 2 shows the light at z = -1: every pixel holds exactly the radiance, in the bin
    of OPL 2 + 3 (delta lobe: MIS weight 1, no emitter sampling at the mirror)."""
    TT = mitr.ScalarTransform4f
    d = _delta_scene(mitr, {
        "mirror": {"type": "rectangle", "to_world": TT().translate([0, 0, 2]).rotate([0, 1, 0], 180),
                   "bsdf": {"type": "conductor", "material": "none"}},
        "light": {"type": "rectangle", "to_world": TT().translate([0, 0, -1]),
                  "bsdf": {"type": "diffuse", "reflectance": 0.0}, "emitter": {"type": "area", "radiance": 2.0}}})
    scene = mitr.load_dict(d)
    r = oracle.render(scene, scene.integrator(), spp=64)
    T = 512
    film = r["transient"][..., :3]
    b = int(math.floor(5.0 / (16.0 / T)))
    got = film[:, :, b - 1:b + 2].sum(axis=2)
    np.testing.assert_allclose(got, 2.0, rtol=1e-4)
    assert np.abs(film).sum() == pytest.approx(got.sum(), rel=1e-6)
    assert r["counters"]["shadow_rays"] == 0                                         # nothing smooth was ever hit


def test_glass_slab_optical_path_length(oracle, mitr):
    """KAT: light at z = 4 seen through a 1-thick glass slab (eta 1.5) at normal incidence: the direct image arrives
    at OPL 1 + 1.5 + 2 = 4.5 with weight T^2 (T = 0.96, the eta^2 radiance factors cancel), the first internal
    double reflection at 4.5 + 3 with weight T^2 R^2 (transientpath.py:154: distance += si.t * eta)."""
    TT = mitr.ScalarTransform4f
    d = _delta_scene(mitr, {
        "slab": {"type": "cube", "to_world": TT().translate([0, 0, 1.5]).scale([5, 5, 0.5]),
                 "bsdf": {"type": "dielectric", "int_ior": 1.5, "ext_ior": 1.0}},
        "light": {"type": "rectangle", "to_world": TT().translate([0, 0, 4]).rotate([0, 1, 0], 180),
                  "bsdf": {"type": "diffuse", "reflectance": 0.0}, "emitter": {"type": "area", "radiance": 2.0}}},
        spp=4096, max_depth=12)
    scene = mitr.load_dict(d)
    r = oracle.render(scene, scene.integrator(), spp=4096)
    T, bw = 512, 16.0 / 512
    film = r["transient"][..., 0].astype(np.float64).mean(axis=(0, 1))
    b0, b1 = int(4.5 / bw), int(7.5 / bw)
    direct = film[b0 - 1:b0 + 2].sum()
    bounce = film[b1 - 1:b1 + 2].sum()
    assert direct == pytest.approx(2.0 * 0.96 ** 2, rel=0.02)
    assert bounce == pytest.approx(2.0 * 0.96 ** 2 * 0.04 ** 2, rel=0.25)
    assert film[:b0 - 1].sum() == 0.0                                                # nothing arrives earlier


def test_crop_window_camera_rays_are_the_full_films_rays(oracle, mitr):
    """Film crop windows (ADIntegrator.sample_rays [upstream]: pos = pixel + crop_offset, sample = fma(pos + jitter,
    1 / crop_size, -crop_offset / crop_size); perspective sensor: sample_to_camera built for the window): pixel (x, y) of
    the window looks along the same ray as pixel (x + ox, y + oy) of the whole film."""
    full = mitr.load_dict(mitr.cornell_box_config(12, 10, 32, 8))
    d = mitr.cornell_box_config(12, 10, 32, 8)
    d["sensor"]["film"].update(crop_width=5, crop_height=4, crop_offset_x=6, crop_offset_y=3)
    win = mitr.load_dict(d)
    for (x, y, jx, jy) in [(0, 0, 0.1, 0.9), (4, 3, 0.5, 0.5), (2, 1, 0.99, 0.01), (0, 3, 0.0, 0.0)]:
        o1, d1, m1 = oracle.camera_ray(win, win.sensors()[0], x, y, jx, jy)
        o2, d2, m2 = oracle.camera_ray(full, full.sensors()[0], x + 6, y + 3, jx, jy)
        assert np.abs(d1 - d2).max() < 2e-7 and np.abs(o1 - o2).max() < 2e-7 and abs(m1 - m2) < 1e-3
    # and the rendered window has the window's shape while the reference-shaped tensor keeps the film's
    r = oracle.render(win, win.integrator(), spp=4, seed=1)
    assert r["steady"].shape == (4, 5, 3) and r["transient"].shape == (10, 12, 32, 4)
    assert np.all(r["transient"][4:] == 0) and np.all(r["transient"][:, 5:] == 0) and r["transient"].sum() > 0
