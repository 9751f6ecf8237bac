"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Tolerances (fp32): the kernel and the oracle follow the same arithmetic contract and the same PCG32
streams, so every path is walked identically -- the integer work counters and the bin-index checksum must
be EQUAL -- and the only difference is the order of fp32 additions into a bin: whole-film relative L1
<= 1e-5, per-time-bin relative L1 <= 1e-4.
"""
import numpy as np
import pytest

from parity_utils import film_l1

pytestmark = pytest.mark.gpu

WHOLE_TOL = 1e-5
PER_BIN_TOL = 1e-4


def _render_both(mitr, oracle, cfg, spp, seed=0, **integ):
    d = mitr.cornell_box_config(*cfg, spp)
    d["integrator"].update(integ)
    scene = mitr.load_dict(d)
    integrator = scene.integrator()
    integrator.collect_counters = True
    steady, transient = integrator.render(scene, seed=seed, spp=spp)
    raw = scene.sensors()[0].film().develop(raw=True)[1]
    ref = oracle.render(scene, integrator, seed=seed, spp=spp)
    return (steady.cpu().numpy(), transient.cpu().numpy(), raw.cpu().numpy(), integrator.last_counters, ref)


def test_c1_cornell_parity(mitr, oracle):
    """BASELINE config C1: 64x64 px x 256 bins, 16 spp, max_depth 8."""
    steady, transient, raw, counters, ref = _render_both(mitr, oracle, (64, 64, 256), 16)
    assert counters == ref["counters"], f"work counters differ: gpu {counters} oracle {ref['counters']}"
    assert raw.shape == (64, 64, 256, 4) and transient.shape == (64, 64, 256, 3) and steady.shape == (64, 64, 3)
    assert np.all(raw[..., 3] == 0.0)                      # weight channel stays 0
    err = film_l1(raw, ref["transient"])
    assert err["zero_bins_mismatch"] == 0
    assert err["whole"] <= WHOLE_TOL and err["max_t"] <= PER_BIN_TOL, err
    np.testing.assert_allclose(transient, ref["transient"][..., :3], rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(steady, ref["steady"], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("w,h,t,spp,seed,integ", [
    (32, 16, 64, 8, 3, {}),                                # non-square film: layout film[y][x][t][c]
    (16, 16, 1000, 5, 1, {"max_depth": 3}),                # spp not a power of two, T not a power of two
    (24, 24, 128, 32, 7, {"camera_unwarp": True}),
    (24, 24, 128, 16, 2, {"discard_direct_light": True}),
    (16, 16, 64, 16, 0, {"max_depth": -1, "rr_depth": 2}),  # unbounded depth, early roulette
    (8, 8, 32, 64, 5, {"max_depth": 1}),                   # only directly visible emitters
])
def test_variants_parity(mitr, oracle, w, h, t, spp, seed, integ):
    steady, transient, raw, counters, ref = _render_both(mitr, oracle, (w, h, t), spp, seed=seed, **integ)
    assert counters == ref["counters"]
    err = film_l1(raw, ref["transient"])
    assert err["zero_bins_mismatch"] == 0
    assert err["whole"] <= WHOLE_TOL and err["max_t"] <= PER_BIN_TOL, err
    np.testing.assert_allclose(steady, ref["steady"], rtol=1e-4, atol=1e-6)


def _write_obj(path, verts, faces):
    with open(path, "w") as f:
        f.write("# test mesh\no test\n")
        for v in verts:
            f.write(f"v {v[0]:.6f} {v[1]:.6f} {v[2]:.6f}\n")
        for t in faces:
            f.write("f " + " ".join(str(i + 1) for i in t) + "\n")


def _uv_sphere(n_lat, n_lon, radius, center):
    verts, faces = [], []
    for i in range(n_lat + 1):
        th = np.pi * i / n_lat
        for j in range(n_lon):
            ph = 2 * np.pi * j / n_lon
            verts.append([center[0] + radius * np.sin(th) * np.cos(ph), center[1] + radius * np.cos(th),
                          center[2] + radius * np.sin(th) * np.sin(ph)])
    for i in range(n_lat):
        for j in range(n_lon):
            a, b = i * n_lon + j, i * n_lon + (j + 1) % n_lon
            c, d = a + n_lon, b + n_lon
            if i > 0:
                faces.append([a, b, c])
            if i < n_lat - 1:
                faces.append([b, d, c])
    return np.array(verts), faces


def test_bvh_mesh_scene_parity(mitr, oracle, tmp_path):
    """More primitives than the lock-step limit (40): the BVH traversal must give the oracle's brute-force hits; the light is an
    OBJ mesh (two triangles), which exercises Mesh::sample_position (area CDF + triangle warp)."""
    sv, sf = _uv_sphere(10, 16, 0.35, (0.3, -0.65, 0.2))
    _write_obj(tmp_path / "sphere.obj", sv, sf)
    lv = np.array([[-0.25, 0.98, -0.2], [0.25, 0.98, -0.2], [0.25, 0.98, 0.2], [-0.25, 0.98, 0.2]])
    _write_obj(tmp_path / "light.obj", lv, [[0, 1, 2], [0, 2, 3]])   # normal (0,-1,0): faces down
    d = mitr.cornell_box_config(40, 40, 200, 16)
    del d["light"], d["small-box"]
    d["sphere"] = {"type": "obj", "filename": str(tmp_path / "sphere.obj"), "bsdf": {"type": "ref", "id": "green"}}
    d["meshlight"] = {"type": "obj", "filename": str(tmp_path / "light.obj"), "bsdf": {"type": "ref", "id": "white"},
                      "merge_quads": False,     # keep two triangles: area CDF + uniform-triangle warp
                      "emitter": {"type": "area", "radiance": {"type": "rgb", "value": [10.0, 9.0, 7.0]}}}
    scene = mitr.load_dict(d)
    integ = scene.integrator()
    integ.collect_counters = True
    steady, transient = integ.render(scene, seed=4, spp=16)
    info = integ.device_scene(scene, scene.sensors()[0]).info()
    assert info["bvh_nodes"] > 0 and info["triangles"] > 64
    raw = scene.sensors()[0].film().develop(raw=True)[1].cpu().numpy()
    ref = oracle.render(scene, integ, seed=4, spp=16)
    assert integ.last_counters == ref["counters"]
    assert ref["steady"].mean() > 1e-3                      # the mesh light actually illuminates the scene
    err = film_l1(raw, ref["transient"])
    assert err["whole"] <= WHOLE_TOL and err["max_t"] <= PER_BIN_TOL, err
    np.testing.assert_allclose(steady.cpu().numpy(), ref["steady"], rtol=1e-4, atol=1e-6)


def test_plain_kernel_equals_counting_kernel(mitr):
    """The counter-free build (768 threads/CTA) and the counting build (512 threads/CTA) walk the same paths."""
    scene = mitr.load_dict(mitr.cornell_box_config(48, 32, 128, 16))
    integ = scene.integrator()
    integ.collect_counters = False
    s0, t0 = integ.render(scene, seed=2, spp=16)
    integ.collect_counters = True
    s1, t1 = integ.render(scene, seed=2, spp=16)
    a, b = t0.cpu().numpy().astype(np.float64), t1.cpu().numpy().astype(np.float64)
    assert np.abs(a - b).sum() / np.abs(b).sum() <= 1e-6
    assert np.array_equal(a != 0, b != 0)


def test_render_to_host_matches_device_render(mitr):
    """tr_render_to_host (banded, D2H overlapped) == tr_render + develop."""
    scene = mitr.load_dict(mitr.cornell_box_config(64, 40, 96, 8))
    integ = scene.integrator()
    steady, transient = integ.render(scene, seed=1, spp=8)
    hs, ht = integ.render_to_host(scene, seed=1, spp=8)
    a, b = ht.numpy().astype(np.float64), transient.cpu().numpy().astype(np.float64)
    assert ht.shape == (40, 64, 96, 3) and np.abs(a - b).sum() / np.abs(b).sum() <= 1e-6
    np.testing.assert_allclose(hs.numpy(), steady.cpu().numpy(), rtol=1e-5, atol=1e-7)
    hs4, ht4 = integ.render_to_host(scene, seed=1, spp=8, raw=True)
    assert ht4.shape == (40, 64, 96, 4) and np.all(ht4.numpy()[..., 3] == 0)


def test_film_scatter_add_and_develop_kernels(mitr, oracle):
    """K4 / K5 on their own against the oracle: a recorded deposit stream, including duplicates within a warp,
    out-of-window and out-of-film deposits."""
    import torch
    scene = mitr.load_dict(mitr.cornell_box_config(16, 12, 64, 4))
    ref = oracle.render(scene, scene.integrator(), seed=0, spp=4, log_deposits=200000)
    dep = ref["deposits"]
    assert len(dep) == ref["counters"]["deposits"]
    extra = np.array([[3, 2, 5.0, 1, 1, 1]] * 40 + [[99, 1, 5.0, 1, 1, 1], [3, -1, 5.0, 1, 1, 1], [3, 2, 1e9, 1, 1, 1],
                                                     [3, 2, np.inf, 1, 1, 1]], dtype=np.float32)
    dep = np.concatenate([dep, extra])
    film = scene.sensors()[0].film()
    film.prepare()
    d = torch.from_numpy(dep).cuda()
    film.add_transient_data(d[:, 0:2], d[:, 2], None, d[:, 3:6])
    got = film.develop(raw=True)[1].cpu().numpy()
    want = np.zeros_like(got)
    oracle.film_scatter_add(want, film.describe(), dep)
    assert np.array_equal(got != 0, want != 0)
    np.testing.assert_allclose(got, want, rtol=2e-5, atol=1e-7)
    dev = film.develop()[1].cpu().numpy()
    np.testing.assert_array_equal(dev, oracle.film_develop(got))


def test_box_as_emitter_and_quads_variant(mitr, oracle):
    """A `cube` that is an area light exercises position sampling on a box primitive (face by area, uniform point);
    the same scene with box_primitive=False goes through six parallelograms."""
    for as_box in (True, False):
        d = mitr.cornell_box_config(24, 24, 96, 32, max_depth=5)
        d["small-box"]["emitter"] = {"type": "area", "radiance": {"type": "rgb", "value": [2.0, 1.5, 1.0]}}
        d["small-box"]["box_primitive"] = as_box
        d["large-box"]["box_primitive"] = as_box
        scene = mitr.load_dict(d)
        assert len(scene.emitters()) == 2
        integ = scene.integrator()
        integ.collect_counters = True
        steady, _ = integ.render(scene, seed=9, spp=32)
        raw = scene.sensors()[0].film().develop(raw=True)[1].cpu().numpy()
        ref = oracle.render(scene, integ, seed=9, spp=32)
        assert integ.last_counters == ref["counters"]
        err = film_l1(raw, ref["transient"])
        assert err["zero_bins_mismatch"] == 0 and err["whole"] <= WHOLE_TOL and err["max_t"] <= PER_BIN_TOL, err
        np.testing.assert_allclose(steady.cpu().numpy(), ref["steady"], rtol=1e-4, atol=1e-6)


def test_sharded_render_to_host_assembles_to_the_full_film(mitr):
    """tr_render_to_host with world > 1: every rank gets the compact film of the pixels it owns (tr_owned_pixels
    order); scattering the shards back gives the film of the single-rank render."""
    import torch
    from mitransient_b200.distributed import owned_pixels
    w, h, t, spp, world = 72, 40, 96, 8, 3
    scene = mitr.load_dict(mitr.cornell_box_config(w, h, t, spp))
    integ = scene.integrator()
    steady, transient = integ.render(scene, spp=spp, seed=5)
    full = torch.zeros((h * w, t, 3)); full_s = torch.zeros((h * w, 3))
    for rank in range(world):
        s_r, t_r = integ.render_to_host(scene, spp=spp, seed=5, rank=rank, world=world, tile_size=16)
        px = torch.from_numpy(owned_pixels(w, h, 16, rank, world).astype(np.int64))
        assert t_r.shape == (len(px), t, 3)
        full[px] = t_r; full_s[px] = s_r
    ref = transient.cpu().reshape(h * w, t, 3)
    err = (full - ref).abs().sum() / ref.abs().sum()
    assert err <= 1e-5, err
    np.testing.assert_allclose(full_s.numpy(), steady.cpu().reshape(h * w, 3).numpy(), rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("w,h,t,spp", [(1, 1, 1, 1), (1, 1, 64, 33), (5, 3, 1, 7), (3, 2, 20000, 8), (33, 17, 7, 1)])
def test_edge_sizes(mitr, oracle, w, h, t, spp):
    """Degenerate and extreme film sizes: one pixel, one bin, one sample, ragged sizes, and a histogram too large
    for shared memory (20000 bins -> global vector-reduction path, cleared by the library)."""
    steady, transient, raw, counters, ref = _render_both(mitr, oracle, (w, h, t), spp, seed=9)
    assert counters == ref["counters"]
    assert raw.shape == (h, w, t, 4)
    err = film_l1(raw, ref["transient"])
    assert err["zero_bins_mismatch"] == 0 and err["whole"] <= WHOLE_TOL, err
    np.testing.assert_allclose(steady, ref["steady"], rtol=1e-4, atol=1e-6)


def test_too_many_samples_is_rejected_like_the_reference_splits(mitr):
    """width*height*spp > 2^32 never reaches the kernel in one pass: prepare() splits it (common.py:51-85) and the
    C ABI rejects a single oversized pass."""
    import ctypes as C
    from mitransient_b200 import _lib
    scene = mitr.load_dict(mitr.cornell_box_config(512, 512, 8, 1))
    integ = scene.integrator()
    passes = integ.prepare(scene, scene.sensors()[0], 0, 2 ** 15)                    # 2^33 samples
    assert len(passes) > 1 and sum(s for _, s in passes) == 2 ** 15 and all(s * 512 * 512 < 2 ** 26 + 1 for _, s in passes)
    dev = integ.device_scene(scene, scene.sensors()[0])
    film = scene.sensors()[0].film()
    p = integ.make_params(2 ** 15, 2 ** 15, 0)
    rc = _lib.load().tr_render(dev.handle, C.byref(p), film.transient_storage.data_ptr(), film.steady_storage.data_ptr(),
                               None, _lib.current_stream_ptr())
    assert rc == mitr._abi.TR_ERR_INVALID_ARGUMENT and b"2^32" in _lib.load().tr_last_error()


def test_multi_pass_accumulation(mitr, oracle):
    """Two passes into the same film (what prepare() produces above 2^32 samples, common.py:56-85): pass 2 adds to
    pass 1 (`accumulate`), both scaled by 1 / total_spp."""
    import ctypes as C
    from mitransient_b200 import _lib
    from oracle import oracle as o
    scene = mitr.load_dict(mitr.cornell_box_config(24, 16, 96, 8))
    integ, sensor = scene.integrator(), scene.sensors()[0]
    film = sensor.film()
    film.prepare()
    dev = integ.device_scene(scene, sensor)
    desc = scene.describe(sensor)
    ref_film = np.zeros((16, 24, 96, 4), np.float32); ref_steady = np.zeros((16, 24, 3), np.float32)
    for i, (seed_i, spp_i) in enumerate(((123, 8), (77, 5))):
        p = integ.make_params(spp_i, 13, seed_i, accumulate=(i > 0))
        _lib.check(_lib.load().tr_render(dev.handle, C.byref(p), film.transient_storage.data_ptr(), film.steady_storage.data_ptr(),
                                         None, _lib.current_stream_ptr()))
        cnt = mitr._abi.TrCounters()
        o._check(o.load().oracle_render(C.byref(desc), C.byref(p), ref_film.ctypes.data, ref_steady.ctypes.data, C.byref(cnt), 0, None, 0, None))
    raw = film.develop(raw=True)[1].cpu().numpy()
    err = film_l1(raw, ref_film)
    assert err["zero_bins_mismatch"] == 0 and err["whole"] <= WHOLE_TOL, err
    np.testing.assert_allclose(film.steady_storage.cpu().numpy(), ref_steady, rtol=1e-4, atol=1e-6)


def test_contract_reciprocal_is_correctly_rounded(mitr):
    """The ray / primitive tests divide through rcp_c(x) = the correctly rounded 1/x (oracle: rcp_c).  The kernels compute
    it as MUFU.RCP + one Newton step without nvcc's range guard: checked here against the IEEE division over EVERY fp32
    x with 2^-126 <= |x| < 2^126 (tr_selftest_rcp), so the oracle's `1.0f / x` and the kernels agree bit for bit."""
    import ctypes as C
    from mitransient_b200 import _lib
    n, bad = C.c_uint64(0), C.c_uint64(0)
    _lib.check(_lib.load().tr_selftest_rcp(C.byref(n), C.byref(bad), _lib.current_stream_ptr()))
    assert n.value == 2 * 252 * (1 << 23), n.value
    assert bad.value == 0, f"{bad.value} of {n.value} reciprocals differ from 1.0f / x"


@pytest.mark.parametrize("subdiv,in_smem", [(1, True), (3, False)])
def test_bvh_icosphere_parity(mitr, oracle, subdiv, in_smem):
    """The SAH BVH path on both sides of the 48 KB blob limit: 87 primitives (whole blob staged in shared memory) and
    1287 primitives (blob in global memory, the breadth-first top of the tree staged by the TMA copy) must give the
    oracle's brute-force hits -- closest hits AND shadow rays: equal counters, film relative L1 <= 1e-5."""
    scene = mitr.load_dict(mitr.cornell_box_sphere_config(40, 32, 160, 8, subdiv))
    integ = scene.integrator()
    integ.collect_counters = True
    steady, transient = integ.render(scene, seed=6, spp=8)
    info = integ.device_scene(scene, scene.sensors()[0]).info()
    assert info["bvh_nodes"] > 0 and (info["blob_bytes"] <= 48 * 1024) == in_smem, info
    raw = scene.sensors()[0].film().develop(raw=True)[1].cpu().numpy()
    ref = oracle.render(scene, integ, seed=6, spp=8)
    assert integ.last_counters == ref["counters"], (integ.last_counters, ref["counters"])
    err = film_l1(raw, ref["transient"])
    assert err["zero_bins_mismatch"] == 0 and err["whole"] <= WHOLE_TOL and err["max_t"] <= PER_BIN_TOL, err
    integ.collect_counters = False                          # the plain 768-thread instantiation
    integ.render(scene, seed=6, spp=8)
    raw2 = scene.sensors()[0].film().develop(raw=True)[1].cpu().numpy()
    assert film_l1(raw2, ref["transient"])["whole"] <= WHOLE_TOL


def test_mesh_cubes_as_box_primitives_parity(mitr, oracle):
    """`merge_box` at ingestion: the mesh Cornell box with its closed cubes turned into box primitives (8 primitives)
    renders like the oracle on the same descriptor."""
    scene = mitr.load_dict(mitr.cornell_box_mesh_config(40, 32, 160, 8, merge_boxes=True))
    assert scene.describe(scene.sensors()[0]).n_prims == 8
    integ = scene.integrator()
    integ.collect_counters = True
    integ.render(scene, seed=2, spp=8)
    raw = scene.sensors()[0].film().develop(raw=True)[1].cpu().numpy()
    ref = oracle.render(scene, integ, seed=2, spp=8)
    assert integ.last_counters == ref["counters"]
    assert film_l1(raw, ref["transient"])["whole"] <= WHOLE_TOL


def test_scene_refresh_reuploads_the_same_scene(mitr):
    """tr_scene_refresh (the e2e leg's per-step H2D input): a render after it equals the render before it."""
    import ctypes as C
    import torch
    from mitransient_b200 import _lib
    scene = mitr.load_dict(mitr.cornell_box_config(24, 16, 64, 4))
    integ, sensor = scene.integrator(), scene.sensors()[0]
    _, t0 = integ.render(scene, seed=1, spp=4)
    t0 = t0.clone()
    dev = integ.device_scene(scene, sensor)
    _lib.check(_lib.load().tr_scene_refresh(dev.handle, _lib.current_stream_ptr()))
    _, t1 = integ.render(scene, seed=1, spp=4)
    assert torch.equal(t0, t1)


def test_histogram_path_preference_is_honoured_and_reported(mitr, oracle, hist_path):
    """TrRenderParams.histogram_path: the diffuse Cornell box can run on either path (tr_last_render_info says which
    one did), with or without specular materials."""
    scene = mitr.load_dict(mitr.cornell_box_config(16, 16, 64, 8))
    integ = scene.integrator()
    assert integ.histogram_path == hist_path
    integ.collect_counters = True
    integ.render(scene, seed=1, spp=8)
    info = integ.last_render_info()
    assert info["histogram_path"] == {"auto": "global", "shared": "shared", "global": "global"}[hist_path], info
    assert info["threads"] == 512 and (info["slots"] >= 2) == (info["histogram_path"] == "shared")
    raw = scene.sensors()[0].film().develop(raw=True)[1].cpu().numpy()
    ref = oracle.render(scene, integ, seed=1, spp=8)
    assert integ.last_counters == ref["counters"]
    err = film_l1(raw, ref["transient"])
    assert err["whole"] <= WHOLE_TOL and err["max_t"] <= PER_BIN_TOL and err["zero_bins_mismatch"] == 0, err

    # a mirror box: the general-material kernels exist for both paths too
    d = mitr.cornell_box_config(16, 16, 64, 8)
    d["large-box"]["bsdf"] = {"type": "conductor"}
    scene = mitr.load_dict(d)
    integ = scene.integrator()
    integ.collect_counters = True
    integ.render(scene, seed=1, spp=8)
    assert integ.last_render_info()["histogram_path"] == info["histogram_path"]
    raw = scene.sensors()[0].film().develop(raw=True)[1].cpu().numpy()
    ref = oracle.render(scene, integ, seed=1, spp=8)
    assert integ.last_counters == ref["counters"]
    err = film_l1(raw, ref["transient"])
    assert err["whole"] <= WHOLE_TOL and err["max_t"] <= PER_BIN_TOL and err["zero_bins_mismatch"] == 0, err
