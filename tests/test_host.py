"""Host-side logic (no GPU): scene ingestion, plugin parameters and the reference's accept/reject behaviour."""
import ctypes as C
import math

import numpy as np
import pytest


def test_plugin_names_and_defaults(mitr):
    from mitransient_b200 import plugins
    assert {"transient_path"} <= set(plugins.INTEGRATORS)
    assert "transient_hdr_film" in plugins.FILMS and "nlos_capture_meter" in plugins.SENSORS
    film = mitr.load_dict({"type": "transient_hdr_film"})
    # transient_hdr_film.py:116-121 + mi.Film defaults
    assert (film.temporal_bins, film.bin_width_opl, film.start_opl) == (2048, 0.003, 0.0)
    assert film.size() == (768, 576) and film.exhaustive_scan is False and film.rfilter() == "box"
    assert abs(film.end_opl() - 2048 * 0.003) < 1e-12
    integ = mitr.load_dict({"type": "transient_path"})
    assert (integ.max_depth, integ.rr_depth, integ.camera_unwarp, integ.discard_direct_light) == (6, 5, False, False)
    assert mitr.load_dict({"type": "transient_path", "max_depth": -1}).max_depth == 0xFFFFFFFF
    # temporal_filter / gaussian_stddev / block_size are accepted and ignored (common.py:28-30)
    mitr.load_dict({"type": "transient_path", "temporal_filter": "gaussian", "gaussian_stddev": 3.0, "block_size": 32})
    with pytest.raises(ValueError):
        mitr.load_dict({"type": "transient_path", "rr_depth": 0})


def test_cornell_box_scene(mitr):
    d = mitr.cornell_box()
    assert d["integrator"]["max_depth"] == 8 and d["sensor"]["film"]["temporal_bins"] == 300
    scene = mitr.load_dict(d)
    assert len(scene.shapes()) == 8 and len(scene.emitters()) == 1 and len(scene.bsdfs_) == 3
    assert sum(len(s.prims) for s in scene.shapes()) == 6 + 2       # 6 rectangles + 2 cubes (one box primitive each)
    light = scene.shapes()[0]
    assert light.kind == "rectangle" and light.emitter_index == 0
    assert abs(light.surface_area() - (2 * 0.23) * (2 * 0.19)) < 1e-9
    n = np.cross(light.prims[0, 1] - light.prims[0, 0], light.prims[0, 2] - light.prims[0, 0])
    assert n[1] < 0 and abs(n[0]) < 1e-12 and abs(n[2]) < 1e-12     # the light faces down
    np.testing.assert_allclose(light.prims[0, 3], light.prims[0, 1] + light.prims[0, 2] - light.prims[0, 0], atol=1e-12)
    for s in scene.shapes()[1:6]:                                    # walls face the inside of the box
        c = s.prims[0, :3].mean(axis=0)
        nn = np.cross(s.prims[0, 1] - s.prims[0, 0], s.prims[0, 2] - s.prims[0, 0])
        assert np.dot(nn, -c) > 0, s.name
    small, large = scene.shapes()[6:]
    assert small.prim_types.tolist() == [2] and abs(small.surface_area() - 6 * 0.36) < 1e-9
    e = large.prims[0, 1:] - large.prims[0, 0]
    assert np.linalg.det(e) > 0 and np.allclose(sorted(np.linalg.norm(e, axis=1)), [0.6, 0.6, 1.22])
    # the same cubes as six outward-facing quads (box_primitive = False)
    d6 = mitr.cornell_box()
    d6["small-box"]["box_primitive"] = False
    s6 = mitr.load_dict(d6).shapes()[6]
    assert len(s6.prims) == 6 and abs(s6.surface_area() - small.surface_area()) < 1e-9
    centre = s6.prims[:, :3].reshape(-1, 3).mean(axis=0)
    for p in s6.prims:
        fc = p[0] + 0.5 * (p[1] - p[0]) + 0.5 * (p[2] - p[0])
        assert np.dot(np.cross(p[1] - p[0], p[2] - p[0]), fc - centre) > 0
    desc = scene.describe(scene.sensors()[0])
    assert desc.n_prims == 8 and desc.n_shapes == 8 and desc.film.temporal_bins == 300
    cfg = mitr.cornell_box_config(512, 512, 1024, 256)
    assert cfg["sensor"]["film"]["bin_width_opl"] == 6.0 / 1024 and cfg["sensor"]["film"]["start_opl"] == 3.5


def test_transforms(mitr):
    T = mitr.ScalarTransform4f
    m = T().translate([1, 2, 3]).rotate([0, 0, 1], 90).scale(2.0)
    np.testing.assert_allclose(m.transform_affine([1, 0, 0]), [1, 4, 3], atol=1e-12)   # scale, rotate, translate
    np.testing.assert_allclose((m @ m.inverse()).matrix, np.eye(4), atol=1e-12)
    la = T().look_at(origin=[0, 0, 3.9], target=[0, 0, 0], up=[0, 1, 0])
    np.testing.assert_allclose(la.matrix[:3, :3], [[-1, 0, 0], [0, 1, 0], [0, 0, -1]], atol=1e-12)
    from mitransient_b200.transforms import parse_fov
    assert parse_fov(40.0, "x", 2.0) == 40.0
    assert abs(parse_fov(40.0, "smaller", 2.0) - math.degrees(2 * math.atan(math.tan(math.radians(20)) * 2))) < 1e-9
    assert parse_fov(40.0, "smaller", 0.5) == 40.0


def test_rejections_match_reference(mitr):
    d = mitr.cornell_box_config(8, 8, 16, 1)
    scene = mitr.load_dict(d)
    integ = scene.integrator()
    with pytest.raises(Exception, match="develop=True"):
        integ.render(scene, develop=False)                           # common.py:130-132
    d2 = mitr.cornell_box_config(8, 8, 16, 1)
    d2["sensor"]["film"]["rfilter"] = {"type": "gaussian"}
    s2 = mitr.load_dict(d2)
    with pytest.raises(ValueError, match="rfilter"):
        s2.sensors()[0].film().check_supported()                     # transient_image_block.py:150-151
    with pytest.raises(ValueError):
        mitr.load_dict({"type": "scene", "x": {"type": "sphere"}})
    with pytest.raises(ValueError):
        mitr.load_dict({"type": "scene", "x": {"type": "rectangle", "bsdf": {"type": "roughplastic"}}})


def test_prepare_pass_split(mitr):
    """common.py:51-85: one pass up to 2^32 samples, else passes of int((2^26-1)/(W*H)) spp with seeds drawn
    from the sampler."""
    from mitransient_b200.integrators.common import TransientADIntegrator
    from mitransient_b200.sampler import independent_sampler_first_float

    class FakeFilm:
        def crop_size(self): return (512, 512)
        def prepare(self, aovs): pass

    class FakeSensor:
        def __init__(self, s): self._s = s
        def film(self): return FakeFilm()
        def sampler(self): return self._s

    from mitransient_b200.scene import IndependentSampler, Properties
    integ = TransientADIntegrator(Properties())
    s = IndependentSampler(Properties({"sample_count": 256}))
    assert integ.prepare(None, FakeSensor(s), seed=7, spp=0) == [(7, 256)]
    assert integ.prepare(None, FakeSensor(s), seed=0, spp=16384) == [(0, 16384)]          # exactly 2^32: one pass
    passes = integ.prepare(None, FakeSensor(s), seed=3, spp=20000)
    spp_per_pass = int((2 ** 26 - 1) / (512 * 512))
    assert spp_per_pass == 255 and sum(p[1] for p in passes) == 20000
    assert len(passes) == 20000 // 255 + 1 and passes[-1][1] == 20000 % 255
    assert passes[0][0] == int(independent_sampler_first_float(3, 0) * 2.0 ** 32)
    assert len({p[0] for p in passes}) == len(passes)


def test_obj_loader_ignores_line_records(mitr, tmp_path):
    p = tmp_path / "q.obj"
    p.write_text("o q\nv 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nvn 0 0 1\nl 1 2\nf 1/1/1 2/1/1 3/1/1 4/1/1\nf -4 -3 -2\n")
    tris = mitr.load_obj(str(p))
    assert tris.shape == (3, 3, 3)                                   # quad -> 2 triangles (fan) + 1 triangle
    np.testing.assert_allclose(tris[2], [[0, 0, 0], [1, 0, 0], [1, 1, 0]])


_XML = """<scene version="3.3.0">
    <default name="spp" value="4"/>
    <default name="res" value="16"/>
    <integrator type="transient_path"><integer name="max_depth" value="5"/><boolean name="camera_unwarp" value="false"/></integrator>
    <sensor type="perspective" id="sensor">
        <string name="fov_axis" value="smaller"/><float name="near_clip" value="0.01"/><float name="far_clip" value="100"/>
        <float name="fov" value="40"/>
        <transform name="to_world"><lookat origin="0, 0, 4" target="0, 0, 0" up="0, 1, 0"/></transform>
        <sampler type="independent"><integer name="sample_count" value="$spp"/></sampler>
        <film type="transient_hdr_film">
            <integer name="width" value="$res"/><integer name="height" value="$res"/>
            <integer name="temporal_bins" value="32"/><float name="start_opl" value="2"/><float name="bin_width_opl" value="0.25"/>
            <rfilter type="box"/>
        </film>
    </sensor>
    <bsdf type="diffuse" id="white"><rgb name="reflectance" value="0.8, 0.7, 0.6"/></bsdf>
    <shape type="obj" id="wall"><string name="filename" value="quad.obj"/><ref id="white"/></shape>
    <shape type="rectangle" id="light">
        <transform name="to_world"><scale value="0.25"/><rotate x="1" angle="180"/><translate x="0" y="0" z="3"/></transform>
        <bsdf type="twosided"><bsdf type="diffuse"><rgb name="reflectance" value="0.5"/></bsdf></bsdf>
        <emitter type="area"><rgb name="radiance" value="10, 9, 8"/></emitter>
    </shape>
</scene>"""


def test_xml_scene_file(mitr, tmp_path):
    (tmp_path / "quad.obj").write_text("v -1 -1 0\nv 1 -1 0\nv 1 1 0\nv -1 1 0\nf 1 2 3\nf 1 3 4\n")
    (tmp_path / "scene.xml").write_text(_XML)
    d = mitr.xml_to_dict(str(tmp_path / "scene.xml"), res=24)
    assert d["sensor"]["film"]["width"] == 24 and d["sensor"]["sampler"]["sample_count"] == 4   # <default> + override
    scene = mitr.load_file(str(tmp_path / "scene.xml"), res=24)
    assert scene.integrator().max_depth == 5 and scene.sensors()[0].film().size() == (24, 24)
    assert len(scene.shapes()) == 2 and len(scene.emitters()) == 1
    wall, light = scene.shapes()
    assert len(wall.prims) == 1 and wall.prim_types[0] == 1 and wall.prims.shape == (1, 4, 3)   # two coplanar triangles -> one quad
    # transform ops apply in file order: scale, then rotate, then translate
    np.testing.assert_allclose(light.to_world.transform_affine([1, 1, 0]), [0.25, -0.25, 3.0], atol=1e-12)
    n = np.cross(light.prims[0, 1] - light.prims[0, 0], light.prims[0, 2] - light.prims[0, 0])
    assert n[2] < 0                                                            # rotated to face the wall
    assert scene.bsdfs_[light.bsdf_index]["flags"] == 1                         # twosided
    with pytest.raises(ValueError):
        bad = _XML.replace('type="diffuse" id="white"', 'type="roughconductor" id="white"')
        (tmp_path / "bad.xml").write_text(bad)
        mitr.load_file(str(tmp_path / "bad.xml"))


def test_merge_quads(mitr):
    from mitransient_b200.scene import merge_quads
    v = np.array([[0, 0, 0], [2, 0, 0], [2, 1, 0], [0, 1, 0], [3, 3, 0]], dtype=float)
    tris = v[np.array([[0, 1, 2], [0, 2, 3], [1, 4, 2]])]        # a rectangle (2 tris) + a lone triangle
    prims, types = merge_quads(tris)
    assert types.tolist() == [1, 0] and prims.shape == (2, 3, 3)
    p0, p1, p2 = prims[0][:3]
    corners = {tuple(p0), tuple(p1), tuple(p2), tuple(p1 + p2 - p0)}
    assert corners == {tuple(x) for x in v[:4]}
    assert np.cross(p1 - p0, p2 - p0)[2] > 0                      # orientation of the source triangles is kept
    trap = v[np.array([[0, 1, 2], [0, 2, 3]])].copy()
    trap[1, 2] = [0.5, 1, 0]                                      # a trapezoid is not a parallelogram: not merged
    assert merge_quads(trap)[1].tolist() == [0, 0]


def test_phasor_film_frequencies_match_reference_formula(mitr):
    """phasor_hdr_film.py:125-137 on the documented example (400 bins x 6.5, wl_mean = wl_sigma = 100)."""
    film = mitr.load_dict({"type": "phasor_hdr_film", "wl_mean": 100, "wl_sigma": 100, "width": 256, "height": 256,
                           "temporal_bins": 400, "start_opl": 1000, "bin_width_opl": 6.5, "rfilter": {"type": "box"}})
    nt, bw = 400, 6.5
    mean_idx, sigma_idx = nt * bw / 100, nt * bw / 600
    lo, hi = max(0, int(np.floor(mean_idx - 3 * sigma_idx))), min(nt // 2, int(np.ceil(mean_idx + 3 * sigma_idx)))
    want = np.fft.fftfreq(nt, d=bw)[lo:hi + 1].astype(np.float32)
    np.testing.assert_array_equal(film.frequencies, want)
    assert len(want) == 27 and film.raw_channels() == 55 and film.channels[0] == "L_fq000_Re" and film.channels[-1] == "W"
    d = film.describe()
    assert d.type == mitr._abi.TR_FILM_PHASOR_HDR and d.n_frequencies == 27 and d.frequencies[26] == want[26]
    defaults = mitr.load_dict({"type": "phasor_hdr_film"})
    assert defaults.temporal_bins == 4096 and defaults.wl_mean == 100.0 and defaults.wl_sigma == 1000.0
    with pytest.raises(NotImplementedError):
        film.create_block()


def test_variant_selects_colour_mode(mitr):
    from mitransient_b200 import colour_mode
    try:
        assert mitr.set_variant("llvm_ad_mono") == "cuda_mono" and colour_mode.is_monochromatic()
        scene = mitr.load_dict(mitr.cornell_box_config(8, 8, 16, 1))
        desc = scene.describe(scene.sensors()[0])
        for i in range(desc.n_bsdfs):
            r = list(desc.bsdfs[i].reflectance)
            assert r[0] == r[1] == r[2]
        green = np.float32(0.212671) * np.float32(0.105421) + np.float32(0.715160) * np.float32(0.37798) \
            + np.float32(0.072169) * np.float32(0.076425)
        assert any(abs(desc.bsdfs[i].reflectance[0] - green) < 1e-7 for i in range(desc.n_bsdfs))
        assert mitr.set_variant("cuda_ad_rgb", "llvm_ad_rgb") == "cuda_rgb" and not colour_mode.is_monochromatic()
        desc = scene.describe(scene.sensors()[0])
        assert any(desc.bsdfs[i].reflectance[0] != desc.bsdfs[i].reflectance[1] for i in range(desc.n_bsdfs))
    finally:
        mitr.set_variant("cuda_rgb")


def test_exhaustive_scan_film_descriptor(mitr):
    film = mitr.load_dict({"type": "transient_hdr_film", "width": 4, "height": 4, "temporal_bins": 8,
                           "exhaustive_scan": True, "laser_scan_width": 4, "laser_scan_height": 4})
    d = film.describe()
    assert (d.laser_scan_width, d.laser_scan_height, d.type) == (4, 4, mitr._abi.TR_FILM_TRANSIENT_HDR)
    assert film.needs_zeroed_accumulate() and film.raw_channels() == 4
    plain = mitr.load_dict({"type": "transient_hdr_film", "width": 4, "height": 4, "temporal_bins": 8})
    assert plain.describe().laser_scan_width == 0 and not plain.needs_zeroed_accumulate()
    bad = mitr.load_dict({"type": "transient_hdr_film", "width": 4, "height": 4, "exhaustive_scan": True})
    with pytest.raises(ValueError, match="laser_scan"):
        bad.check_supported()


def test_tonemap_transient_matches_numpy_quantile(mitr):
    import torch
    rng = np.random.default_rng(0)
    x = (rng.standard_normal((6, 5, 7, 3)) ** 3).astype(np.float32)
    want = x / np.quantile(np.abs(x), 0.99) * 2.0
    np.testing.assert_allclose(mitr.vis.tonemap_transient(x, scaling=2.0), want, rtol=1e-6)
    got = mitr.vis.tonemap_transient(torch.from_numpy(x), scaling=2.0)
    np.testing.assert_allclose(got.numpy(), want, rtol=1e-5)


def test_ply_loader_ascii_and_binary(mitr, tmp_path):
    """`ply` shapes (examples/transient-nlos/nlos-z-room.xml): binary little/big endian and ascii, extra vertex
    properties skipped, polygons fan-triangulated."""
    import struct
    verts = [(-1.0, -1.0, 0.0, 0.0, 0.0), (1.0, -1.0, 0.0, 1.0, 0.0), (1.0, 1.0, 0.0, 1.0, 1.0), (-1.0, 1.0, 0.0, 0.0, 1.0),
             (0.0, 2.0, 0.5, 0.5, 0.5)]
    faces = [(0, 1, 2, 3), (3, 2, 4)]
    header = ("ply\nformat {fmt} 1.0\ncomment made by a test\nelement vertex 5\nproperty float x\nproperty float y\n"
              "property float z\nproperty float s\nproperty float t\nelement face 2\n"
              "property list uchar int vertex_indices\nend_header\n")
    want = np.array([[verts[a][:3], verts[b][:3], verts[c][:3]] for a, b, c in ((0, 1, 2), (0, 2, 3), (3, 2, 4))])
    for fmt, e in (("binary_little_endian", "<"), ("binary_big_endian", ">")):
        path = tmp_path / f"{fmt}.ply"
        with open(path, "wb") as f:
            f.write(header.format(fmt=fmt).encode())
            for v in verts:
                f.write(struct.pack(e + "5f", *v))
            for fc in faces:
                f.write(struct.pack(e + "B" + f"{len(fc)}i", len(fc), *fc))
        np.testing.assert_allclose(mitr.load_ply(str(path)), want)
    path = tmp_path / "ascii.ply"
    with open(path, "w") as f:
        f.write(header.format(fmt="ascii"))
        for v in verts:
            f.write(" ".join(str(x) for x in v) + "\n")
        for fc in faces:
            f.write(f"{len(fc)} " + " ".join(str(i) for i in fc) + "\n")
    np.testing.assert_allclose(mitr.load_ply(str(path)), want)
    scene = mitr.load_dict({"type": "scene", "m": {"type": "ply", "filename": str(path), "bsdf": {"type": "diffuse"}}})
    shape = scene.shapes()[0]
    assert len(shape.prims) == 2 and sorted(shape.prim_types.tolist()) == [0, 1]     # the quad was merged, the triangle kept


def test_crop_window_film_descriptor(mitr):
    """A crop window renders crop_size() pixels (film.describe: width / height = the window, plus its offset); the
    exhaustive layout and the NLOS sensor reject it like the reference's documentation says / code implies."""
    props = {"type": "transient_hdr_film", "width": 12, "height": 10, "temporal_bins": 8, "crop_width": 5, "crop_height": 4,
             "crop_offset_x": 6, "crop_offset_y": 3}
    film = mitr.load_dict(props)
    d = film.describe()
    assert (d.width, d.height, d.crop_offset_x, d.crop_offset_y) == (5, 4, 6, 3)
    assert film.size() == (12, 10) and film.crop_size() == (5, 4) and film.is_cropped()
    film.check_supported()
    plain = mitr.load_dict({"type": "transient_hdr_film", "width": 12, "height": 10, "temporal_bins": 8})
    assert not plain.is_cropped() and (plain.describe().crop_offset_x, plain.describe().crop_offset_y) == (0, 0)
    with pytest.raises(ValueError, match="crop window must lie inside"):
        mitr.load_dict({**props, "crop_offset_x": 9}).check_supported()
    with pytest.raises(ValueError, match="exhaustive_scan"):
        mitr.load_dict({**props, "exhaustive_scan": True, "laser_scan_width": 2, "laser_scan_height": 2}).check_supported()
    d = mitr.nlos_z_scene(8, 8, 16, 4)
    d["relay_wall"]["nlos_sensor"]["transient_film"].update(crop_width=4, crop_height=4)
    with pytest.raises(NotImplementedError, match="crop window"):
        mitr.load_dict(d)


def _fake_mitsuba(mitr):
    """The few Mitsuba objects mitsuba_adapter touches, as plain Python (Mitsuba itself cannot be installed here)."""
    import types
    T = mitr.ScalarTransform4f

    class Obj:
        def __init__(self, kind, params, ident=""): self._kind, self._params, self._id = kind, params, ident
        def __str__(self): return self._kind + "[\n]"
        def id(self): return self._id

    class Shape(Obj):
        def __init__(self, kind, params, ident, bsdf, emitter=None, mesh=False):
            super().__init__(kind, params, ident); self._bsdf, self._emitter, self._mesh = bsdf, emitter, mesh
        def is_mesh(self): return self._mesh
        def bsdf(self): return self._bsdf
        def emitter(self): return self._emitter
        def interior_medium(self): return None
        def exterior_medium(self): return None

    class Emitter(Obj):
        def __init__(self, kind, params, shape=None): super().__init__(kind, params); self._shape = shape
        def get_shape(self): return self._shape

    class Film:
        temporal_bins, start_opl, bin_width_opl, exhaustive_scan = 32, 3.5, 0.25, False
        def size(self): return (8, 6)
        def crop_size(self): return (8, 6)
        def crop_offset(self): return (0, 0)
        def rfilter(self): return types.SimpleNamespace(is_box_filter=lambda: True)

    class Sensor(Obj):
        def __init__(self, params): super().__init__("PerspectiveCamera", params); self._film = Film()
        def film(self): return self._film
        def sampler(self): return types.SimpleNamespace(sample_count=lambda: 4)

    white = Obj("SmoothDiffuse", {"reflectance.value": [0.8, 0.8, 0.8]})
    green = Obj("TwoSided", {"brdf_0.reflectance.value": [0.1, 0.9, 0.2]})
    light_shape = Shape("Rectangle", {"to_world": T().translate([0, 0.99, 0]).rotate([1, 0, 0], 90).scale([0.2, 0.2, 1])}, "light", white)
    light = Emitter("AreaLight", {"radiance.value": [18.0, 14.0, 7.0]}, light_shape)
    light_shape._emitter = light
    tri_v = np.array([[-1, -1, -1], [1, -1, -1], [1, -1, 1], [-1, -1, 1]], dtype=np.float32)
    floor = Shape("OBJMesh", {"vertex_positions": tri_v.reshape(-1), "faces": np.array([0, 1, 2, 0, 2, 3], dtype=np.uint32)}, "floor",
                  green, mesh=True)
    back = Shape("Rectangle", {"to_world": T().translate([0, 0, -1])}, "back", white)
    sensor = Sensor({"to_world": T().look_at(origin=[0, 0, 3.9], target=[0, 0, 0], up=[0, 1, 0]), "x_fov": 39.3, "near_clip": 0.01,
                     "far_clip": 100.0})
    scene = types.SimpleNamespace(shapes=lambda: [floor, back, light_shape], emitters=lambda: [light], sensors=lambda: [sensor])
    registered = {}
    mi = types.ModuleType("mitsuba")
    mi.traverse = lambda o: o._params
    mi.variant = lambda: "cuda_ad_rgb"
    mi.SamplingIntegrator = type("SamplingIntegrator", (), {"__init__": lambda self, props: None})
    mi.register_integrator = lambda name, factory: registered.__setitem__(name, factory)
    mi.TensorXf = lambda t: t
    return mi, scene, sensor, registered


def test_mitsuba_adapter_exports_a_scene_the_loader_accepts(mitr, monkeypatch):
    """mitsuba_adapter.scene_to_dict over a stand-in of the Mitsuba objects it reads (NOT a live Mitsuba: see the
    module's status note): the exported dictionary loads, and describes the same primitives, BSDFs, emitter, sensor and
    film as the same scene written as a dictionary by hand."""
    import sys
    from mitransient_b200 import mitsuba_adapter as ad
    mi, scene, sensor, registered = _fake_mitsuba(mitr)
    monkeypatch.setitem(sys.modules, "mitsuba", mi)
    assert ad.available()
    d = ad.scene_to_dict(scene, sensor, {"type": "transient_path", "max_depth": 5})
    ours = mitr.load_dict(d)
    T = mitr.ScalarTransform4f
    hand = mitr.load_dict({
        "type": "scene", "integrator": {"type": "transient_path", "max_depth": 5},
        "floor": {"type": "trimesh", "vertices": np.array([[-1, -1, -1], [1, -1, -1], [1, -1, 1], [-1, -1, 1]], dtype=np.float64),
                  "faces": np.array([[0, 1, 2], [0, 2, 3]]),
                  "bsdf": {"type": "twosided", "bsdf": {"type": "diffuse", "reflectance": {"type": "rgb", "value": [0.1, 0.9, 0.2]}}}},
        "back": {"type": "rectangle", "to_world": T().translate([0, 0, -1]),
                 "bsdf": {"type": "diffuse", "reflectance": {"type": "rgb", "value": [0.8, 0.8, 0.8]}}},
        "light": {"type": "rectangle", "to_world": T().translate([0, 0.99, 0]).rotate([1, 0, 0], 90).scale([0.2, 0.2, 1]),
                  "bsdf": {"type": "diffuse", "reflectance": {"type": "rgb", "value": [0.8, 0.8, 0.8]}},
                  "emitter": {"type": "area", "radiance": {"type": "rgb", "value": [18.0, 14.0, 7.0]}}},
        "sensor": {"type": "perspective", "to_world": T().look_at(origin=[0, 0, 3.9], target=[0, 0, 0], up=[0, 1, 0]), "fov": 39.3,
                   "near_clip": 0.01, "far_clip": 100.0, "sampler": {"type": "independent", "sample_count": 4},
                   "film": {"type": "transient_hdr_film", "width": 8, "height": 6, "temporal_bins": 32, "start_opl": 3.5,
                            "bin_width_opl": 0.25, "rfilter": {"type": "box"}}}})
    a, b = ours.describe(ours.sensors()[0]), hand.describe(hand.sensors()[0])
    assert (a.n_prims, a.n_shapes, a.n_bsdfs, a.n_emitters) == (b.n_prims, b.n_shapes, b.n_bsdfs, b.n_emitters) == (3, 3, b.n_bsdfs, 1)
    assert np.allclose(np.ctypeslib.as_array(a.prim_verts, (a.n_prims * 12,)), np.ctypeslib.as_array(b.prim_verts, (b.n_prims * 12,)))
    assert list(a.sensor.sample_to_camera) == list(b.sensor.sample_to_camera) and list(a.sensor.to_world) == list(b.sensor.to_world)
    assert (a.film.width, a.film.height, a.film.temporal_bins, a.film.start_opl) == (8, 6, 32, 3.5)
    assert list(a.emitters[0].radiance) == list(b.emitters[0].radiance)
    assert ours.integrator().max_depth == 5
    # specular BSDFs
    scene.shapes()[1]._bsdf = type(scene.shapes()[1]._bsdf)("SmoothConductor", {"eta.value": [0.14, 0.37, 1.44], "k.value": [3.98, 2.38, 1.6],
                                                                                 "specular_reflectance.value": [1.0, 1.0, 1.0]})
    scene.shapes()[0]._bsdf = type(scene.shapes()[1]._bsdf)("SmoothDielectric", {"eta": 1.5})
    d2 = ad.scene_to_dict(scene, sensor, {"type": "transient_path"})
    assert d2["back"]["bsdf"]["type"] == "conductor" and d2["back"]["bsdf"]["k"]["value"] == [3.98, 2.38, 1.6]
    assert d2["floor"]["bsdf"] == {"type": "dielectric", "int_ior": 1.5, "ext_ior": 1.0, "specular_reflectance": 1.0, "specular_transmittance": 1.0}
    s2 = mitr.load_dict(d2)
    assert s2.describe(s2.sensors()[0]).n_bsdfs >= 3
    # unsupported content is refused, not guessed
    scene.shapes()[1]._bsdf._kind = "RoughConductor"
    with pytest.raises(NotImplementedError, match="RoughConductor"):
        ad.scene_to_dict(scene, sensor, {"type": "transient_path"})
    scene.shapes()[1]._bsdf._kind = "SmoothDiffuse"
    assert ad.register(prefix="b200_") == ["b200_transient_path", "b200_transient_nlos_path", "b200_transient_prbvolpath"]
    integ = registered["b200_transient_path"]({"max_depth": 7})
    assert integ._impl.max_depth == 7
    with pytest.raises(Exception, match="develop=True"):
        integ.render(scene, develop=False)


def test_dict_to_xml_round_trip(mitr, tmp_path):
    """`dict_to_xml` writes what `load_file` reads: the same flat scene descriptor (vertices, sensor, film) and the same
    integrator as `load_dict` of the dictionary -- for the four benchmark scenes, meshes going through OBJ files."""
    import ctypes as C
    from mitransient_b200.xml import dict_to_xml

    def flat(scene):
        d = scene.describe(scene.sensors()[0])
        v = np.ctypeslib.as_array(C.cast(d.prim_verts, C.POINTER(C.c_float)), shape=(d.n_prims, 12)).copy()
        return v, bytes(d.sensor), bytes(d.film), d.n_shapes, d.n_bsdfs, d.n_emitters, d.n_media

    cases = {"c2": lambda: mitr.cornell_box_config(32, 32, 64, 4), "vol": lambda: mitr.cornell_box_volumetric(16, 24, 32, 4),
             "nlos": lambda: mitr.nlos_z_scene(16, 16, 64, 4), "mesh": lambda: mitr.cornell_box_mesh_config(16, 16, 32, 4)}
    for name, make in cases.items():          # (load_dict consumes its dictionary: build a fresh one per use)
        s0 = mitr.load_dict(make())
        a = flat(s0)
        s1 = mitr.load_file(dict_to_xml(make(), str(tmp_path / (name + ".xml"))))
        b = flat(s1)
        assert np.array_equal(a[0], b[0]) and a[1:] == b[1:], name
        assert s1.integrator().__dict__ == s0.integrator().__dict__
    assert (tmp_path / "mesh_1_light.obj").exists() or any(p.suffix == ".obj" for p in tmp_path.iterdir())


def test_multi_pass_seeds_include_the_sampler_seed(mitr, monkeypatch):
    """common.py:70-80: per-pass seeds are drawn from a sampler seeded with `seed` (+ the sampler's own `seed` property)
    and handed to fresh clones of the sensor's sampler, whose seed() adds that property again."""
    from mitransient_b200.films.transient_hdr_film import TransientHDRFilm
    from mitransient_b200.sampler import independent_sampler_first_float
    monkeypatch.setattr(TransientHDRFilm, "prepare", lambda self, aovs=(): 4)
    for base in (0, 7):
        d = mitr.cornell_box_config(512, 512, 2, 4)
        d["sensor"]["sampler"]["seed"] = base
        scene = mitr.load_dict(d)
        passes = scene.integrator().prepare(scene, scene.sensors()[0], 3, 2 ** 15)
        assert len(passes) == 129
        for i, (seed_i, _) in enumerate(passes[:5]):
            u = independent_sampler_first_float((3 + base) & 0xFFFFFFFF, i)
            assert seed_i == (int(u * 2.0 ** 32) + base) & 0xFFFFFFFF


def test_closed_mesh_cubes_become_box_primitives(mitr):
    """`merge_box`: six outward parallelograms of one parallelepiped (an OBJ cube after `merge_quads`) -> one box primitive;
    emitters, open or inward-facing meshes and `box_primitive=False` keep their parallelograms."""
    from mitransient_b200.scene import merge_box, merge_quads
    s18 = mitr.load_dict(mitr.cornell_box_mesh_config(8, 8, 4, 1))
    assert s18.describe(s18.sensors()[0]).n_prims == 18
    s8 = mitr.load_dict(mitr.cornell_box_mesh_config(8, 8, 4, 1, merge_boxes=True))
    assert s8.describe(s8.sensors()[0]).n_prims == 8
    v, f = mitr.utils._unit_mesh("cube")
    tris = np.array(v)[np.array(f)]
    q, t = merge_quads(tris)
    box, bt = merge_box(q, t)
    assert box.shape == (1, 4, 3) and abs(abs(np.linalg.det(box[0, 1:] - box[0, 0])) - 8.0) < 1e-9
    assert merge_box(q[:, [0, 2, 1]], t) is None          # inward-facing faces: a room, not a solid
    assert merge_box(q[:5], t[:5]) is None                # open


def test_histogram_path_is_a_host_side_attribute_not_a_plugin_property(mitr, monkeypatch):
    """TrRenderParams.histogram_path: "auto" unless the integrator attribute (or MITRANSIENT_B200_HISTOGRAM_PATH, which the
    GPU tests use to run every test on both device paths) says otherwise; never read from the plugin dictionary."""
    from mitransient_b200 import _abi
    monkeypatch.delenv("MITRANSIENT_B200_HISTOGRAM_PATH", raising=False)
    integ = mitr.load_dict(mitr.cornell_box_config(4, 4, 8, 1)).integrator()
    assert integ.histogram_path == "auto" and integ.make_params(1, 1, 0).histogram_path == _abi.TR_HIST_AUTO
    for name, value in (("shared", _abi.TR_HIST_SHARED), ("global", _abi.TR_HIST_GLOBAL)):
        integ.histogram_path = name
        assert integ.make_params(1, 1, 0).histogram_path == value
    integ.histogram_path = "registers"
    with pytest.raises(ValueError):
        integ.make_params(1, 1, 0)
    monkeypatch.setenv("MITRANSIENT_B200_HISTOGRAM_PATH", "global")
    assert mitr.load_dict(mitr.cornell_box_config(4, 4, 8, 1)).integrator().histogram_path == "global"
    assert C.sizeof(_abi.TrRenderInfo) == 16
