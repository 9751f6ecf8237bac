"""The reference's own example scene files, read where they lie (/root/reference exists only in the build
container; skipped elsewhere): every file of the supported plugin subset must load through ``load_file`` and render
through the CPU oracle -- the "on-disk format either side of the path" row of SURVEY 8f."""
import glob
import os

import numpy as np
import pytest

EXAMPLES = "/root/reference/examples"
pytestmark = pytest.mark.skipif(not os.path.isdir(EXAMPLES), reason="the reference tree is not available here")

SUPPORTED = ["transient/cornell-box/cbox_diffuse.xml", "transient/cornell-box/cbox_mirror.xml",
             "transient/cornell-box/cbox_volumetric.xml", "transient/cornell-box/cbox_diffuse_freq.xml",
             "diff-transient/cornell-box/cbox_volumetric.xml", "transient-nlos/nlos_Z.xml",
             "transient-nlos/nlos-z-simple.xml", "transient-nlos/nlos-z-room.xml",
             "angulararea-emitter/scenes/area_1light.xml", "angulararea-emitter/scenes/angular_1light.xml"]
# what the others need (documented in DESIGN.md, "Scope")
UNSUPPORTED = {"polarization/cornell-box/cbox_polarized.xml": "roughconductor",
               "polarization/cornell-box/cbox_polarized_steady.xml": "stokes",
               "diff-transient/staircase/scene.xml": ""}


def test_every_example_file_is_classified():
    found = {os.path.relpath(p, EXAMPLES) for p in glob.glob(EXAMPLES + "/**/*.xml", recursive=True)}
    assert found == set(SUPPORTED) | set(UNSUPPORTED)


@pytest.mark.parametrize("rel", SUPPORTED)
def test_example_scene_loads_and_renders_with_the_oracle(mitr, oracle, rel):
    import warnings
    phasor = rel.endswith("_freq.xml")
    mitr.set_variant("llvm_ad_mono" if phasor else "llvm_ad_rgb")
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            scene = mitr.load_file(os.path.join(EXAMPLES, rel), res=6, spp=2)
        sensor = scene.sensors()[0]
        film = sensor.film()
        w, h = film.size()
        assert w * h <= 64 * 64, "the res default was not substituted"
        integ = scene.integrator()
        integ.check_transient_(scene, sensor)
        integ.check_scene_(scene, sensor)
        r = oracle.render(scene, integ, spp=2)
        assert r["counters"]["paths"] == w * h * 2
        assert np.isfinite(r["transient"]).all() and np.isfinite(r["steady"]).all()
        if phasor:
            assert r["transient"].shape == (h, w, 2 * len(film.frequencies) + 1)
        else:
            assert r["transient"].shape == (h, w, film.temporal_bins, 4)
    finally:
        mitr.set_variant("cuda_rgb")


FIXTURES = {"cbox_diffuse": ("transient/cornell-box/cbox_diffuse.xml", {"res": 64, "spp": 16}),
            "cbox_volumetric": ("transient/cornell-box/cbox_volumetric.xml", {"res": 64, "spp": 16}),
            "nlos_Z": ("transient-nlos/nlos_Z.xml", {})}


@pytest.mark.parametrize("name", sorted(FIXTURES))
def test_scene_fixture_equals_the_reference_file(mitr, name):
    """tests/golden/scenes/<name>.xml (written by make_scene_fixtures.py with OUR exporter, for the GPU box) builds the
    same flat scene descriptor as the reference's own file: primitives, shapes, BSDFs, emitters, media, sensor, film,
    integrator."""
    import ctypes as C
    import warnings
    rel, params = FIXTURES[name]
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "scenes", name + ".xml")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a, b = mitr.load_file(os.path.join(EXAMPLES, rel), **params), mitr.load_file(here)
    da, db = a.describe(a.sensors()[0]), b.describe(b.sensors()[0])
    assert da.n_prims == db.n_prims and da.n_shapes == db.n_shapes and da.n_bsdfs == db.n_bsdfs
    assert da.n_emitters == db.n_emitters and da.n_media == db.n_media
    va = np.ctypeslib.as_array(C.cast(da.prim_verts, C.POINTER(C.c_float)), shape=(da.n_prims, 12))
    vb = np.ctypeslib.as_array(C.cast(db.prim_verts, C.POINTER(C.c_float)), shape=(db.n_prims, 12))
    assert np.array_equal(va, vb)
    assert bytes(da.sensor) == bytes(db.sensor) and bytes(da.film) == bytes(db.film)
    for i in range(da.n_bsdfs):
        assert bytes(da.bsdfs[i]) == bytes(db.bsdfs[i])
    for i in range(da.n_emitters):
        assert bytes(da.emitters[i]) == bytes(db.emitters[i])
    for i in range(da.n_media):
        assert bytes(da.media[i]) == bytes(db.media[i])
    assert a.integrator().__dict__ == b.integrator().__dict__


@pytest.mark.parametrize("rel,needle", sorted(UNSUPPORTED.items()))
def test_unsupported_example_scene_fails_loudly(mitr, rel, needle):
    with pytest.raises((ValueError, NotImplementedError)) as e:
        mitr.load_file(os.path.join(EXAMPLES, rel))
    assert needle in str(e.value)
