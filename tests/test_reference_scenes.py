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


@pytest.mark.parametrize("rel,needle", sorted(UNSUPPORTED.items()))
def test_unsupported_example_scene_fails_loudly(mitr, rel, needle):
    with pytest.raises((ValueError, NotImplementedError)) as e:
        mitr.load_file(os.path.join(EXAMPLES, rel))
    assert needle in str(e.value)
