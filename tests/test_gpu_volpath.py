"""GPU parity of transient_prbvolpath (homogeneous media, HG phase, null boundaries) against the CPU oracle."""
import numpy as np
import pytest

from parity_utils import film_l1

pytestmark = pytest.mark.gpu
WHOLE_TOL, PER_BIN_TOL = 1e-5, 1e-4


@pytest.mark.parametrize("w,h,t,spp,seed,integ", [
    (32, 48, 128, 16, 0, {"max_depth": 16}),
    (24, 24, 96, 32, 1, {"max_depth": 64, "rr_depth": 3}),
    (24, 16, 64, 16, 2, {"max_depth": 8, "camera_unwarp": True}),
    (16, 16, 64, 32, 3, {"max_depth": 12, "hide_emitters": True}),
    (16, 16, 64, 16, 4, {"max_depth": 2}),
])
def test_volumetric_cornell_parity(mitr, oracle, w, h, t, spp, seed, integ):
    d = mitr.cornell_box_volumetric(w, h, t, spp, max_depth=integ.get("max_depth", 16))
    d["integrator"].update(integ)
    scene = mitr.load_dict(d)
    integrator = scene.integrator()
    integrator.collect_counters = True
    steady, transient = integrator.render(scene, seed=seed, spp=spp)
    raw = scene.sensors()[0].film().develop(raw=True)[1].cpu().numpy()
    ref = oracle.render(scene, integrator, seed=seed, spp=spp)
    assert integrator.last_counters == ref["counters"], (integrator.last_counters, ref["counters"])
    err = film_l1(raw, ref["transient"])
    assert err["zero_bins_mismatch"] == 0 and err["whole"] <= WHOLE_TOL and err["max_t"] <= PER_BIN_TOL, err
    np.testing.assert_allclose(steady.cpu().numpy(), ref["steady"], rtol=1e-4, atol=1e-6)


def test_volpath_equals_path_without_media(mitr, oracle):
    """On a scene without media the volumetric integrator is an ordinary path tracer: same expectation as
    transient_path (different estimator details -- RR rule, RNG use -- so only statistically equal)."""
    d = mitr.cornell_box_config(24, 24, 64, 256, max_depth=6)
    s_path = mitr.load_dict(d)
    st_path, _ = s_path.integrator().render(s_path, seed=0, spp=256)
    d2 = mitr.cornell_box_config(24, 24, 64, 256, max_depth=6)
    d2["integrator"] = {"type": "transient_prbvolpath", "max_depth": 6}
    s_vol = mitr.load_dict(d2)
    st_vol, _ = s_vol.integrator().render(s_vol, seed=0, spp=256)
    a, b = st_path.cpu().numpy().mean(), st_vol.cpu().numpy().mean()
    assert abs(a - b) / a < 0.03, (a, b)
