"""GPU parity of the conductor / dielectric BSDF branches of transient_path (general-material kernel instantiations)
against the CPU oracle: equal integer work counters, film relative L1 <= 1e-5."""
import numpy as np
import pytest

from parity_utils import film_l1

pytestmark = pytest.mark.gpu
WHOLE_TOL, PER_BIN_TOL = 1e-5, 1e-4


def _both(mitr, oracle, d, spp, seed=0):
    scene = mitr.load_dict(d)
    integ = scene.integrator()
    integ.collect_counters = True
    steady, transient = integ.render(scene, seed=seed, spp=spp)
    counted = dict(integ.last_counters)
    raw = scene.sensors()[0].film().develop(raw=True)[1].cpu().numpy()
    ref = oracle.render(scene, integ, seed=seed, spp=spp)
    assert counted == ref["counters"], (counted, ref["counters"])
    err = film_l1(raw, ref["transient"])
    assert err["zero_bins_mismatch"] == 0 and err["whole"] <= WHOLE_TOL and err["max_t"] <= PER_BIN_TOL, err
    np.testing.assert_allclose(steady.cpu().numpy(), ref["steady"], rtol=1e-4, atol=1e-6)
    # the plain (counter-free, 768-thread) instantiation renders the same film
    integ.collect_counters = False
    integ.render(scene, seed=seed, spp=spp)
    raw2 = scene.sensors()[0].film().develop(raw=True)[1].cpu().numpy()
    assert film_l1(raw2, ref["transient"])["whole"] <= WHOLE_TOL
    return raw, ref


def test_cornell_mirror_and_glass(mitr, oracle):
    d = mitr.cornell_box_config(48, 40, 384, 16, max_depth=10, window=(3.5, 12.5))
    d["large-box"]["bsdf"] = {"type": "conductor", "material": "none"}
    d["small-box"]["bsdf"] = {"type": "dielectric", "int_ior": "bk7", "ext_ior": "air"}
    raw, ref = _both(mitr, oracle, d, 16, seed=2)
    assert ref["counters"]["deposits_in_range"] > 0


def test_coloured_conductor_twosided_and_mesh_glass(mitr, oracle):
    d = mitr.cornell_box_config(40, 40, 256, 8, max_depth=12, window=(3.5, 14.0))
    # gold-like complex IOR per channel, two-sided, on the back wall; glass cube built from quads (mesh path)
    d["back"]["bsdf"] = {"type": "twosided", "bsdf": {"type": "conductor", "eta": {"type": "rgb", "value": [0.143, 0.375, 1.442]},
                                                      "k": {"type": "rgb", "value": [3.983, 2.386, 1.603]},
                                                      "specular_reflectance": 0.9}}
    d["small-box"]["bsdf"] = {"type": "dielectric", "int_ior": 1.33, "ext_ior": 1.0,
                              "specular_transmittance": {"type": "rgb", "value": [0.9, 1.0, 0.8]}}
    d["small-box"]["box_primitive"] = False
    _both(mitr, oracle, d, 8, seed=7)


def test_delta_bsdfs_under_transient_prbvolpath(mitr, oracle):
    """conductor / dielectric BSDFs next to the media (transient_prbvolpath.py:364-400 is BSDF-generic): the
    general-material instantiations of the volumetric kernel walk the oracle's paths."""
    d = mitr.cornell_box_volumetric(24, 20, 192, 8)
    d["floor"]["bsdf"] = {"type": "conductor"}
    _both(mitr, oracle, d, 8, seed=3)
    with pytest.raises(ValueError, match="twosided"):
        mitr.load_dict({"type": "scene", "s": {"type": "rectangle", "bsdf": {"type": "twosided", "bsdf": {"type": "dielectric"}}}})
