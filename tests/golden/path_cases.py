"""The transient_path configurations whose deposits were produced by executing the reference's render() + sample()
(make_reference_golden.py, section H) -- shared by the generator and by tests/test_reference_golden.py."""
import numpy as np

import mitransient_b200 as mitr

PATH_CASES = {
    "path_a": ((8, 6, 64, 4), {"max_depth": 8}, {}),
    "path_b": ((6, 6, 32, 8), {"max_depth": -1, "rr_depth": 2}, {}),
    "path_c": ((6, 4, 32, 8), {"max_depth": 5, "camera_unwarp": True}, {}),
    "path_d": ((5, 5, 16, 8), {"max_depth": 6, "discard_direct_light": True}, {}),
    "path_e": ((8, 8, 48, 8), {"max_depth": 10}, {"large-box": {"type": "conductor", "material": "none"},
                                                   "small-box": {"type": "dielectric", "int_ior": "bk7", "ext_ior": "air"}}),
    "path_f": ((6, 6, 32, 6), {"max_depth": 12, "rr_depth": 3},
               {"back": {"type": "twosided", "bsdf": {"type": "conductor", "eta": {"type": "rgb", "value": [0.143, 0.375, 1.442]},
                                                       "k": {"type": "rgb", "value": [3.983, 2.386, 1.603]}}},
                "small-box": {"type": "dielectric", "int_ior": 1.33, "ext_ior": 1.0}}),
}


def path_case_scene(name):
    cfg, integ, bsdfs = PATH_CASES[name]
    d = mitr.cornell_box_config(*cfg, window=(3.5, 12.5))
    d["integrator"].update(integ)
    for shape, b in bsdfs.items():
        d[shape]["bsdf"] = b
    return mitr.load_dict(d), cfg, {k: v for k, v in d["integrator"].items() if k != "type"}


def sort_rows(a):
    return a[np.lexsort(a.T[::-1])]


