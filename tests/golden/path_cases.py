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
    "path_g": ((6, 6, 48, 8), {"max_depth": 8}, "volumetric"),
    # a crop window (crop_width, crop_height, crop_offset_x, crop_offset_y) of a 12 x 10 film: sample_rays shifts the pixels,
    # the block subtracts the offset again and keeps the film's size (transient_hdr_film.py:192-195)
    "path_h": ((12, 10, 32, 8), {"max_depth": 6, "crop": (5, 4, 6, 3)}, {}),
}


def path_case_scene(name):
    cfg, integ, bsdfs = PATH_CASES[name]
    if bsdfs == "volumetric":       # cbox_volumetric under transient_path: null boundaries are walked through as delta lobes
        d = mitr.cornell_box_volumetric(*cfg, max_depth=8)
        d["integrator"]["type"] = "transient_path"
        d["integrator"].update(integ)
        return mitr.load_dict(d), cfg, {k: v for k, v in d["integrator"].items() if k != "type"}
    d = mitr.cornell_box_config(*cfg, window=(3.5, 12.5))
    integ = dict(integ)
    crop = integ.pop("crop", None)
    if crop is not None:
        d["sensor"]["film"].update(crop_width=crop[0], crop_height=crop[1], crop_offset_x=crop[2], crop_offset_y=crop[3])
    d["integrator"].update(integ)
    for shape, b in bsdfs.items():
        d[shape]["bsdf"] = b
    return mitr.load_dict(d), cfg, {k: v for k, v in d["integrator"].items() if k != "type"}


def sort_rows(a):
    return a[np.lexsort(a.T[::-1])]




# ---- transient_nlos_path + nlos_capture_meter: (film (w, h, bins, spp), nlos_z_scene keyword arguments, laser focus pixel,
#      projector fov override, exhaustive_scan film)
NLOS_CASES = {
    "nlos_a": ((6, 6, 64, 8), {}, None, None, False),
    "nlos_b": ((5, 5, 64, 8), {"nlos_hidden_geometry_sampling_do_rroulette": True}, None, None, False),
    "nlos_c": ((5, 5, 64, 8), {"nlos_hidden_geometry_sampling": False, "account_first_and_last_bounces": True, "start_opl": 0.0,
                               "window": 8.0}, None, None, False),
    "nlos_d": ((5, 5, 64, 8), {"filter_depth": 3, "start_opl": 0.0, "window": 8.0}, None, None, False),
    "nlos_e": ((5, 5, 64, 8), {"discard_direct_paths": True, "nlos_hidden_geometry_sampling_includes_relay_wall": True, "max_depth": 7,
                               "rr_depth": 2, "start_opl": 0.0, "window": 8.0}, None, None, False),
    "nlos_f": ((6, 6, 64, 16), {"capture_type": "single", "start_opl": 0.0, "window": 8.0}, (3, 2), None, False),
    "nlos_g": ((6, 6, 64, 64), {"capture_type": "single", "nlos_laser_sampling": False, "start_opl": 0.0, "window": 8.0}, (3, 3), 60.0, False),
    # plain emitter sampling + account_first_and_last_bounces: the reference's in-place `distance += ds.dist` inside
    # emitter_nee_sample leaks into the loop's own distance -- deposits carry the accumulated lengths
    "nlos_n": ((5, 5, 64, 64), {"capture_type": "single", "nlos_laser_sampling": False, "nlos_hidden_geometry_sampling": False,
                                "account_first_and_last_bounces": True, "start_opl": 0.0, "window": 16.0, "max_depth": 6},
               (2, 2), 120.0, False),
    # ... and with an occluder between part of the scene and the laser: occluded connections deposit nothing but still add
    # ds.dist (the projector fills its DirectionSample regardless of the mask, :497-498 is unmasked)
    "nlos_o": ((5, 5, 64, 64), {"capture_type": "single", "nlos_laser_sampling": False, "nlos_hidden_geometry_sampling": False,
                                "account_first_and_last_bounces": True, "start_opl": 0.0, "window": 16.0, "max_depth": 6,
                                "occluder": True}, (2, 2), 120.0, False),
    # exhaustive capture with laser sampling and an occluder: some laser targets are invisible from some vertices.  The
    # `active_e &= ...` inside emitter_laser_targets_sample mutates the mask it was handed, but that mask is not state of the
    # recorded (symbolic) laser loop, so every laser target starts from the caller's mask (refshim._SymbolicLoopState)
    "nlos_p": ((3, 3, 48, 4), {"capture_type": "exhaustive", "nlos_hidden_geometry_sampling": False, "start_opl": 0.0, "window": 16.0,
                               "max_depth": 6, "rr_depth": 2, "occluder": True}, None, None, True),
    # a mirror, a metal plate and a glass block among the hidden geometry: delta lobes under transient_nlos_path (no emitter
    # sample at those vertices, bsdf.sample continues the path, eta scales the path length; with plain emitter sampling the
    # aliased distance still grows there).  GPU legs NOT YET RUN (branch next/specular-nlos).
    "nlos_q": ((4, 4, 48, 64), {"start_opl": 0.0, "window": 24.0, "max_depth": 6, "plate": {"type": "conductor", "material": "none"}},
               None, None, False),
    "nlos_r": ((4, 4, 48, 64), {"capture_type": "single", "nlos_laser_sampling": False, "nlos_hidden_geometry_sampling": False,
                                "account_first_and_last_bounces": True, "start_opl": 0.0, "window": 24.0, "max_depth": 6,
                                "plate": {"type": "twosided", "bsdf": {"type": "conductor", "eta": {"type": "rgb", "value": [0.143, 0.375, 1.442]},
                                                                       "k": {"type": "rgb", "value": [3.983, 2.386, 1.603]}}}},
               (2, 1), 120.0, False),
    "nlos_s": ((4, 4, 48, 64), {"nlos_hidden_geometry_sampling": False, "nlos_hidden_geometry_sampling_do_rroulette": False, "start_opl": 0.0,
                                "window": 24.0, "max_depth": 8, "rr_depth": 3, "plate": {"type": "dielectric", "int_ior": 1.5, "ext_ior": 1.0}},
               None, None, False),
    "nlos_h": ((3, 3, 32, 4), {"capture_type": "exhaustive", "start_opl": 0.0, "window": 8.0}, None, None, True),
    "nlos_i": ((3, 3, 32, 4), {"capture_type": "exhaustive", "nlos_laser_sampling": False, "start_opl": 0.0, "window": 8.0}, None, 60.0, True),
    # 1x1 "confocal" sensor (original_film_width / height): every ray goes to the focused laser point
    # Exhaustive capture with its own (non-square) laser grid, scanned through a 30-degree copy of the projector
    "nlos_m": ((4, 4, 32, 4), {"capture_type": "exhaustive", "force_equal_illumination_scanning": False, "illumination_scan_fov": 30.0,
                               "laser_grid": (3, 2), "start_opl": 0.0, "window": 8.0}, (2, 2), None, True),
    "nlos_k": ((8, 6, 128, 64), {"perspective": "single"}, None, None, False),
    "nlos_l": ((8, 6, 128, 64), {"perspective": "confocal"}, None, None, False),
    "nlos_j": ((1, 1, 128, 256), {"capture_type": "single", "start_opl": 0.0, "window": 8.0, "original_film": (8, 8)}, (5, 3), None, False),
}


def _perspective_nlos_scene(cfg, capture):
    """transient_nlos_path looking at the relay wall through an ordinary perspective camera (the setup of the reference's
    examples/transient-nlos/nlos-z-simple.xml)."""
    T = mitr.ScalarTransform4f
    v, f = mitr.letter_z_mesh()
    cam = T().look_at(origin=[-1.2, 0.0, 1.6], target=[0.0, 0.0, 0.0], up=[0, 1, 0])
    return {"type": "scene",
            "integrator": {"type": "transient_nlos_path", "nlos_laser_sampling": True, "nlos_hidden_geometry_sampling": True,
                           "capture_type": capture, "account_first_and_last_bounces": capture == "single", "max_depth": 5},
            "sensor": {"type": "perspective", "fov": 40.0, "near_clip": 0.1, "far_clip": 100.0, "to_world": cam,
                       "sampler": {"type": "independent", "sample_count": cfg[3]},
                       "film": {"type": "transient_hdr_film", "width": cfg[0], "height": cfg[1], "temporal_bins": cfg[2],
                                "start_opl": 0.0, "bin_width_opl": 12.0 / cfg[2], "rfilter": {"type": "box"}}},
            "emitter": {"type": "projector", "irradiance": 100.0, "fov": 0.2, "to_world": cam},
            "relay_wall": {"type": "rectangle", "to_world": T().scale([2.0, 2.0, 1.0]),
                           "bsdf": {"type": "twosided", "bsdf": {"type": "diffuse", "reflectance": 0.8}}},
            "hidden": {"type": "trimesh", "vertices": v, "faces": f, "to_world": T().translate([0.0, 0.0, 1.0]),
                       "bsdf": {"type": "twosided", "bsdf": {"type": "diffuse", "reflectance": 1.0}}}}


def nlos_case_scene(name):
    cfg, kw, focus, fov, exhaustive = NLOS_CASES[name]
    if "perspective" in kw:
        d = _perspective_nlos_scene(cfg, kw["perspective"])
        return mitr.load_dict(d), cfg, {k: v for k, v in d["integrator"].items() if k != "type"}
    kw = dict(kw)
    original = kw.pop("original_film", None)
    laser_grid = kw.pop("laser_grid", None)
    occluder = kw.pop("occluder", False)
    plate = kw.pop("plate", None)
    d = mitr.nlos_z_scene(*cfg, **kw)
    if plate is not None:        # a tilted plate (a block for glass) next to the letter
        glass = plate["type"] == "dielectric"
        d["plate"] = {"type": "cube" if glass else "rectangle",
                      "to_world": mitr.ScalarTransform4f().translate([0.5, 0.0, 0.8]).rotate([0, 1, 0], -135.0).scale([0.4, 0.4, 0.1 if glass else 1.0]),
                      "bsdf": plate}
    if occluder:
        d["occluder"] = {"type": "rectangle", "to_world": mitr.ScalarTransform4f().translate([-0.25, 0.1, 0.5]).scale([0.3, 0.3, 1.0]),
                         "bsdf": {"type": "twosided", "bsdf": {"type": "diffuse", "reflectance": 0.7}}}
    if original is not None:
        d["relay_wall"]["nlos_sensor"]["original_film_width"], d["relay_wall"]["nlos_sensor"]["original_film_height"] = original
    if fov is not None:
        d["emitter"]["fov"] = fov
    if exhaustive:
        f = d["relay_wall"]["nlos_sensor"]["transient_film"]
        f["exhaustive_scan"], f["laser_scan_width"], f["laser_scan_height"] = True, cfg[0], cfg[1]
        if laser_grid is not None:
            f["laser_scan_width"], f["laser_scan_height"] = laser_grid
    scene = mitr.load_dict(d)
    if focus is not None:
        wall = [v for v in d.values() if isinstance(v, dict) and v.get("type") == "rectangle"][0]
        mitr.nlos.focus_emitter_at_relay_wall_pixel(focus, mitr.PluginSpec(wall), scene.emitters()[0]["spec"])
    return scene, cfg, {k: v for k, v in d["integrator"].items() if k != "type"}


# ---- mitransient's `angulararea` emitter under transient_path: (film, integrator overrides, cutoff_angle, beam_width in degrees)
ANGULAR_CASES = {
    "angular_a": ((8, 8, 64, 16), {"max_depth": 5}, 60.0, 35.0),
    "angular_b": ((6, 6, 48, 16), {"max_depth": 8, "rr_depth": 2}, 80.0, 80.0),      # cutoff == beam: hard edge, inv_transition = inf
}


def angular_case_scene(name):
    cfg, integ, cutoff, beam = ANGULAR_CASES[name]
    d = mitr.cornell_box_config(*cfg, window=(3.5, 12.5))
    d["integrator"].update(integ)
    d["light"]["emitter"] = {"type": "angulararea", "radiance": {"type": "rgb", "value": [18.387, 10.9873, 2.75357]},
                             "cutoff_angle": cutoff, "beam_width": beam}
    return mitr.load_dict(d), cfg, {k: v for k, v in d["integrator"].items() if k != "type"}


# ---- transient_prbvolpath: (film, integrator overrides, volumetric scene?)
VOL_CASES = {
    "vol_a": ((6, 6, 48, 8), {"max_depth": 6}, True),
    "vol_b": ((6, 6, 48, 8), {"max_depth": 12, "camera_unwarp": True, "rr_depth": 2}, True),
    "vol_c": ((5, 7, 48, 16), {"max_depth": -1, "rr_depth": 1, "hide_emitters": True}, True),
    "vol_d": ((6, 6, 48, 8), {"max_depth": 6}, False),          # no medium at all: the surface-only branch of the same loop
    # a second light standing in the room: medium vertices see lights from behind, whose radiance-free emitter samples still
    # walk to the light and draw per medium segment (transient_prbvolpath.py:429-457)
    "vol_e": ((5, 7, 48, 8), {"max_depth": 6, "rr_depth": 2, "lamp2": True}, True),
    # mirrors, metals and glass next to (and around) the media: delta lobes, eta-scaled path lengths, specular_chain
    "vol_f": ((6, 6, 48, 8), {"max_depth": 8, "rr_depth": 2, "bsdfs": {"back": {"type": "conductor", "material": "none"},
                                                                     "floor": {"type": "dielectric", "int_ior": 1.5, "ext_ior": 1.0}}}, True),
    "vol_g": ((5, 7, 48, 8), {"max_depth": -1, "rr_depth": 3, "camera_unwarp": True,
                              "bsdfs": {"small-box": {"type": "dielectric", "int_ior": 1.33, "ext_ior": 1.0},
                                        "red-wall": {"type": "twosided", "bsdf": {"type": "conductor", "eta": {"type": "rgb", "value": [0.143, 0.375, 1.442]},
                                                                              "k": {"type": "rgb", "value": [3.983, 2.386, 1.603]}}}}}, True),
}
VOL_CASES_GPU = tuple(sorted(VOL_CASES))                      # vol_e, vol_f, vol_g: GPU legs NOT YET RUN (branch next/specular-volpath)


def vol_case_scene(name):
    cfg, integ, volumetric = VOL_CASES[name]
    if volumetric:
        d = mitr.cornell_box_volumetric(*cfg, max_depth=8)
    else:
        d = mitr.cornell_box_config(*cfg, max_depth=8)
        d["integrator"]["type"] = "transient_prbvolpath"
    integ = dict(integ)
    for shape, b in integ.pop("bsdfs", {}).items():
        d[shape]["bsdf"] = b
    if integ.pop("lamp2", False):
        d["lamp2"] = {"type": "rectangle", "to_world": mitr.ScalarTransform4f().translate([0.5, -0.3, 0.3]).rotate([0, 1, 0], -90.0).scale([0.2, 0.2, 1.0]),
                      "bsdf": {"type": "diffuse", "reflectance": 0.5},
                      "emitter": {"type": "area", "radiance": {"type": "rgb", "value": [9.0, 3.0, 4.0]}}}
    d["integrator"].update(integ)
    return mitr.load_dict(d), cfg, {k: v for k, v in d["integrator"].items() if k != "type"}


# ---- phasor_hdr_film filled by the reference's transient_path loop (monochromatic variant)
PHASOR_CASES = {"phasor_path_a": ((6, 5, 256, 8), {"max_depth": 8}, 0.5, 0.5),
                "phasor_path_b": ((4, 4, 1024, 16), {"max_depth": 5, "camera_unwarp": True}, 0.2, 0.1)}


def phasor_case_scene(name):
    """Call with the monochromatic variant active (mitr.set_variant('..._mono'))."""
    cfg, integ, wl_mean, wl_sigma = PHASOR_CASES[name]
    d = mitr.cornell_box_config(*cfg)
    d["integrator"].update(integ)
    d["sensor"]["film"] = {"type": "phasor_hdr_film", "width": cfg[0], "height": cfg[1], "temporal_bins": cfg[2], "start_opl": 3.5,
                           "bin_width_opl": 6.0 / cfg[2], "wl_mean": wl_mean, "wl_sigma": wl_sigma}
    return mitr.load_dict(d), cfg, {k: v for k, v in d["integrator"].items() if k != "type"}
