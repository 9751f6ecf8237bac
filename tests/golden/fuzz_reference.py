"""Random sweeps of integrator options, BSDFs, media and occluders: the oracle's deposits against the REFERENCE'S OWN
render() + sample(), executed over refshim (build container only: needs /root/reference).  Test infrastructure.

    python tests/golden/fuzz_reference.py nlos|nlos_scene|nlos_sensor|path|vol|angular|phasor  N  FIRST_SEED
"""
import os
import random
import sys
import traceback
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, HERE)
import numpy as np  # noqa: E402

import refshim  # noqa: E402
import mitransient_b200 as mitr  # noqa: E402
from oracle import oracle  # noqa: E402

T = mitr.ScalarTransform4f
SPP_SCALE = int(os.environ.get("FUZZ_SPP_SCALE", "1"))       # more samples per configuration: rarer events per scene


def sortrows(a): return a[np.lexsort(a.T[::-1])]

def cmp(tag, cfgdesc, dep, r):
    a, b = sortrows(dep), sortrows(r["deposits"])
    eq = a.shape == b.shape and np.array_equal(a, b)
    print(("OK  " if eq else "FAIL"), tag, len(a), len(b), cfgdesc, flush=True)
    return eq

def fuzz_nlos(rng, i):
    kw = {}
    cap = rng.choice(["confocal", "single", "exhaustive"])
    kw["capture_type"] = cap
    kw["nlos_laser_sampling"] = rng.random() < 0.6
    kw["nlos_hidden_geometry_sampling"] = rng.random() < 0.6
    if kw["nlos_hidden_geometry_sampling"]:
        kw["nlos_hidden_geometry_sampling_do_rroulette"] = rng.random() < 0.5
        kw["nlos_hidden_geometry_sampling_includes_relay_wall"] = rng.random() < 0.5
    kw["account_first_and_last_bounces"] = rng.random() < 0.5
    kw["discard_direct_paths"] = rng.random() < 0.3
    if rng.random() < 0.3: kw["filter_depth"] = rng.choice([2, 3, 4])
    kw["max_depth"] = rng.choice([3, 4, 6, 8, -1]) if "filter_depth" not in kw else rng.choice([5, 6, 8])
    kw["rr_depth"] = rng.choice([1, 2, 5])
    w, h = rng.choice([(3, 3), (4, 3), (5, 4)])
    spp = rng.choice([4, 8, 16]) * SPP_SCALE
    if cap == "exhaustive": spp = 4
    d = mitr.nlos_z_scene(w, h, 48, spp, start_opl=0.0, window=16.0, **kw)
    if not kw["nlos_laser_sampling"]:
        d["emitter"]["fov"] = rng.choice([60.0, 120.0])
    if rng.random() < 0.4:
        d["occluder"] = {"type": "rectangle", "to_world": T().translate([-0.25, 0.1, 0.5]).scale([0.3, 0.3, 1.0]),
                         "bsdf": {"type": "twosided", "bsdf": {"type": "diffuse", "reflectance": 0.7}}}
    if cap == "exhaustive":
        f = d["relay_wall"]["nlos_sensor"]["transient_film"]
        f["exhaustive_scan"], f["laser_scan_width"], f["laser_scan_height"] = True, w, h
        if rng.random() < 0.4:       # its own laser grid, scanned through a widened copy of the projector (:349-379)
            f["laser_scan_width"], f["laser_scan_height"] = rng.choice([(2, 2), (3, 2), (2, 4)])
            d["integrator"].update(force_equal_illumination_scanning=False, illumination_scan_fov=rng.choice([10.0, 30.0, 50.0]))
            kw["laser_grid"] = (f["laser_scan_width"], f["laser_scan_height"], d["integrator"]["illumination_scan_fov"])
    scene = mitr.load_dict(d)
    if cap == "single":
        mitr.nlos.focus_emitter_at_relay_wall_pixel((rng.randrange(w), rng.randrange(h)), mitr.PluginSpec(d["relay_wall"]), scene.emitters()[0]["spec"])
    props = {k: v for k, v in d["integrator"].items() if k != "type"}
    seed = rng.randrange(1000)
    steady, raw, dep = refshim.reference_render_nlos(scene, props, spp=spp, seed=seed)
    r = oracle.render(scene, scene.integrator(), spp=spp, seed=seed, log_deposits=1 << 21)
    ok = cmp("nlos", (kw, w, h, spp, seed, "occ" in "".join(d.keys())), dep, r)
    if cap == "exhaustive" and ok:
        ok = np.array_equal(raw != 0, r["transient"] != 0) and np.abs(raw - r["transient"]).max() <= 1e-6 * max(1, np.abs(raw).max())
        if not ok: print("FAIL exhaustive film", kw)
    return ok

def fuzz_nlos_scene(rng, i):
    """Scene variety under transient_nlos_path: several hidden shapes of different kinds and areas (the hidden-geometry
    area distribution), a scaled / shifted relay wall, two-sided and coloured BSDFs, sensor and laser at different places."""
    kw = {"capture_type": rng.choice(["confocal", "single"]), "nlos_laser_sampling": rng.random() < 0.7,
          "nlos_hidden_geometry_sampling": rng.random() < 0.8, "account_first_and_last_bounces": rng.random() < 0.5,
          "max_depth": rng.choice([4, 6, -1]), "rr_depth": rng.choice([2, 5])}
    if kw["nlos_hidden_geometry_sampling"]:
        kw["nlos_hidden_geometry_sampling_do_rroulette"] = rng.random() < 0.5
        kw["nlos_hidden_geometry_sampling_includes_relay_wall"] = rng.random() < 0.3
    w, h = rng.choice([(3, 3), (4, 3), (5, 4)])
    spp = rng.choice([8, 16]) * SPP_SCALE
    d = mitr.nlos_z_scene(w, h, 48, spp, start_opl=0.0, window=24.0, **kw)
    if not kw["nlos_laser_sampling"]:
        d["emitter"]["fov"] = rng.choice([60.0, 120.0])
    colour = lambda: {"type": "rgb", "value": [round(rng.uniform(0.2, 1.0), 3) for _ in range(3)]}
    diffuse = lambda: ({"type": "diffuse", "reflectance": colour()} if rng.random() < 0.5
                       else {"type": "twosided", "bsdf": {"type": "diffuse", "reflectance": round(rng.uniform(0.3, 1.0), 3)}})
    d["geometry"]["bsdf"] = diffuse()
    if rng.random() < 0.5:
        d["relay_wall"]["bsdf"] = {"type": "diffuse", "reflectance": colour()}
    if rng.random() < 0.5:       # a bigger relay wall, shifted sideways
        d["relay_wall"]["to_world"] = T().translate([rng.choice([0.0, 0.2]), 0.0, 0.0]).scale([rng.choice([1.0, 1.5]), rng.choice([1.0, 0.75]), 1.0])
    n_extra = rng.choice([0, 1, 2])
    specular = rng.random() < 0.35
    for k in range(n_extra):
        kind = rng.choice(["rectangle", "cube"])
        pos = [rng.uniform(-0.6, 0.6), rng.uniform(-0.6, 0.6), rng.uniform(0.7, 1.4)]
        sc = rng.uniform(0.1, 0.35)
        tw = T().translate(pos).rotate([0, 1, 0], 180.0 if kind == "rectangle" else rng.uniform(0, 90)).scale([sc, sc * rng.uniform(0.5, 1.5), sc])
        d["hidden%d" % k] = {"type": kind, "to_world": tw, "bsdf": diffuse()}
        if specular and k == 0:      # a mirror / metal / glass object among the hidden geometry: oracle-only, see DESIGN 8.2
            d["hidden0"]["bsdf"] = rng.choice([BSDFS[1], BSDFS[3], BSDFS[8], BSDFS[2]])
    if rng.random() < 0.4:       # sensor somewhere else than the laser
        d["relay_wall"]["nlos_sensor"]["sensor_origin"] = [rng.uniform(-0.8, 0.8), rng.uniform(-0.3, 0.3), rng.uniform(0.2, 0.6)]
    scene = mitr.load_dict(d)
    if kw["capture_type"] == "single":
        mitr.nlos.focus_emitter_at_relay_wall_pixel((rng.randrange(w), rng.randrange(h)), mitr.PluginSpec(d["relay_wall"]), scene.emitters()[0]["spec"])
    props = {k: v for k, v in d["integrator"].items() if k != "type"}
    seed = rng.randrange(1000)
    steady, raw, dep = refshim.reference_render_nlos(scene, props, spp=spp, seed=seed)
    r = oracle.render(scene, scene.integrator(), spp=spp, seed=seed, log_deposits=1 << 21)
    return cmp("nlos_scene", (kw, w, h, spp, seed, n_extra), dep, r)


def fuzz_nlos_sensor(rng, i):
    """The other two ways of looking at the relay wall: the 1x1 "confocal" nlos_capture_meter (original_film_width / height)
    and an ordinary perspective camera driving transient_nlos_path (examples/transient-nlos/nlos-z-simple.xml)."""
    kw = {"nlos_laser_sampling": rng.random() < 0.7, "nlos_hidden_geometry_sampling": rng.random() < 0.7,
          "account_first_and_last_bounces": rng.random() < 0.5, "max_depth": rng.choice([4, 6, -1]), "rr_depth": rng.choice([2, 5])}
    spp = rng.choice([16, 64]) * SPP_SCALE
    if rng.random() < 0.5:
        ow, oh = rng.choice([(4, 4), (8, 6), (5, 7)])
        d = mitr.nlos_z_scene(1, 1, 64, spp, capture_type="single", start_opl=0.0, window=16.0, **kw)
        d["relay_wall"]["nlos_sensor"]["original_film_width"], d["relay_wall"]["nlos_sensor"]["original_film_height"] = ow, oh
        if not kw["nlos_laser_sampling"]: d["emitter"]["fov"] = 120.0
        scene = mitr.load_dict(d)
        mitr.nlos.focus_emitter_at_relay_wall_pixel((rng.randrange(ow), rng.randrange(oh)), mitr.PluginSpec(d["relay_wall"]), scene.emitters()[0]["spec"])
        tag = ("confocal1x1", ow, oh)
    else:
        w, h = rng.choice([(4, 3), (6, 5)])
        capture = rng.choice(["single", "confocal"])
        v, f = mitr.letter_z_mesh()
        cam = T().look_at(origin=[rng.uniform(-1.5, -0.8), rng.uniform(-0.3, 0.3), rng.uniform(1.2, 1.9)], target=[0.0, 0.0, 0.0], up=[0, 1, 0])
        d = {"type": "scene",
             "integrator": {"type": "transient_nlos_path", "capture_type": capture, **kw},
             "sensor": {"type": "perspective", "fov": rng.choice([30.0, 40.0]), "near_clip": 0.1, "far_clip": 100.0, "to_world": cam,
                        "sampler": {"type": "independent", "sample_count": spp},
                        "film": {"type": "transient_hdr_film", "width": w, "height": h, "temporal_bins": 64, "start_opl": 0.0,
                                 "bin_width_opl": 16.0 / 64, "rfilter": {"type": "box"}}},
             "emitter": {"type": "projector", "irradiance": 100.0, "fov": 0.2 if kw["nlos_laser_sampling"] else 60.0, "to_world": cam},
             "relay_wall": {"type": "rectangle", "to_world": T().scale([2.0, 2.0, 1.0]),
                            "bsdf": {"type": "twosided", "bsdf": {"type": "diffuse", "reflectance": 0.8}}},
             "hidden": {"type": "trimesh", "vertices": v, "faces": f, "to_world": T().translate([0.0, 0.0, 1.0]),
                        "bsdf": {"type": "twosided", "bsdf": {"type": "diffuse", "reflectance": 1.0}}}}
        scene = mitr.load_dict(d)
        tag = ("perspective", capture, w, h)
    props = {k: v for k, v in d["integrator"].items() if k != "type"}
    seed = rng.randrange(1000)
    steady, raw, dep = refshim.reference_render_nlos(scene, props, spp=spp, seed=seed)
    r = oracle.render(scene, scene.integrator(), spp=spp, seed=seed, log_deposits=1 << 21)
    return cmp("nlos_sensor", (tag, kw, spp, seed), dep, r)


GOLD = {"type": "conductor", "eta": {"type": "rgb", "value": [0.143, 0.375, 1.442]}, "k": {"type": "rgb", "value": [3.983, 2.386, 1.603]}}
BSDFS = [None, {"type": "conductor", "material": "none"}, {"type": "dielectric", "int_ior": 1.5, "ext_ior": 1.0},
         {"type": "twosided", "bsdf": GOLD}, {"type": "null"}, GOLD, {"type": "dielectric", "int_ior": "water", "ext_ior": "air"},
         {"type": "dielectric", "int_ior": 1.0, "ext_ior": 1.5},
         {"type": "conductor", "eta": 0.2, "k": 3.9, "specular_reflectance": {"type": "rgb", "value": [0.9, 0.5, 0.3]}},
         {"type": "diffuse", "reflectance": {"type": "rgb", "value": [0.2, 0.9, 0.4]}},
         {"type": "twosided", "bsdf": {"type": "diffuse", "reflectance": 0.3}}]

def fuzz_path(rng, i):
    integ = {"max_depth": rng.choice([2, 3, 5, 8, 12, -1]), "rr_depth": rng.choice([1, 2, 3, 5]),
             "camera_unwarp": rng.random() < 0.4, "discard_direct_light": rng.random() < 0.3}
    w, h = rng.choice([(4, 4), (6, 5), (5, 7)])
    spp = rng.choice([4, 8]) * SPP_SCALE
    d = mitr.cornell_box_config(w, h, 32, spp, max_depth=8)
    d["integrator"].update(integ)
    bs = {}
    for shape in ("small-box", "large-box", "back", "floor"):
        b = rng.choice(BSDFS if shape.endswith("box") else [None, None] + BSDFS[:4] + BSDFS[5:])
        if b is not None: d[shape]["bsdf"] = b; bs[shape] = b["type"] + str(b.get("material", b.get("int_ior", "")))
    if rng.random() < 0.25:              # a crop window of the film
        cw, ch = rng.randrange(1, w + 1), rng.randrange(1, h + 1)
        d["sensor"]["film"].update(crop_width=cw, crop_height=ch, crop_offset_x=rng.randrange(0, w - cw + 1), crop_offset_y=rng.randrange(0, h - ch + 1))
        bs["crop"] = (cw, ch)
    extra = rng.random() < 0.35          # a second (and third) area light: uniform emitter selection with sample reuse
    if extra:
        d["small-box"]["emitter"] = {"type": "area", "radiance": {"type": "rgb", "value": [2.0, 5.0, 1.0]}}
        if rng.random() < 0.5:
            d["lamp2"] = {"type": "rectangle", "to_world": T().translate([0.5, -0.2, 0.3]).rotate([0, 1, 0], -90.0).scale([0.2, 0.2, 1.0]),
                          "bsdf": {"type": "diffuse", "reflectance": 0.5},
                          "emitter": {"type": "area", "radiance": {"type": "rgb", "value": [9.0, 3.0, 4.0]}}}
        bs["emitters"] = 2 + ("lamp2" in d)
    scene = mitr.load_dict(d)
    props = {k: v for k, v in d["integrator"].items() if k != "type"}
    seed = rng.randrange(1000)
    steady, raw, dep = refshim.reference_render(scene, "mitransient.integrators.transientpath", "TransientPath", props, spp=spp, seed=seed)
    r = oracle.render(scene, scene.integrator(), spp=spp, seed=seed, log_deposits=1 << 21)
    ok = cmp("path", (integ, bs, w, h, spp, seed), dep, r)
    if ok and not np.allclose(steady, r["steady"], rtol=3e-7, atol=0): print("FAIL steady", np.abs(steady - r["steady"]).max()); ok = False
    if ok and not (raw.shape == r["transient"].shape and np.abs(raw - r["transient"]).max() <= 1e-6 * max(1.0, np.abs(raw).max())):
        print("FAIL film", raw.shape, r["transient"].shape); ok = False
    return ok

def fuzz_vol(rng, i):
    integ = {"max_depth": rng.choice([2, 3, 6, 12, -1]), "rr_depth": rng.choice([1, 2, 5]),
             "camera_unwarp": rng.random() < 0.4, "hide_emitters": rng.random() < 0.3}
    w, h = rng.choice([(4, 4), (6, 5), (5, 7)])
    spp = rng.choice([4, 8]) * SPP_SCALE
    media = rng.random() < 0.8
    d = mitr.cornell_box_volumetric(w, h, 32, spp, max_depth=8) if media else mitr.cornell_box_config(w, h, 32, spp, max_depth=8)
    if not media: d["integrator"] = {"type": "transient_prbvolpath", "temporal_filter": "box"}
    d["integrator"].update(integ)
    if media:
        for shape in ("small-box", "large-box"):
            d[shape]["interior"]["sigma_t"] = rng.choice([0.5, 2.0, 5.0, 14.0])
            d[shape]["interior"]["phase"]["g"] = rng.choice([-0.5, 0.0, 0.1, 0.7])
    if rng.random() < 0.4:           # a second light low in the room: medium vertices see lights from the front and from behind
        d["lamp2"] = {"type": "rectangle", "to_world": T().translate([rng.choice([0.5, -0.2]), -0.3, 0.3]).rotate([0, 1, 0], rng.choice([-90.0, 90.0, 0.0])).scale([0.2, 0.2, 1.0]),
                      "bsdf": {"type": "diffuse", "reflectance": 0.5},
                      "emitter": {"type": "area", "radiance": {"type": "rgb", "value": [9.0, 3.0, 4.0]}}}
    spec = {}
    if rng.random() < 0.4:           # specular walls / a glass or metal box next to (or around) the media: oracle-only, see DESIGN 8.2
        for shape in rng.sample(["back", "floor", "small-box", "large-box"], rng.choice([1, 2])):
            b = rng.choice([BSDFS[1], BSDFS[2], BSDFS[3], BSDFS[5], BSDFS[6], BSDFS[8]])
            d[shape]["bsdf"] = b; spec[shape] = b["type"]
    scene = mitr.load_dict(d)
    props = {k: v for k, v in d["integrator"].items() if k != "type"}
    seed = rng.randrange(1000)
    steady, raw, dep = refshim.reference_render_vol(scene, props, spp=spp, seed=seed)
    r = oracle.render(scene, scene.integrator(), spp=spp, seed=seed, log_deposits=1 << 21)
    ok = cmp("vol", (integ, media, spec, "lamp2" in d, w, h, spp, seed), dep, r)
    if ok and not np.allclose(steady, r["steady"], rtol=3e-7, atol=0): print("FAIL steady", np.abs(steady - r["steady"]).max()); ok = False
    return ok

def fuzz_angular(rng, i):
    """mitransient's `angulararea` emitter (emitters/angulararea.py) with random cutoff / beam angles inside transient_path."""
    integ = {"max_depth": rng.choice([2, 3, 5, 8, -1]), "rr_depth": rng.choice([1, 2, 5]), "camera_unwarp": rng.random() < 0.3,
             "discard_direct_light": rng.random() < 0.2}
    w, h = rng.choice([(4, 4), (6, 5), (5, 7)])
    spp = rng.choice([8, 16]) * SPP_SCALE
    cutoff = rng.choice([10.0, 25.0, 45.0, 60.0, 80.0, 89.0])
    beam = cutoff if rng.random() < 0.25 else round(rng.uniform(0.0, cutoff), 2)
    d = mitr.cornell_box_config(w, h, 32, spp, max_depth=8, window=(3.5, 12.5))
    d["integrator"].update(integ)
    d["light"]["emitter"] = {"type": "angulararea", "radiance": {"type": "rgb", "value": [18.387, 10.9873, 2.75357]},
                             "cutoff_angle": cutoff, "beam_width": beam}
    if rng.random() < 0.4:
        d["large-box"]["bsdf"] = rng.choice(BSDFS[1:])
    scene = mitr.load_dict(d)
    props = {k: v for k, v in d["integrator"].items() if k != "type"}
    seed = rng.randrange(1000)
    steady, raw, dep = refshim.reference_render_angular(scene, props, spp=spp, seed=seed)
    r = oracle.render(scene, scene.integrator(), spp=spp, seed=seed, log_deposits=1 << 21)
    ok = cmp("angular", (integ, cutoff, beam, w, h, spp, seed), dep, r)
    if ok and not np.allclose(steady, r["steady"], rtol=3e-7, atol=0): print("FAIL steady", np.abs(steady - r["steady"]).max()); ok = False
    return ok


def fuzz_phasor(rng, i):
    """The reference's PhasorHDRFilm (frequency selection + per-deposit cos / sin) filled by its transient_path loop."""
    integ = {"max_depth": rng.choice([3, 5, 8]), "camera_unwarp": rng.random() < 0.3}
    w, h = rng.choice([(3, 3), (4, 3), (5, 4)])
    spp = rng.choice([4, 8]) * SPP_SCALE
    bins = rng.choice([64, 256, 1000])
    wl_mean, wl_sigma = rng.choice([0.1, 0.2, 0.5, 1.0]), rng.choice([0.05, 0.1, 0.5])
    mitr.set_variant("llvm_ad_mono")
    try:
        d = mitr.cornell_box_config(w, h, bins, spp)
        d["integrator"].update(integ)
        d["sensor"]["film"] = {"type": "phasor_hdr_film", "width": w, "height": h, "temporal_bins": bins, "start_opl": 3.5,
                               "bin_width_opl": 6.0 / bins, "wl_mean": wl_mean, "wl_sigma": wl_sigma}
        scene = mitr.load_dict(d)
        props = {k: v for k, v in d["integrator"].items() if k != "type"}
        seed = rng.randrange(1000)
        steady, raw, dep = refshim.reference_render(scene, "mitransient.integrators.transientpath", "TransientPath", props, spp=spp,
                                                    seed=seed, monochromatic=True)
        r = oracle.render(scene, scene.integrator(), spp=spp, seed=seed)
    finally:
        mitr.set_variant("cuda_rgb")
    ok = raw.shape == r["transient"].shape and np.abs(raw - r["transient"]).max() <= 2e-5 * max(1.0, np.abs(raw).max())
    print(("OK  " if ok else "FAIL"), "phasor", raw.shape, (integ, w, h, spp, bins, wl_mean, wl_sigma, seed),
          float(np.abs(raw - r["transient"]).max()) if raw.shape == r["transient"].shape else None, flush=True)
    return ok


FUZZERS = {"nlos": fuzz_nlos, "nlos_scene": fuzz_nlos_scene, "nlos_sensor": fuzz_nlos_sensor, "path": fuzz_path, "vol": fuzz_vol, "angular": fuzz_angular,
           "phasor": fuzz_phasor}


def run_one(which, seed):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return FUZZERS[which](random.Random(seed), seed)


if __name__ == "__main__":
    warnings.simplefilter("ignore")
    which, n, s0 = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    fn = FUZZERS[which]
    bad = 0
    for i in range(n):
        rng = random.Random(s0 + i)
        try:
            if not fn(rng, i): bad += 1; print("   ^ seed", s0 + i)
        except Exception as e:
            print("EXC seed", s0 + i, type(e).__name__, str(e)[:300]); traceback.print_exc(limit=3)
    print("bad:", bad, "of", n)
