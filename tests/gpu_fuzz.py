"""Random GPU-vs-oracle sweep (run on a B200 box; test infrastructure, not collected by pytest):

    python tests/gpu_fuzz.py  N  [FIRST_SEED]  [kinds: path,mesh,bvh,mats,vol,nlos]

Every configuration draws a scene (Cornell-box variants with rotated / rescaled boxes, mesh boxes, icospheres behind
the BVH, mirrors / metals / glass, participating media, NLOS setups with random capture types and sampling flags), film
size, sample count, depth limits and seed, renders it with the CUDA path (counting and plain kernel instantiations) and
with the CPU oracle, and requires EQUAL integer work counters and a film within the parity tolerances (whole-film
relative L1 <= 1e-5, per-time-bin <= 1e-4, no bin that is empty on one side only).  One line per configuration.
"""
import os
import random
import sys
import traceback
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import mitransient_b200 as mitr  # noqa: E402
from oracle import oracle  # noqa: E402
from parity_utils import film_l1  # noqa: E402

T = mitr.ScalarTransform4f
WHOLE_TOL, PER_BIN_TOL = 1e-5, 1e-4


def colour(rng):
    return {"type": "rgb", "value": [round(rng.uniform(0.2, 1.0), 3) for _ in range(3)]}


def some_bsdf(rng, specular):
    if specular:
        return rng.choice([
            {"type": "conductor", "material": "none"},
            {"type": "twosided", "bsdf": {"type": "conductor", "eta": {"type": "rgb", "value": [0.143, 0.375, 1.442]},
                                          "k": {"type": "rgb", "value": [3.983, 2.386, 1.603]}}},
            {"type": "dielectric", "int_ior": rng.choice([1.33, 1.5]), "ext_ior": 1.0},
        ])
    if rng.random() < 0.3:
        return {"type": "twosided", "bsdf": {"type": "diffuse", "reflectance": colour(rng)}}
    return {"type": "diffuse", "reflectance": colour(rng)}


def film_dims(rng):
    w, h = rng.choice([(8, 8), (17, 9), (24, 16), (32, 40), (48, 32)])
    t = rng.choice([16, 64, 200, 600, 1500])
    return w, h, t


def cfg_path(rng, kind):
    w, h, t = film_dims(rng)
    spp = rng.choice([1, 3, 8, 16, 32])
    depth = rng.choice([2, 3, 5, 8, 12, -1])
    window = rng.choice([(3.5, 9.5), (0.0, 20.0), (4.0, 6.0)])
    if kind == "mesh":
        d = mitr.cornell_box_mesh_config(w, h, t, spp, 8, window, merge_boxes=rng.random() < 0.5)
    elif kind == "bvh":
        d = mitr.cornell_box_sphere_config(w, h, t, spp, rng.choice([1, 2, 3]), 8, window)
    else:
        d = mitr.cornell_box_config(w, h, t, spp, 8, window)
    d["integrator"]["max_depth"] = depth
    d["integrator"]["rr_depth"] = rng.choice([1, 2, 5])
    d["integrator"]["camera_unwarp"] = rng.random() < 0.3
    d["integrator"]["discard_direct_light"] = rng.random() < 0.2
    if kind in ("path", "mats") and rng.random() < 0.7:      # move / turn / rescale the boxes
        d["small-box"]["to_world"] = T().translate([rng.uniform(0.1, 0.5), -0.7, rng.uniform(0.1, 0.5)]).rotate(
            [0, 1, 0], rng.uniform(-60, 60)).scale(rng.uniform(0.2, 0.35))
        if rng.random() < 0.5:
            d["large-box"]["to_world"] = T().translate([-0.33, -0.4, -0.28]).rotate([0, 1, 0], rng.choice([0.0, 18.25, 45.0])).scale(
                [0.3, 0.61, 0.3])
        if rng.random() < 0.3:
            d["small-box"]["box_primitive"] = False
    if kind == "mats":
        d["large-box"]["bsdf"] = some_bsdf(rng, True)
        if rng.random() < 0.5:
            d["small-box"]["bsdf"] = some_bsdf(rng, True)
        if rng.random() < 0.3:
            d["back"]["bsdf"] = some_bsdf(rng, True) if rng.random() < 0.5 else some_bsdf(rng, False)
    elif rng.random() < 0.4:
        d["back"]["bsdf"] = some_bsdf(rng, False)
    return d, spp, None


def cfg_vol(rng, kind):
    w, h, t = film_dims(rng)
    spp = rng.choice([1, 4, 8, 16])
    d = mitr.cornell_box_volumetric(w, h, t, spp, rng.choice([4, 8, 16, -1]), rng.choice([(3.5, 13.5), (0.0, 30.0)]))
    d["integrator"]["camera_unwarp"] = rng.random() < 0.3
    d["integrator"]["rr_depth"] = rng.choice([2, 5])
    d["integrator"]["hide_emitters"] = rng.random() < 0.2
    for name in ("small-box", "large-box"):
        m = d[name]["interior"]
        m["sigma_t"] = rng.choice([0.5, 3.0, 14.0])
        m["phase"] = rng.choice([{"type": "isotropic"}, {"type": "hg", "g": rng.choice([-0.5, 0.1, 0.7])}])
    if rng.random() < 0.3:
        d["floor"]["bsdf"] = some_bsdf(rng, True)
    return d, spp, None


def cfg_nlos(rng, kind):
    kw = {"capture_type": rng.choice(["confocal", "single"]), "nlos_laser_sampling": rng.random() < 0.7,
          "nlos_hidden_geometry_sampling": rng.random() < 0.7, "account_first_and_last_bounces": rng.random() < 0.5,
          "discard_direct_paths": rng.random() < 0.2, "max_depth": rng.choice([3, 4, 6, -1]), "rr_depth": rng.choice([2, 5])}
    if kw["nlos_hidden_geometry_sampling"]:
        kw["nlos_hidden_geometry_sampling_do_rroulette"] = rng.random() < 0.5
        kw["nlos_hidden_geometry_sampling_includes_relay_wall"] = rng.random() < 0.3
    w, h = rng.choice([(4, 4), (9, 7), (16, 16), (32, 24)])
    t = rng.choice([64, 300, 1200, 4096])
    spp = rng.choice([4, 16, 64])
    d = mitr.nlos_z_scene(w, h, t, spp, start_opl=rng.choice([0.0, 1.85]), window=rng.choice([1.8, 12.0]), **kw)
    if not kw["nlos_laser_sampling"]:
        d["emitter"]["fov"] = rng.choice([60.0, 120.0])
    for k in range(rng.choice([0, 0, 1, 2])):
        kindk = rng.choice(["rectangle", "cube"])
        pos = [rng.uniform(-0.6, 0.6), rng.uniform(-0.6, 0.6), rng.uniform(0.7, 1.4)]
        sc = rng.uniform(0.1, 0.35)
        tw = T().translate(pos).rotate([0, 1, 0], 180.0 if kindk == "rectangle" else rng.uniform(0, 90)).scale([sc, sc * rng.uniform(0.5, 1.5), sc])
        d["hidden%d" % k] = {"type": kindk, "to_world": tw, "bsdf": some_bsdf(rng, rng.random() < 0.25)}
    if rng.random() < 0.3:
        v, f = mitr.icosphere(rng.choice([1, 2]))
        d["ball"] = {"type": "trimesh", "vertices": v, "faces": f, "merge_quads": False,
                     "to_world": T().translate([0.3, -0.3, 1.1]).scale(0.25), "bsdf": some_bsdf(rng, False)}
    focus = (rng.randrange(w), rng.randrange(h)) if kw["capture_type"] == "single" else None
    return d, spp, focus


def cfg_exhaustive(rng, kind):
    """Exhaustive NLOS capture into an `exhaustive_scan` film (global vector reductions instead of the shared-memory
    histograms), with the sensor's grid or the film's own laser grid."""
    kw = {"capture_type": "exhaustive", "nlos_laser_sampling": rng.random() < 0.7,
          "nlos_hidden_geometry_sampling": rng.random() < 0.6, "account_first_and_last_bounces": rng.random() < 0.5,
          "max_depth": rng.choice([3, 4, 6]), "rr_depth": rng.choice([2, 5])}
    w, h = rng.choice([(2, 2), (3, 4), (5, 4)])
    t = rng.choice([32, 96, 400])
    spp = rng.choice([2, 4, 8])
    d = mitr.nlos_z_scene(w, h, t, spp, start_opl=0.0, window=rng.choice([8.0, 16.0]), **kw)
    if not kw["nlos_laser_sampling"]:
        d["emitter"]["fov"] = rng.choice([60.0, 120.0])
    f = d["relay_wall"]["nlos_sensor"]["transient_film"]
    f["exhaustive_scan"], f["laser_scan_width"], f["laser_scan_height"] = True, w, h
    if kw["nlos_laser_sampling"] and rng.random() < 0.4:       # its own laser grid (transientnlospath.py:349-379)
        f["laser_scan_width"], f["laser_scan_height"] = rng.choice([(2, 2), (3, 2), (2, 4)])
        d["integrator"].update(force_equal_illumination_scanning=False, illumination_scan_fov=rng.choice([10.0, 30.0, 50.0]))
    return d, spp, None


def cfg_phasor(rng, kind):
    """phasor_hdr_film (monochromatic variant) under transient_path on Cornell variants or transient_nlos_path."""
    spp = rng.choice([2, 8, 16])
    if rng.random() < 0.5:
        w, h = rng.choice([(8, 8), (17, 9), (24, 16)])
        bins = rng.choice([64, 256, 1024])
        d = mitr.cornell_box_config(w, h, bins, spp, rng.choice([3, 8]))
        wl = rng.choice([(0.5, 0.5), (0.2, 0.1), (1.0, 0.3)])
        d["sensor"]["film"] = {"type": "phasor_hdr_film", "width": w, "height": h, "temporal_bins": bins, "start_opl": 3.5,
                               "bin_width_opl": 6.0 / bins, "wl_mean": wl[0], "wl_sigma": wl[1], "rfilter": {"type": "box"}}
    else:
        w, h = rng.choice([(4, 4), (9, 7), (16, 16)])
        bins = rng.choice([100, 300])
        d = mitr.nlos_z_scene(w, h, bins, spp, nlos_hidden_geometry_sampling=rng.random() < 0.6,
                              account_first_and_last_bounces=rng.random() < 0.5)
        wl = rng.choice([(0.06, 0.06), (0.1, 0.03)])
        d["relay_wall"]["nlos_sensor"]["transient_film"] = {
            "type": "phasor_hdr_film", "width": w, "height": h, "temporal_bins": bins, "start_opl": 1.85,
            "bin_width_opl": 1.8 / bins, "wl_mean": wl[0], "wl_sigma": wl[1], "rfilter": {"type": "box"}}
    return d, spp, None


KINDS = {"path": cfg_path, "mesh": cfg_path, "bvh": cfg_path, "mats": cfg_path, "vol": cfg_vol, "nlos": cfg_nlos,
         "exh": cfg_exhaustive, "phasor": cfg_phasor}


def one(rng, kind):
    mitr.set_variant("cuda_mono" if kind == "phasor" else "cuda_rgb")
    d, spp, focus = KINDS[kind](rng, kind)
    seed = rng.randrange(1000)
    scene = mitr.load_dict(d)
    if focus is not None:
        mitr.nlos.focus_emitter_at_relay_wall_pixel(focus, mitr.PluginSpec(d["relay_wall"]), scene.emitters()[0]["spec"])
    integ, sensor = scene.integrator(), scene.sensors()[0]
    if os.environ.get("FUZZ_DRY"):          # generator check on a box without a GPU: the oracle alone
        integ.check_transient_(scene, sensor); integ.check_scene_(scene, sensor)
        try:
            ref = oracle.render(scene, integ, seed=seed, spp=spp)
            print("DRY ", kind, ref["counters"], flush=True)
        except RuntimeError as exc:
            print("DRY ", kind, "oracle refuses:", exc, flush=True)
        return True
    integ.collect_counters = True
    # the counting render prefers one histogram path, the plain render below the other (honoured where the scene / film
    # allow both, tr_last_render_info says what ran)
    prefs = ("shared", "global") if seed % 2 else ("global", "shared")
    integ.histogram_path = prefs[0]
    try:
        steady, _ = integ.render(scene, seed=seed, spp=spp)
    except (mitr.TransientError, ValueError, AssertionError) as exc:
        # a configuration the reference refuses (e.g. a laser grid that misses the scene): the oracle must refuse it too
        try:
            oracle.render(scene, integ, seed=seed, spp=spp)
        except RuntimeError as oexc:
            print("OK  ", kind, "both refuse:", str(exc)[:60], "|", str(oexc)[:60], flush=True)
            return True
        print("FAIL", kind, "only the GPU path refuses:", exc, flush=True)
        return False
    counted = dict(integ.last_counters)
    ran = integ.last_render_info()["histogram_path"][0]
    raw = sensor.film().develop(raw=True)[1].cpu().numpy()
    ref = oracle.render(scene, integ, seed=seed, spp=spp)
    ok = counted == ref["counters"]

    def l1(a, b):           # exhaustive_scan films are (H, W, Lh, Lw, T, C): fold the laser grid into the pixel axis
        if a.ndim == 3:     # phasor film (H, W, 2F + 1): one "time bin" per stored component
            a, b = a[:, :, :, None], b[:, :, :, None]
        if a.ndim == 6:
            a, b = a.reshape(a.shape[0], -1, a.shape[4], a.shape[5]), b.reshape(b.shape[0], -1, b.shape[4], b.shape[5])
        return film_l1(a, b)
    err = l1(raw, ref["transient"])
    ok = ok and err["zero_bins_mismatch"] == 0 and err["whole"] <= WHOLE_TOL and err["max_t"] <= PER_BIN_TOL
    if kind == "phasor":    # every finite deposit counts (no time window): compare component-wise too
        ok = ok and np.allclose(raw, ref["transient"], rtol=1e-4, atol=1e-6)
        ok = ok and np.allclose(steady.cpu().numpy()[..., 0], ref["steady"][..., 0], rtol=1e-4, atol=1e-6)
    elif kind != "exh":     # (an exhaustive film's develop() returns the channel mean as "steady", transient_hdr_film.py:213-214)
        ok = ok and np.allclose(steady.cpu().numpy(), ref["steady"], rtol=1e-4, atol=1e-6)
    integ.collect_counters = False
    integ.histogram_path = prefs[1]
    integ.render(scene, seed=seed, spp=spp)
    ran += integ.last_render_info()["histogram_path"][0]
    raw2 = sensor.film().develop(raw=True)[1].cpu().numpy()
    err2 = l1(raw2, ref["transient"])
    ok = ok and err2["zero_bins_mismatch"] == 0 and err2["whole"] <= WHOLE_TOL
    info = integ.device_scene(scene, sensor).info()
    film = sensor.film()
    fw, fh = film.size()
    print(("OK  " if ok else "FAIL"), kind, f"prims={info['primitives']} nodes={info['bvh_nodes']} film={fw}x{fh}x{film.temporal_bins} spp={spp} "
          f"seed={seed} depth={integ.max_depth if integ.max_depth < 1 << 31 else -1} vertices={counted['vertices']} deposits={counted['deposits_in_range']} "
          f"L1={err['whole']:.1e}/{err2['whole']:.1e} hist={ran}", flush=True)
    if not ok:
        print("     counters gpu", counted, "\n     counters ref", ref["counters"], "\n     err", err, err2, flush=True)
    return ok


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    kinds = sys.argv[3].split(",") if len(sys.argv) > 3 else list(KINDS)
    warnings.simplefilter("ignore")
    bad = 0
    for i in range(n):
        rng = random.Random(first + i)
        kind = kinds[i % len(kinds)]
        try:
            bad += 0 if one(rng, kind) else 1
        except Exception:
            bad += 1
            print("FAIL", kind, "seed", first + i, "exception:", flush=True)
            traceback.print_exc()
    print(f"# {n - bad} / {n} configurations equal (first seed {first}, kinds {','.join(kinds)})")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
