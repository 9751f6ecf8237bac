"""Runs the UNMODIFIED reference (mitransient on mitsuba/drjit) on the benchmark configurations and saves
goldens + timings.  NOT RUNNABLE IN THIS IMAGE: mitsuba >=3.6,<3.9 and drjit are not installed and cannot be
(no wheel in /opt/wheelhouse, no network) -- see DESIGN.md section 2.  Kept so that parity can be pinned the
day a Mitsuba install is available:

    python baseline/run_reference.py --config C1 --out tests/golden/reference_c1.npz
"""
import argparse
import time

import numpy as np

CONFIGS = {
    "C1": dict(width=64, height=64, temporal_bins=256, spp=16, max_depth=8),
    "C2": dict(width=512, height=512, temporal_bins=1024, spp=256, max_depth=8),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C1", choices=sorted(CONFIGS))
    ap.add_argument("--variant", default="llvm_ad_rgb")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    import drjit as dr
    import mitsuba as mi
    mi.set_variant(a.variant)
    import mitransient as mitr

    c = CONFIGS[a.config]
    d = mitr.cornell_box()
    film = d["sensor"]["film"]
    film["width"], film["height"], film["temporal_bins"] = c["width"], c["height"], c["temporal_bins"]
    film["start_opl"], film["bin_width_opl"] = 3.5, 6.0 / c["temporal_bins"]
    d["integrator"]["max_depth"] = c["max_depth"]
    scene = mi.load_dict(d)
    times = []
    for rep in range(3):  # the first call includes JIT compilation
        t0 = time.perf_counter()
        steady, transient = mi.render(scene, spp=c["spp"], seed=0)
        dr.eval(steady, transient)
        dr.sync_thread()
        times.append(time.perf_counter() - t0)
    paths = c["width"] * c["height"] * c["spp"]
    print(f"{a.config} {a.variant}: {[round(t, 3) for t in times]} s, "
          f"{paths / min(times[1:]) / 1e6:.2f} M-paths/s (excluding the first, JIT-compiling call)")
    if a.out:
        raw = scene.sensors()[0].film().develop(raw=True)[1]
        np.savez_compressed(a.out, steady=np.array(steady), transient=np.array(transient), raw=np.array(raw))


if __name__ == "__main__":
    main()
