"""Time the NLOS letter-Z configuration (BASELINE config C3: 256x256 relay wall x 4096 bins, confocal)."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import mitransient_b200 as mitr

ap = argparse.ArgumentParser()
ap.add_argument("--res", type=int, default=256)
ap.add_argument("--bins", type=int, default=4096)
ap.add_argument("--spp", type=int, default=1024)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--counters", action="store_true")
a = ap.parse_args()
scene = mitr.load_dict(mitr.nlos_z_scene(a.res, a.res, a.bins, a.spp))
integ = scene.integrator()
integ.collect_counters = a.counters
for i in range(a.reps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    steady, transient = mitr.render(scene, spp=a.spp, seed=0)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"rep {i}: {dt * 1e3:.2f} ms (render + develop), {a.res * a.res * a.spp / dt / 1e6:.1f} M-paths/s, {integ.last_counters}",
          flush=True)
