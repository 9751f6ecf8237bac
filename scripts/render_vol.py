"""Time the volumetric Cornell configuration (BASELINE config C4-i: 256x384 x 512 bins)."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import mitransient_b200 as mitr

ap = argparse.ArgumentParser()
ap.add_argument("--width", type=int, default=256)
ap.add_argument("--height", type=int, default=384)
ap.add_argument("--bins", type=int, default=512)
ap.add_argument("--spp", type=int, default=64)
ap.add_argument("--depth", type=int, default=16)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--counters", action="store_true")
a = ap.parse_args()
scene = mitr.load_dict(mitr.cornell_box_volumetric(a.width, a.height, a.bins, a.spp, max_depth=a.depth))
integ = scene.integrator()
integ.collect_counters = a.counters
for i in range(a.reps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    steady, transient = mitr.render(scene, spp=a.spp, seed=0)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"rep {i}: {dt * 1e3:.2f} ms (render + develop), {a.width * a.height * a.spp / dt / 1e6:.1f} M-paths/s, "
          f"{integ.last_counters}", flush=True)
