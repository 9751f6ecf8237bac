"""CUDA-event timing of tr_render alone (the bench's `value` leg) -- used for A/B runs of kernel variants.
    [MITRANSIENT_B200_LIB=path/to/variant.so] python scripts/time_kernel.py [--steps 10]"""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import mitransient_b200 as mitr
from mitransient_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--width", type=int, default=512)
ap.add_argument("--height", type=int, default=512)
ap.add_argument("--bins", type=int, default=1024)
ap.add_argument("--spp", type=int, default=256)
ap.add_argument("--depth", type=int, default=8)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--scene", default="cornell", choices=["cornell", "nlos", "vol", "mesh", "sphere0", "sphere1", "sphere2", "sphere3", "sphere4"])
ap.add_argument("--hist", default="auto", choices=["auto", "shared", "global"], help="TrRenderParams.histogram_path")
ap.add_argument("--world", type=int, default=1, help="render only the tiles of rank 0 of `world` ranks (the per-GPU share of an N-GPU job)")
a = ap.parse_args()
if a.scene == "nlos":          # BASELINE config C3: letter Z, confocal, 256x256 x 4096 bins, 1024 spp
    a.width = a.height = 256 if a.width == 512 else a.width
    a.bins = 4096 if a.bins == 1024 else a.bins
    a.spp = 1024 if a.spp == 256 else a.spp
    scene = mitr.load_dict(mitr.nlos_z_scene(a.width, a.height, a.bins, a.spp))
elif a.scene == "vol":         # BASELINE config C4(i): volumetric Cornell box 256x384 x 512 bins
    a.width, a.height, a.bins = 256, 384, 512
    scene = mitr.load_dict(mitr.cornell_box_volumetric(a.width, a.height, a.bins, a.spp))
elif a.scene == "mesh":        # the Cornell box as 36 mesh triangles (18 parallelograms, brute-force loop)
    a.spp = 64 if a.spp == 256 else a.spp
    scene = mitr.load_dict(mitr.cornell_box_mesh_config(a.width, a.height, a.bins, a.spp, a.depth))
elif a.scene.startswith("sphere"):   # Cornell box + icosphere: the BVH path (320 / 1280 / 5120 triangles)
    scene = mitr.load_dict(mitr.cornell_box_sphere_config(a.width, a.height, a.bins, a.spp, int(a.scene[-1]), a.depth))
else:
    scene = mitr.load_dict(mitr.cornell_box_config(a.width, a.height, a.bins, a.spp, a.depth))
integ, sensor = scene.integrator(), scene.sensors()[0]
dev = integ.device_scene(scene, sensor)
lib = _lib.load()
film = torch.zeros((a.height, a.width, a.bins, 4), dtype=torch.float32, device="cuda")
steady = torch.zeros((a.height, a.width, 3), dtype=torch.float32, device="cuda")
params = integ.make_params(a.spp, a.spp, 0, rank=0, world=a.world)
integ.fill_params_(params, scene, sensor)
params.histogram_path = {"auto": 0, "shared": 1, "global": 2}[a.hist]
st = torch.cuda.current_stream()
for _ in range(3):
    _lib.check(lib.tr_render(dev.handle, C.byref(params), film.data_ptr(), steady.data_ptr(), None, st.cuda_stream))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    _lib.check(lib.tr_render(dev.handle, C.byref(params), film.data_ptr(), steady.data_ptr(), None, st.cuda_stream))
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
print(f"{os.environ.get('MITRANSIENT_B200_LIB', 'default'):40s} {ms:8.3f} ms/step  "
      f"{a.width * a.height * a.spp / a.world / ms / 1e3:9.1f} M-paths/s  film.sum={film.sum().item():.6e}")
