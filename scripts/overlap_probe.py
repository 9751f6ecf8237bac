"""Does a large cudaMemsetAsync on a second stream overlap with the persistent render kernel?  (experiment, run 49)"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import mitransient_b200 as mitr
from mitransient_b200 import _lib

rt = C.CDLL("libcudart.so.12")
rt.cudaMemsetAsync.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p]
W = H = 512; T = 1024; spp = 256
scene = mitr.load_dict(mitr.cornell_box_config(W, H, T, spp, 8))
integ, sensor = scene.integrator(), scene.sensors()[0]
dev = integ.device_scene(scene, sensor)
lib = _lib.load()
film = torch.zeros((H, W, T, 4), dtype=torch.float32, device="cuda")
other = torch.zeros((H, W, T, 4), dtype=torch.float32, device="cuda")
steady = torch.zeros((H, W, 3), dtype=torch.float32, device="cuda")
integ.histogram_path = "global"
params = integ.make_params(spp, spp, 0)
integ.fill_params_(params, scene, sensor)
params.accumulate = 1                       # no memset inside tr_render: the kernel alone
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
nbytes = other.numel() * 4


def run(kernel, memset, reps=5):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start = torch.cuda.Event()
    e0.record(s1); start.record(s1)
    for _ in range(reps):
        s2.wait_stream(s1)
        if memset:
            assert rt.cudaMemsetAsync(other.data_ptr(), 0, nbytes, s2.cuda_stream) == 0
        if kernel:
            _lib.check(lib.tr_render(dev.handle, C.byref(params), film.data_ptr(), steady.data_ptr(), None, s1.cuda_stream))
        s1.wait_stream(s2)
    e1.record(s1)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for _ in range(2):
    run(True, True, 2)
print(f"kernel alone {run(True, False):.3f} ms | memset alone {run(False, True):.3f} ms | both, two streams {run(True, True):.3f} ms")
print(integ.last_render_info())
