"""Regenerates tests/golden/cornell_c1_oracle.npz from the CPU oracle (config C1 of BASELINE.md).

These are NOT outputs of the reference (mitsuba/drjit cannot be installed here, parity stays *unpinned*):
they pin the oracle against accidental change, and are what the GPU path is compared with at C1.
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import mitransient_b200 as mitr  # noqa: E402
from oracle import oracle  # noqa: E402

scene = mitr.load_dict(mitr.cornell_box_config(64, 64, 256, 16))
r = oracle.render(scene, scene.integrator(), spp=16, seed=0)
out = {"steady": r["steady"], "temporal_profile": r["transient"][..., :3].sum(axis=(0, 1)),
       "pixel_32_32": r["transient"][32, 32, :, :3]}
for k, v in r["counters"].items():
    out["counter_" + k] = np.uint64(v)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "cornell_c1_oracle.npz"), **out)
print({k: (v.shape if hasattr(v, "shape") and v.shape else v) for k, v in out.items()})
