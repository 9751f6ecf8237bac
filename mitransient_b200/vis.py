"""Visualisation helpers of ``mitransient.vis`` that sit directly on the film output
(mitransient/unpolarized_visualization.py:14-18).  Video / frame writers need OpenCV + a Mitsuba Bitmap and are
not part of this build."""
from __future__ import annotations

import numpy as np


def _quantile_abs(t, q: float) -> float:
    """``np.quantile(np.abs(t), q)`` (linear interpolation) for a torch tensor of any size: two radix selects
    (``kthvalue``) instead of a full sort, so a 3 GB developed film stays on the device."""
    import torch
    flat = t.detach().abs().reshape(-1)
    n = flat.numel()
    pos = q * (n - 1)
    lo = int(np.floor(pos))
    hi = min(lo + 1, n - 1)
    v_lo = float(torch.kthvalue(flat, lo + 1).values)
    v_hi = float(torch.kthvalue(flat, hi + 1).values) if hi != lo else v_lo
    return v_lo + (v_hi - v_lo) * (pos - lo)


def tonemap_transient(transient, scaling: float = 1.0):
    """Linear tonemapping of a transient image: divide by the 99th percentile of ``|transient|``
    (unpolarized_visualization.py:14-18).  Accepts the CUDA tensor returned by ``render`` (stays on the device)
    or anything ``np.array`` understands (returns a NumPy array like the reference)."""
    try:
        import torch
        if isinstance(transient, torch.Tensor):
            return transient / _quantile_abs(transient, 0.99) * scaling
    except ImportError:
        pass
    tnp = np.array(transient)
    channel_top = np.quantile(np.abs(tnp), 0.99)
    return tnp / channel_top * scaling
