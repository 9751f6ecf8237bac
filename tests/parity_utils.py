"""Parity metrics of BASELINE.md section 6: per-time-bin relative L1 and whole-film relative L1."""
import numpy as np


def film_l1(new: np.ndarray, ref: np.ndarray):
    """new/ref: (H, W, T, C).  Returns dict(whole, max_t, mean_t) of relative L1 errors."""
    new = np.asarray(new, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    num_t = np.abs(new - ref).sum(axis=(0, 1, 3))
    den_t = np.abs(ref).sum(axis=(0, 1, 3))
    nz = den_t > 0
    per_t = num_t[nz] / den_t[nz]
    return {
        "whole": float(num_t.sum() / max(den_t.sum(), 1e-300)),
        "max_t": float(per_t.max()) if per_t.size else 0.0,
        "mean_t": float(per_t.mean()) if per_t.size else 0.0,
        "zero_bins_mismatch": int(((np.abs(new).sum(axis=(0, 1, 3)) > 0) != nz).sum()),
    }
