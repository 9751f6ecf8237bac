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


def film_l1_device(film_dev, ref: np.ndarray, rows: int = 32):
    """film_l1 for films too large for a host-side float64 pass (4 GiB at 512x512x1024): `film_dev` is a CUDA tensor
    (H, W, T, C), `ref` the oracle's numpy film; streamed over row blocks, reduced in float64 on the device."""
    import torch
    H, T = film_dev.shape[0], film_dev.shape[2]
    num = torch.zeros(T, dtype=torch.float64, device="cuda"); den = torch.zeros_like(num); mag = torch.zeros_like(num)
    for y0 in range(0, H, rows):
        r = torch.from_numpy(ref[y0:y0 + rows]).cuda().double()
        g = film_dev[y0:y0 + rows].double()
        num += (g - r).abs().sum(dim=(0, 1, 3)); den += r.abs().sum(dim=(0, 1, 3)); mag += g.abs().sum(dim=(0, 1, 3))
    nz = den > 0
    per_t = num[nz] / den[nz]
    return {"whole": float(num.sum() / max(float(den.sum()), 1e-300)),
            "max_t": float(per_t.max()) if per_t.numel() else 0.0,
            "mean_t": float(per_t.mean()) if per_t.numel() else 0.0,
            "zero_bins_mismatch": int(((mag > 0) != nz).sum())}
