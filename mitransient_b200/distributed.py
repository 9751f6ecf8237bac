"""Multi-GPU plumbing (one process per GPU, torch.distributed): pixel-tile ownership and the film reduce.

Paths never interact and, with the box reconstruction filter, a path only deposits into its own pixel's
(T, C) column (mitransient/render/transient_image_block.py:131,142-146), so pixels shard with no data-path
collective.  Tiles of ``tile_size x tile_size`` pixels are dealt round-robin to the ranks (the C ABI's
``tile_size / rank / world``); RNG streams are indexed by the global sample index, so the film does not
depend on the number of ranks.  The only collective is ONE sum-reduce of the film (and of the small steady
image) to rank 0 at the end; because ownership is disjoint the sum is exact (x + 0).
"""
from __future__ import annotations

import numpy as np


def tile_owner_map(width: int, height: int, tile_size: int, world: int) -> np.ndarray:
    """(H, W) int array: rank owning each pixel (same rule as csrc/tr_render.cu ensure_pixel_list)."""
    tiles_x = (width + tile_size - 1) // tile_size
    ty, tx = np.meshgrid(np.arange(height) // tile_size, np.arange(width) // tile_size, indexing="ij")
    return ((ty * tiles_x + tx) % world).astype(np.int64)


def reduce_film(film, steady=None, dst: int = 0):
    """The single collective of the job: sum-reduce the raw film (and steady image) to ``dst``.
    NCCL when the tensors are on the GPU, gloo on the CPU."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    dist.reduce(film, dst=dst, op=dist.ReduceOp.SUM)
    if steady is not None:
        dist.reduce(steady, dst=dst, op=dist.ReduceOp.SUM)


def render_distributed(scene, sensor=0, seed: int = 0, spp: int = 0, dst: int = 0, raw: bool = False):
    """Every rank renders its tiles (``integrator.render(..., rank, world)``), then one reduce to ``dst``.
    Returns ``(steady, transient)`` on ``dst`` (other ranks get their partial films)."""
    import torch.distributed as dist
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    integ = scene.integrator()
    integ.render(scene, sensor=sensor, seed=seed, spp=spp, rank=rank, world=world, evaluate=False)
    s = scene.sensors()[sensor] if isinstance(sensor, int) else sensor
    film = s.film()
    reduce_film(film.transient_storage, film.steady_storage, dst=dst)
    return film.develop(raw=raw)
