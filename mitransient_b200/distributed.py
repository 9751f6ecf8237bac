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
    return ((ty * tiles_x + tx + ty) % world).astype(np.int64)     # round-robin, skewed by one tile per tile row


def owned_pixels(width: int, height: int, tile_size: int, rank: int, world: int) -> np.ndarray:
    """Pixel ids ``y * W + x`` owned by ``rank`` in the order ``tr_render_to_host`` lays out its compact per-rank
    film (``tr_owned_pixels``: owned tiles row-major, pixels row-major inside a tile).  Needs no GPU."""
    import ctypes as C

    from . import _lib
    lib = _lib.load()
    n = C.c_uint64(0)
    _lib.check(lib.tr_owned_pixels(width, height, tile_size, rank, world, None, C.byref(n)))
    out = np.zeros(int(n.value), dtype=np.uint32)
    _lib.check(lib.tr_owned_pixels(width, height, tile_size, rank, world, out.ctypes.data, C.byref(n)))
    return out


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


class _DeviceArray:
    """A raw device pointer with a shape: ``data_ptr()`` for the C ABI, ``__cuda_array_interface__`` for torch."""

    def __init__(self, ptr: int, shape):
        self._ptr, self.shape = int(ptr), tuple(int(x) for x in shape)
        self.__cuda_array_interface__ = {"shape": self.shape, "typestr": "<f4", "data": (self._ptr, False), "version": 3}

    def data_ptr(self) -> int:
        return self._ptr


class PeerFilm:
    """The film of rank ``dst`` mapped into every rank's address space (CUDA IPC over NVLink peer memory).

    With pixel-tile ownership every film byte has exactly one writer, so instead of rendering into N private
    films and reducing them, each rank's render kernel flushes its finished pixel histograms STRAIGHT INTO the
    destination rank's film with 128-bit stores over NVLink: the transfer is fused into the kernel's flush and
    overlaps the path tracing; no reduce, no memset, no N-fold film memory.  One tiny NCCL all-reduce at the end
    orders the remote stores before ``dst`` reads the film.

    Contract when the film is REUSED: there is one buffer and no double buffering, so a new render may only start once
    ``dst`` has finished reading the previous film.  :func:`render_distributed_peer` fences on entry as well (the entry
    fence is enqueued on ``dst``'s stream after its reads); callers that launch ``tr_render`` themselves (bench.py)
    must do the same whenever something reads the film between two renders."""

    def __init__(self, shape_film, shape_steady, dst: int = 0):
        import ctypes as C

        import torch
        import torch.distributed as dist

        from . import _abi, _lib
        self._lib = _lib.load()
        self.dst, self.rank = dst, dist.get_rank()
        n_film = int(np.prod(shape_film)) * 4
        n_steady = int(np.prod(shape_steady)) * 4
        self._film_bytes = (n_film + 255) // 256 * 256
        total = self._film_bytes + n_steady
        ptr = C.c_void_p()
        payload = [None]
        if self.rank == dst:
            handle = C.create_string_buffer(_abi.TR_PEER_HANDLE_BYTES)
            _lib.check(self._lib.tr_peer_alloc(total, C.byref(ptr), handle))
            payload = [handle.raw]
        dist.broadcast_object_list(payload, src=dst)
        if self.rank != dst:
            _lib.check(self._lib.tr_peer_open(payload[0], C.byref(ptr)))
        self._ptr = ptr.value
        self.film = _DeviceArray(self._ptr, shape_film)
        self.steady = _DeviceArray(self._ptr + self._film_bytes, shape_steady)
        self._nccl = dist.get_backend() == "nccl"
        self._flag = torch.zeros(1, device="cuda" if self._nccl else "cpu")

    def tensors(self):
        """(steady, film) as torch tensors -- only meaningful on ``dst`` (the memory lives on its device)."""
        import torch
        return (torch.as_tensor(self.steady, device="cuda"), torch.as_tensor(self.film, device="cuda"))

    def close(self):
        if getattr(self, "_ptr", None):
            (self._lib.tr_peer_free if self.rank == self.dst else self._lib.tr_peer_close)(self._ptr)
            self._ptr = None

    def fence(self):
        """Stream-ordered after this rank's render kernel: when it completes on ``dst``, every rank's stores have
        been issued and completed (kernel completion flushes them) before anything enqueued later reads the film.
        (On a non-NCCL backend -- the two-ranks-on-one-GPU test -- the stream is drained on the host first.)"""
        import torch
        import torch.distributed as dist
        if not self._nccl:
            torch.cuda.current_stream().synchronize()
        dist.all_reduce(self._flag)


def render_distributed_peer(scene, peer: PeerFilm, sensor=0, seed: int = 0, spp: int = 0):
    """Fused variant of :func:`render_distributed`: every rank writes its tiles directly into ``peer.film``
    (owned by ``peer.dst``) over NVLink; returns when the film is complete on ``peer.dst``."""
    import ctypes as C

    import torch
    import torch.distributed as dist

    from . import _lib
    rank, world = dist.get_rank(), dist.get_world_size()
    integ = scene.integrator()
    s = scene.sensors()[sensor] if isinstance(sensor, int) else sensor
    integ.check_transient_(scene, s)
    integ.check_scene_(scene, s)
    s.film().check_supported()
    sampler = s.sampler().clone()
    if spp:
        sampler.set_sample_count(spp)
    spp = sampler.sample_count()
    dev = integ.device_scene(scene, s)
    params = integ.make_params(spp, spp, (int(seed) + sampler.base_seed) & 0xFFFFFFFF, rank=rank, world=world,
                               film_channels=int(peer.film.shape[-1]))
    integ.fill_params_(params, scene, s)
    # The film is ONE buffer that every rank overwrites: order this render after whatever `dst` enqueued on its stream to
    # read the previous one (develop, device->host copy) -- without this fence another rank's kernel could store into the
    # film while `dst` is still reading it (write-after-read).
    peer.fence()
    _lib.check(_lib.load().tr_render(dev.handle, C.byref(params), peer.film.data_ptr(), peer.steady.data_ptr(), None,
                                     torch.cuda.current_stream().cuda_stream))
    peer.fence()
    return peer.tensors() if rank == peer.dst else (None, None)
