"""N > 1 host logic on CPU: two gloo ranks each render their pixel tiles, one SUM-reduce of the film
reproduces the single-rank film bit for bit (ownership is disjoint, so the sum is x + 0).  The renderer in
this test is the CPU oracle, which implements the same (tile_size, rank, world) ownership rule as the CUDA
library; the GPU version of this check is tests/test_gpu_multi.py."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, outdir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import mitransient_b200 as mitr
    from mitransient_b200.distributed import reduce_film, tile_owner_map
    from oracle import oracle
    scene = mitr.load_dict(mitr.cornell_box_config(40, 24, 64, 4))
    r = oracle.render(scene, scene.integrator(), spp=4, seed=0, rank=rank, world=world, n_threads=2)
    film = torch.from_numpy(r["transient"])
    steady = torch.from_numpy(r["steady"])
    own = tile_owner_map(40, 24, 16, world) == rank
    assert np.all(r["transient"][~own] == 0)                     # nothing written outside the owned tiles
    reduce_film(film, steady, dst=0)
    cnt = torch.tensor([r["counters"]["paths"], r["counters"]["vertices"]], dtype=torch.int64)
    dist.all_reduce(cnt)
    if rank == 0:
        np.save(os.path.join(outdir, "film.npy"), film.numpy())
        np.save(os.path.join(outdir, "steady.npy"), steady.numpy())
        np.save(os.path.join(outdir, "cnt.npy"), cnt.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_tile_sharding_reduces_to_single_rank_film(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, ROOT)
    import mitransient_b200 as mitr
    from oracle import oracle
    scene = mitr.load_dict(mitr.cornell_box_config(40, 24, 64, 4))
    ref = oracle.render(scene, scene.integrator(), spp=4, seed=0)
    film = np.load(tmp_path / "film.npy")
    assert np.array_equal(film, ref["transient"])                # KAT-7: independent of the number of ranks
    assert np.array_equal(np.load(tmp_path / "steady.npy"), ref["steady"])
    cnt = np.load(tmp_path / "cnt.npy")
    assert cnt[0] == ref["counters"]["paths"] and cnt[1] == ref["counters"]["vertices"]


def test_tile_owner_map():
    sys.path.insert(0, ROOT)
    from mitransient_b200.distributed import tile_owner_map
    m = tile_owner_map(70, 40, 16, 4)
    assert m.shape == (40, 70)
    assert m[0, 0] == 0 and m[0, 16] == 1 and m[0, 32] == 2 and m[16, 0] == 2 and m[16, 16] == 3     # (tile + tile_row) % world, 5 tiles per row
    counts = np.bincount(m.reshape(-1), minlength=4)
    assert counts.sum() == 70 * 40 and np.all(counts > 0)


def test_owned_pixels_is_the_owner_map_in_tile_order():
    """tr_owned_pixels (the layout of the sharded tr_render_to_host output) partitions the film exactly like
    tile_owner_map; no GPU needed."""
    sys.path.insert(0, ROOT)
    from mitransient_b200.distributed import owned_pixels, tile_owner_map
    W, H, tile, world = 70, 40, 16, 3
    m = tile_owner_map(W, H, tile, world).reshape(-1)
    seen = np.zeros(W * H, dtype=np.int64)
    for rank in range(world):
        px = owned_pixels(W, H, tile, rank, world)
        assert np.array_equal(np.sort(px), np.nonzero(m == rank)[0])
        seen[px] += 1
        # tile-major order: the first pixels are the first owned tile, row by row
        first_tile = np.nonzero((m == rank))[0][0]
        assert px[0] == first_tile and px[1] == first_tile + 1
    assert (seen == 1).all()
    assert len(owned_pixels(W, H, tile, 0, 1)) == W * H
