"""N > 1 on GPUs: two NCCL ranks render their pixel tiles, one reduce of the film to rank 0 reproduces the
single-GPU film (KAT-7: the film does not depend on the number of ranks).  Skipped on a 1-GPU box; the CPU
version of this check (gloo + oracle) is tests/test_dist_gloo.py."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, outdir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import mitransient_b200 as mitr
    from mitransient_b200.distributed import render_distributed, tile_owner_map
    scene = mitr.load_dict(mitr.cornell_box_config(96, 64, 128, 8))
    integ = scene.integrator()
    integ.render(scene, seed=0, spp=8, rank=rank, world=world)
    part = scene.sensors()[0].film().develop(raw=True)[1].cpu().numpy()
    own = tile_owner_map(96, 64, 16, world) == rank
    assert np.all(part[~own] == 0) and np.abs(part[own]).sum() > 0
    steady, raw = render_distributed(scene, seed=0, spp=8, dst=0, raw=True)
    torch.cuda.synchronize()
    if rank == 0:
        np.save(os.path.join(outdir, "film.npy"), raw.cpu().numpy())
        np.save(os.path.join(outdir, "steady.npy"), steady.cpu().numpy())
    # fused variant: every rank flushes its pixels straight into rank 0's film over NVLink peer memory
    from mitransient_b200.distributed import PeerFilm, render_distributed_peer
    peer = PeerFilm((64, 96, 128, 4), (64, 96, 3), dst=0)
    st_p, film_p = render_distributed_peer(scene, peer, seed=0, spp=8)
    torch.cuda.synchronize()
    if rank == 0:
        np.save(os.path.join(outdir, "film_peer.npy"), film_p.cpu().numpy())
        np.save(os.path.join(outdir, "steady_peer.npy"), st_p.cpu().numpy())
    dist.barrier()
    peer.close()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_gpu_tile_sharding_matches_single_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    import mitransient_b200 as mitr
    scene = mitr.load_dict(mitr.cornell_box_config(96, 64, 128, 8))
    steady, _ = scene.integrator().render(scene, seed=0, spp=8)
    raw = scene.sensors()[0].film().develop(raw=True)[1].cpu().numpy()
    film2 = np.load(tmp_path / "film.npy")
    # disjoint ownership: the reduce adds zeros, so the 2-GPU film equals the 1-GPU film up to the order of
    # the fp32 adds inside a bin (which the CTA schedule may change)
    num = np.abs(film2.astype(np.float64) - raw).sum()
    assert num / np.abs(raw).sum() <= 1e-6
    np.testing.assert_allclose(np.load(tmp_path / "steady.npy"), steady.cpu().numpy(), rtol=1e-5, atol=1e-7)
    film3 = np.load(tmp_path / "film_peer.npy")
    assert np.abs(film3.astype(np.float64) - raw).sum() / np.abs(raw).sum() <= 1e-6
    np.testing.assert_allclose(np.load(tmp_path / "steady_peer.npy"), steady.cpu().numpy(), rtol=1e-5, atol=1e-7)


def _worker_one_gpu(rank, world, port, outdir):
    """Two processes, BOTH on cuda:0 (gloo for the rendezvous / fence): rank 1 maps rank 0's film through CUDA IPC and
    its render kernel flushes its tiles into it -- the same kernel code path as the NVLink peer stores."""
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import mitransient_b200 as mitr
    from mitransient_b200.distributed import PeerFilm, render_distributed_peer
    scene = mitr.load_dict(mitr.cornell_box_config(96, 64, 128, 8))
    peer = PeerFilm((64, 96, 128, 4), (64, 96, 3), dst=0)
    for rep in range(2):                       # the film is reused: the second render overwrites the first
        st_p, film_p = render_distributed_peer(scene, peer, seed=rep, spp=8)
        torch.cuda.synchronize()
        if rank == 0:
            np.save(os.path.join(outdir, f"film_peer{rep}.npy"), film_p.cpu().numpy())
            np.save(os.path.join(outdir, f"steady_peer{rep}.npy"), st_p.cpu().numpy())
    dist.barrier()
    peer.close()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_ranks_on_one_gpu_assemble_the_film_through_ipc(tmp_path):
    """KAT-7 on the 1-GPU test box: a film assembled by two ranks (tile sharding, peer film through CUDA IPC) equals the
    single-rank film; the sample streams are indexed globally, so only the fp32 add order inside a bin may differ."""
    import torch.multiprocessing as mp
    mp.spawn(_worker_one_gpu, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    import mitransient_b200 as mitr
    scene = mitr.load_dict(mitr.cornell_box_config(96, 64, 128, 8))
    for rep in range(2):
        steady, _ = scene.integrator().render(scene, seed=rep, spp=8)
        raw = scene.sensors()[0].film().develop(raw=True)[1].cpu().numpy()
        film = np.load(tmp_path / f"film_peer{rep}.npy")
        assert np.abs(film.astype(np.float64) - raw).sum() / np.abs(raw).sum() <= 1e-6
        np.testing.assert_allclose(np.load(tmp_path / f"steady_peer{rep}.npy"), steady.cpu().numpy(), rtol=1e-5, atol=1e-7)
