"""world_size-2 gloo test of the multi-GPU host logic (partition -> per-rank render -> gather), run on CPU.

The per-rank renderer is the oracle here (this is a test; the GPU box runs the same parallel.py with the CUDA
renderer under NCCL — see tests/test_gpu_*.py and bench.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import orc

W, H, SPP, DEPTH = 48, 30, 3, 10


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tile_rows, out_path):
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from pathtrace_rs_b200 import parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    scene = orc.Scene("random_spheres", W, H)

    def render_rows(part):
        img = np.zeros((H, W, 3), np.float32)
        rays = 0
        for r in parallel.owned_rows(part, H):
            _, n = scene.update(SPP, DEPTH, buffer=img, rows=(int(r), int(r) + 1), nthreads=1)
            rays += n
        return torch.from_numpy(img), rays

    full, total = parallel.render_distributed(render_rows, H, W, rank, world, dist, tile_rows=tile_rows)
    if rank == 0:
        np.savez(out_path, image=full.numpy(), rays=total)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("tile_rows", [4, 7])
def test_two_rank_row_tiles_reassemble_the_single_process_image(tmp_path, tile_rows):
    out = str(tmp_path / "out.npz")
    mp.spawn(_worker, args=(2, _free_port(), tile_rows, out), nprocs=2, join=True)
    got = np.load(out)
    ref, rays = orc.Scene("random_spheres", W, H).update(SPP, DEPTH)
    assert int(got["rays"]) == rays
    assert np.array_equal(got["image"], ref)  # per-pixel seeds: the split cannot change a single bit


def _slice_worker(rank, world, port, out_path):
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from pathtrace_rs_b200 import parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    scene = orc.Scene("random_spheres", W, H)

    def render_frame(frame):
        img, n = scene.update(SPP, DEPTH, frame_num=frame, nthreads=1)  # zeroed buffer: holds col / (frame + 1)
        return torch.from_numpy(img), n

    mean, total = parallel.render_sample_slices(render_frame, rank, world, dist)
    if rank == 0:
        np.savez(out_path, image=mean.numpy(), rays=total)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sample_slices_equal_progressive_accumulation(tmp_path):
    """Sample-slice partition: rank r renders frame seed r, ONE reduce forms the mean.  Must equal the reference's own
    progressive accumulation of frames 0 and 1 into one buffer (scene.rs:86-87,113-116) up to rounding."""
    out = str(tmp_path / "slices.npz")
    mp.spawn(_slice_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    sc = orc.Scene("random_spheres", W, H)
    ref = np.zeros((H, W, 3), np.float32)
    rays = 0
    for f in range(2):
        _, n = sc.update(SPP, DEPTH, frame_num=f, buffer=ref)
        rays += n
    assert int(got["rays"]) == rays
    np.testing.assert_allclose(got["image"], ref, rtol=2e-6, atol=1e-7)
