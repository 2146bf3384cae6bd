"""Multi-GPU orchestration of one Scene::update: one process per GPU, interleaved row tiles, no data-path
collective.  torch.distributed is plumbing only: a barrier, the ray-count sum and — when a single
image is wanted on rank 0 — a gather of the rows each rank owns.

The per-rank renderer is injected (`render_rows`) so the same partition/gather logic runs under gloo on CPU in
the tests and under NCCL on B200s in bench.py.
"""
import numpy as np

from . import ffi

DEFAULT_TILE_ROWS = 4


def partition_for(rank, world_size, tile_rows=DEFAULT_TILE_ROWS):
    return ffi.PtPartition(tile_rows, rank, world_size, 0)


def owned_rows(part, height):
    """Row indices `part` owns, ascending — pt_partition_rows of the C ABI (host logic only, no GPU)."""
    import ctypes as C
    L = ffi.libptgpu()
    n = L.pt_partition_rows(C.byref(part), height, None, 0)
    rows = np.zeros(n, np.uint32)
    if n:
        L.pt_partition_rows(C.byref(part), height, rows.ctypes.data_as(C.c_void_p), n)
    return rows


def gather_image(local_image, height, width, rank, world_size, tile_rows=DEFAULT_TILE_ROWS, dist=None, device=None):
    """Assemble the full image on rank 0 from per-rank images that hold valid data only in their owned rows.

    local_image: torch tensor [height, width, 3] (device for NCCL, CPU for gloo).  Returns the full tensor on rank 0,
    None elsewhere.  Row-tile mode needs no reduction: rows are disjoint, so this is a pure gather."""
    import torch
    if world_size == 1:
        return local_image
    counts = [len(owned_rows(partition_for(r, world_size, tile_rows), height)) for r in range(world_size)]
    max_rows = max(counts)
    mine = torch.from_numpy(owned_rows(partition_for(rank, world_size, tile_rows), height).astype(np.int64)).to(local_image.device)
    send = torch.zeros((max_rows, width, 3), dtype=local_image.dtype, device=local_image.device)
    send[: len(mine)] = local_image.index_select(0, mine)
    if rank == 0:
        bufs = [torch.empty_like(send) for _ in range(world_size)]
        dist.gather(send, bufs, dst=0)
        full = torch.empty_like(local_image)
        for r in range(world_size):
            rows = torch.from_numpy(owned_rows(partition_for(r, world_size, tile_rows), height).astype(np.int64)).to(full.device)
            full.index_copy_(0, rows, bufs[r][: counts[r]])
        return full
    dist.gather(send, None, dst=0)
    return None


def render_distributed(render_rows, height, width, rank, world_size, dist=None, tile_rows=DEFAULT_TILE_ROWS, gather=True):
    """render_rows(part) -> (torch tensor [height,width,3] with the owned rows filled, ray_count).
    Returns (full image on rank 0 or None, total ray count on every rank)."""
    import torch
    part = partition_for(rank, world_size, tile_rows)
    local, rays = render_rows(part)
    total = torch.tensor([rays], dtype=torch.int64, device=local.device)
    if world_size > 1:
        dist.all_reduce(total)
    full = gather_image(local, height, width, rank, world_size, tile_rows, dist) if gather else None
    return full, int(total.item())


def combine_sample_slices(local_image, frame_num, world_size, dist=None):
    """Sample-slice mode: rank r rendered the WHOLE image with frame seed `frame_num` = r into a zeroed
    buffer, so it holds col_r / (r + 1) (the blend of scene.rs:86-87,113-116 against an empty previous frame).  Undo
    that weight, sum over ranks with ONE reduce and scale by 1/G: the equal-weight mean of G frames — exactly what the
    reference's progressive accumulation over frames 0..G-1 converges to (running mean algebra of scene.rs:86-87).
    In place; the result is valid on rank 0."""
    local_image.mul_(float(frame_num + 1))
    if world_size > 1:
        dist.reduce(local_image, dst=0)
        local_image.mul_(1.0 / world_size)
    return local_image


def render_sample_slices(render_frame, rank, world_size, dist=None):
    """render_frame(frame_num) -> (torch tensor [h,w,3] rendered into a zeroed buffer, ray_count).
    Returns (mean image on rank 0, total ray count on every rank)."""
    import torch
    local, rays = render_frame(rank)
    total = torch.tensor([rays], dtype=torch.int64, device=local.device)
    if world_size > 1:
        dist.all_reduce(total)
    return combine_sample_slices(local, rank, world_size, dist), int(total.item())
