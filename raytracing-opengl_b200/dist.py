"""Multi-GPU plumbing: one process per GPU, frame partitioned by interleaved row blocks, ONE gather per frame.

The path shards naturally (pixels are independent; 2x2 derivative quads never straddle a block because
block_rows is a multiple of 4), so no data-path collective exists except the frame gather at the end
(SURVEY.md 8e).  torch.distributed is used for rendezvous and the gather (NCCL over NVLink on GPUs,
gloo on CPU in the tests); nothing here touches pixel values.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

BLOCK_ROWS = 16


def local_row_map(height: int, rank: int, world: int, block_rows: int = BLOCK_ROWS) -> torch.Tensor:
    """Canvas scanline of every local scanline of `rank` (same rule as rtb_set_partition)."""
    rows = []
    b = rank
    while b * block_rows < height:
        rows.extend(range(b * block_rows, min(height, (b + 1) * block_rows)))
        b += world
    return torch.tensor(rows, dtype=torch.long)


_ROW_MAPS = {}


def _device_row_map(height, rank, world, block_rows, device):
    key = (height, rank, world, block_rows, str(device))
    if key not in _ROW_MAPS:
        _ROW_MAPS[key] = local_row_map(height, rank, world, block_rows).to(device)
    return _ROW_MAPS[key]


def max_local_rows(height: int, world: int, block_rows: int = BLOCK_ROWS) -> int:
    return max(len(local_row_map(height, r, world, block_rows)) for r in range(world))


def init_from_env():
    """(rank, local_rank, world).  Initialises the default process group when launched by torchrun."""
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    return rank, local_rank, world


def gather_frame(local: torch.Tensor, height: int, rank: int, world: int, block_rows: int = BLOCK_ROWS, out: torch.Tensor | None = None,
                 scratch: list | None = None):
    """Gather every rank's packed scanlines [max_local_rows, W, 4] on rank 0 and de-interleave into [H, W, 4].

    `local` must already be padded to max_local_rows (equal sizes on all ranks, as the collective requires).
    Returns the full frame on rank 0, None elsewhere."""
    if world == 1:
        return local[:height]
    if rank == 0:
        if scratch is None:
            scratch = [torch.empty_like(local) for _ in range(world)]
        dist.gather(local, gather_list=scratch, dst=0)
        if out is None:
            out = torch.empty((height,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        for r in range(world):
            rows = _device_row_map(height, r, world, block_rows, local.device)      # cached: no per-frame host work
            out.index_copy_(0, rows, scratch[r][: len(rows)])
        return out
    dist.gather(local, gather_list=None, dst=0)
    return None
