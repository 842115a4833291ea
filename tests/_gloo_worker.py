"""world_size-2 worker (gloo, CPU): row-block partition + one gather reassemble the frame exactly.
Pixel data comes from the CPU oracle here (no GPU in this container); the partition/gather code under test is
raytracing-opengl_b200/dist.py, the same code bench.py runs over NCCL."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rtb200  # noqa: E402
from rtb200 import dist as rdist, scenes  # noqa: E402
from oracle.binding import Oracle  # noqa: E402


def main():
    rank, _, world = rdist.init_from_env()
    w, h, br = 48, 52, 16                       # 52 rows: the last block is partial (4 rows)
    sc = scenes.synthetic_scene("mini2", w, h, 3)
    o = Oracle(sc)
    rows = rdist.local_row_map(h, rank, world, br)
    pad = rdist.max_local_rows(h, world, br)
    local = torch.zeros((pad, w, 4), dtype=torch.float32)
    b = rank
    cur = 0
    while b * br < h:                            # render only the blocks this rank owns
        n = min(br, h - b * br)
        local[cur:cur + n] = torch.from_numpy(o.render(0, b * br, w, n, threads=1))
        cur += n
        b += world
    assert cur == len(rows)
    full = rdist.gather_frame(local, h, rank, world, br)
    if rank == 0:
        want = torch.from_numpy(o.render(threads=2))
        assert full.shape == want.shape
        assert torch.equal(full, want), float((full - want).abs().max())
        # the numpy re-interleave used by single-process callers agrees with the collective path
        print("GLOO_OK", world, flush=True)
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
