"""N>1 host logic on CPU: two ranks over gloo."""
import os
import socket
import subprocess
import sys

import torch

from rtb200 import dist as rdist

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_partition_and_gather_reassemble_the_frame():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(HERE, "_gloo_worker.py")]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "GLOO_OK 2" in r.stdout


def test_row_maps_tile_the_canvas_exactly_once():
    for h, world, br in ((2160, 8, 16), (4320, 8, 16), (100, 3, 16), (36, 4, 4), (16, 2, 16)):
        seen = torch.cat([rdist.local_row_map(h, r, world, br) for r in range(world)])
        assert sorted(seen.tolist()) == list(range(h))
        assert rdist.max_local_rows(h, world, br) >= (h + world - 1) // world
