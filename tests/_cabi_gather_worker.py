"""Worker of tests/test_multigpu_cabi.py::test_one_process_per_gpu_gather_through_the_c_abi (launched by torchrun, one rank per GPU)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rtb200  # noqa: E402
from rtb200 import scenes, textures  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")                       # the launcher's plumbing only ships the 128-byte NCCL id; the data path is rtb_gather
ts = textures.procedural_textures()
sc = scenes.synthetic_scene("mini4", 250, 130, 6)
w, h = 250, 130
gl = rtb200.GLWrapper(w, h, device=local)
gl.init_window()
ids = [gl.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
gl.comm_init(ids[0], rank, world, 4)
rtb200.setup_scene(gl, sc, ts)
stream = torch.cuda.Stream()
localbuf = torch.zeros((max(gl.local_rows(), 1), w, 4), dtype=torch.float32, device="cuda")
full = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda") if rank == 0 else None
ok = True
for strict in (1, 0):
    gl.set_option("strict", strict)
    for _ in range(2):
        gl.draw_to(localbuf.data_ptr(), stream.cuda_stream)
        gl.gather(localbuf.data_ptr(), full.data_ptr() if rank == 0 else 0, stream.cuda_stream)
    stream.synchronize()
    if rank == 0:
        one = rtb200.GLWrapper(w, h, device=local)
        one.init_window()
        rtb200.setup_scene(one, sc, ts)
        one.set_option("strict", strict)
        one.draw()
        ok = ok and np.array_equal(one.read_pixels().view(np.uint32), full.cpu().numpy().view(np.uint32))
        one.stop()
dist.barrier()
gl.stop()
if rank == 0:
    print("CABI_GATHER_OK" if ok else "CABI_GATHER_MISMATCH", flush=True)
dist.destroy_process_group()
