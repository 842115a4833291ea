"""The C-ABI beyond one device and one frame at a time: glBufferSubData semantics of rtb_update, stream ordering of uploads
against frames in flight, and the multi-GPU entry points (rtb_create_multi: one process, N devices; rtb_comm_init + rtb_gather:
one process per device).  The multi-GPU cases need >= 2 visible GPUs and are skipped otherwise
(`gpurun --gpus 2 -- python -m pytest tests/test_multigpu_cabi.py -m gpu`; logs under profiles/)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import rtb200
from rtb200 import api, scenes
from rtb200.textures import TextureSet

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def _render(sc, ts, n_gpus=1, strict=1, gather=0, block_rows=4):
    w, h = int(sc.scene["canvas_width"]), int(sc.scene["canvas_height"])
    gl = rtb200.GLWrapper(w, h, n_gpus=n_gpus, block_rows=block_rows)
    gl.init_window()
    try:
        rtb200.setup_scene(gl, sc, ts)
        gl.set_option("strict", strict)
        if n_gpus > 1:
            gl.set_option("gather", gather)
        gl.draw()
        img = gl.read_pixels()
        times = gl.rank_times()
        return img, times
    finally:
        gl.stop()


def test_update_buffer_is_glBufferSubData(procedural):
    """GLWrapper.cpp:381-386: update_buffer writes a prefix and leaves the rest; init_buffer (glBufferData) replaces the block."""
    sc = scenes.synthetic_scene("mini2", 64, 40, 3)
    ts = TextureSet(cube=procedural.cube)
    gl = rtb200.GLWrapper(64, 40)
    gl.init_window()
    try:
        handles = rtb200.setup_scene(gl, sc, ts)
        gl.draw()
        want = gl.read_pixels()
        spheres = sc.array("spheres")
        gl.update_buffer(handles["spheres_buf"], spheres[:2])           # a prefix: the other spheres must survive
        gl.draw()
        assert np.array_equal(gl.read_pixels().view(np.uint32), want.view(np.uint32))
        bigger = np.empty(2 * len(spheres), dtype=spheres.dtype)        # (np.concatenate would repack the std140 layout)
        bigger[: len(spheres)] = spheres
        bigger[len(spheres):] = spheres
        with pytest.raises(rtb200.RtbError, match="exceeds the block"):
            gl.update_buffer(handles["spheres_buf"], bigger)
        moved = spheres.copy()
        moved["obj"][0][1] += 1.0
        gl.update_buffer(handles["spheres_buf"], moved)
        gl.draw()
        assert not np.array_equal(gl.read_pixels().view(np.uint32), want.view(np.uint32))
    finally:
        gl.stop()


def test_uploads_do_not_overtake_frames_on_a_caller_stream(procedural):
    """upload A, draw on a caller stream, upload B, draw again — no synchronisation in between: every frame must see its own
    scene (the context stream is ordered behind the caller's frame before the next upload touches the arrays)."""
    import torch
    ts = TextureSet(cube=procedural.cube)
    sc = scenes.build_config("spheres4k", 1 / 8)
    w, h = int(sc.scene["canvas_width"]), int(sc.scene["canvas_height"])
    cams = [(0.0, 2.0, -12.0), (3.0, 4.0, -10.0), (-4.0, 1.0, -14.0), (1.0, 6.0, -9.0)]
    gl = rtb200.GLWrapper(w, h)
    gl.init_window()
    try:
        handles = rtb200.setup_scene(gl, sc, ts)
        want = []
        for c in cams:                                                  # reference frames, one at a time
            sc.scene["camera_pos"] = c
            gl.update_buffer(handles["scene_buf"], np.ascontiguousarray(sc.scene).reshape(1))
            gl.draw()
            want.append(gl.read_pixels())
        stream = torch.cuda.Stream()
        bufs = [torch.empty((h, w, 4), dtype=torch.float32, device="cuda") for _ in cams]
        staging = np.ascontiguousarray(sc.scene).reshape(1).copy()
        for rep in range(3):
            for c, b in zip(cams, bufs):
                staging["camera_pos"] = c
                gl.update_buffer(handles["scene_buf"], staging)         # the caller's buffer is reused at once
                gl.draw_to(b.data_ptr(), stream.cuda_stream)
            stream.synchronize()
            for b, ref in zip(bufs, want):
                assert np.array_equal(b.cpu().numpy().view(np.uint32), ref.view(np.uint32))
    finally:
        gl.stop()


def test_create_multi_with_one_gpu_is_a_plain_context(procedural):
    sc = scenes.synthetic_scene("mini3", 64, 40, 3)
    a, _ = _render(sc, procedural, n_gpus=1)
    L = api.load_library()
    ctx = L.rtb_create_multi(64, 40, 1, 4)
    assert ctx and L.rtb_n_gpus(ctx) == 1
    L.rtb_destroy(ctx)
    assert not L.rtb_create_multi(64, 40, 99, 4)                       # more GPUs than the box has: refused, with a message
    assert b"out of range" in L.rtb_last_error(None)
    assert a.shape == (40, 64, 4)


@pytest.mark.skipif(_n_gpus() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("gather", [api.GATHER_NCCL, api.GATHER_P2P])
@pytest.mark.parametrize("case", ["mixed1024/16", "default1080/8", "odd"])
def test_one_process_n_gpus_is_bit_identical_to_one_gpu(case, gather, procedural):
    """rtb_create_multi: the frame is tile-partitioned over all GPUs of the box and gathered on device 0 by ONE draw()."""
    sc = {"mixed1024/16": lambda: scenes.build_config("mixed1024_4k", 1 / 16), "default1080/8": lambda: scenes.build_config("default1080", 1 / 8),
          "odd": lambda: scenes.synthetic_scene("mini4", 250, 130, 6)}[case]()           # 130 rows: the last block is partial
    n = _n_gpus()
    for strict in (1, 0):
        one, _ = _render(sc, procedural, 1, strict)
        many, (kernel_ms, frame_ms) = _render(sc, procedural, n, strict, gather)
        assert many.shape == one.shape
        assert np.array_equal(many.view(np.uint32), one.view(np.uint32)), f"{case} strict={strict} gather={gather}"
        assert len(kernel_ms) == n and all(k > 0 for k in kernel_ms) and frame_ms >= max(kernel_ms) * 0.99


@pytest.mark.skipif(_n_gpus() < 2, reason="needs >= 2 GPUs")
def test_one_process_per_gpu_gather_through_the_c_abi(tmp_path):
    """torchrun-style: N processes, rtb_comm_unique_id / rtb_comm_init / rtb_render_to + rtb_gather; rank 0 compares with one GPU."""
    n = _n_gpus()
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                          "--master-port", "29671", os.path.join(ROOT, "tests", "_cabi_gather_worker.py")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "CABI_GATHER_OK" in out.stdout


@pytest.mark.skipif(_n_gpus() < 2, reason="needs >= 2 GPUs")
def test_unchanged_main_cpp_renders_on_all_gpus(tmp_path):
    """RT_GPUS=n rt_headless: the reference's own main.cpp / SceneManager.cpp, unchanged, drive n GPUs through the same draw()."""
    exe = os.path.join(ROOT, "raytracing-opengl_b200", "host", "build", "rt_headless")
    if not os.path.isfile(exe):
        pytest.skip("rt_headless is built where /root/reference exists and travels with the snapshot")
    frames = {}
    for n in (1, _n_gpus()):
        d = tmp_path / f"gpus{n}"
        d.mkdir()
        env = dict(os.environ, RT_DUMP_DIR=str(d), RT_FRAMES="2", RT_WIDTH="320", RT_HEIGHT="180", RT_GPUS=str(n), RT_STRICT="1")
        r = subprocess.run([exe], env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        frames[n] = np.load(d / "frame_0001.npy")
    a, b = frames[1], frames[_n_gpus()]
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
