"""Python host side of the drop-in boundary: `GLWrapper`, method for method.

Mirrors the public interface of the reference's render driver
(src/GLWrapper.h:17-38) on top of the C-ABI in include/rtb200.h — the same
library the C++ host (host/GLWrapper.cpp) calls.  The reference's error
behaviour is "print and exit" (utils.h:25,62; GLWrapper.cpp:371-375); here every
failing call raises `RtbError` carrying rtb_last_error().

There is no CPU path: if librtb200.so is missing, or no sm_100 device is
present, construction / init_window() raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .scene import UBO_BLOCKS, rt_defines

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RTB200_LIB") or os.path.join(_HERE, "librtb200.so")     # RTB200_LIB: development builds only

KERNEL_AUTO, KERNEL_QUAD, KERNEL_PERSISTENT = 0, 1, 2
GATHER_NCCL, GATHER_P2P = 0, 1
SMAA_LOW, SMAA_MEDIUM, SMAA_HIGH, SMAA_ULTRA = 0, 1, 2, 3        # enum SMAA_PRESET, SMAA_Builder.h:9-12
COMM_ID_BYTES = 128


class RtbError(RuntimeError):
    pass


class RtbStats(C.Structure):
    _fields_ = [("pixels", C.c_uint64), ("rays_nearest", C.c_uint64), ("rays_shadow", C.c_uint64), ("tests", C.c_uint64 * 7),
                ("dk_iterations", C.c_uint64), ("shaded_hits", C.c_uint64 * 7), ("light_evals", C.c_uint64),
                ("flops", C.c_double), ("kernel_ms", C.c_float), ("kernel_used", C.c_int32), ("grid", C.c_int32),
                ("block", C.c_int32), ("smem_bytes", C.c_int32)]

    def as_dict(self):
        return {"pixels": self.pixels, "rays_nearest": self.rays_nearest, "rays_shadow": self.rays_shadow,
                "tests": list(self.tests), "dk_iterations": self.dk_iterations, "shaded_hits": list(self.shaded_hits),
                "light_evals": self.light_evals, "flops": self.flops, "kernel_ms": self.kernel_ms,
                "kernel_used": self.kernel_used, "grid": self.grid, "block": self.block, "smem_bytes": self.smem_bytes}

    @property
    def rays(self):
        return self.rays_nearest + self.rays_shadow


_lib = None


def load_library() -> C.CDLL:
    """Load librtb200.so (built in-tree by __graft_entry__.build() / make -C csrc).  Fails loudly when absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RtbError(f"{LIB_PATH} is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback for the ray-trace pass)")
    L = C.CDLL(LIB_PATH)
    vp, i, sz = C.c_void_p, C.c_int, C.c_size_t
    L.rtb_create.restype = vp
    L.rtb_create.argtypes = [i, i, i]
    L.rtb_destroy.argtypes = [vp]
    L.rtb_set_partition.argtypes = [vp, i, i, i]
    L.rtb_local_rows.argtypes = [vp]
    L.rtb_set_defines.argtypes = [vp, vp]
    L.rtb_upload.argtypes = [vp, i, vp, sz]
    L.rtb_update.argtypes = [vp, i, vp, sz]
    L.rtb_create_multi.restype = vp
    L.rtb_create_multi.argtypes = [i, i, i, i]
    L.rtb_n_gpus.argtypes = [vp]
    L.rtb_rank_times.argtypes = [vp, vp, C.POINTER(C.c_float)]
    L.rtb_comm_unique_id.argtypes = [vp]
    L.rtb_comm_init.argtypes = [vp, vp, i, i, i]
    L.rtb_gather.argtypes = [vp, vp, vp, vp]
    L.rtb_enable_smaa.argtypes = [vp, i]
    L.rtb_smaa_set_tables.argtypes = [vp, vp, vp]
    L.rtb_smaa_apply.argtypes = [vp, vp, vp, vp, vp, C.POINTER(C.c_float)]
    L.rtb_smaa_last_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.rtb_set_cubemap.argtypes = [vp, C.POINTER(vp), i, i, i]
    L.rtb_set_texture2d.argtypes = [vp, i, vp, i, i, i]
    L.rtb_set_option.argtypes = [vp, C.c_char_p, i]
    L.rtb_render.argtypes = [vp]
    L.rtb_render_to.argtypes = [vp, vp, vp]
    L.rtb_render_counted.argtypes = [vp, C.POINTER(RtbStats)]
    L.rtb_sync.argtypes = [vp]
    L.rtb_read_rgba32f.argtypes = [vp, vp]
    L.rtb_read_rgba8.argtypes = [vp, vp]
    L.rtb_device_framebuffer.restype = vp
    L.rtb_device_framebuffer.argtypes = [vp]
    L.rtb_get_stats.argtypes = [vp, C.POINTER(RtbStats)]
    L.rtb_tile_order.argtypes = [vp, vp, vp, i]
    L.rtb_measure_fp32_peak.argtypes = [i, C.POINTER(C.c_double)]
    L.rtb_measure_fp32_peak3.argtypes = [i, C.POINTER(C.c_double)]
    L.rtb_last_error.restype = C.c_char_p
    L.rtb_last_error.argtypes = [vp]
    L.rtb_version.restype = C.c_char_p
    _lib = L
    return L


def measure_fp32_peak(device: int = 0, three_registers: bool = False) -> float:
    """FFMA-only microbenchmark, TFLOP/s.  three_registers: every FFMA reads three distinct registers (the operand-delivery
    limit) instead of one register + a uniform register + a reused operand (the pipe's peak)."""
    L = load_library()
    out = C.c_double(0)
    fn = L.rtb_measure_fp32_peak3 if three_registers else L.rtb_measure_fp32_peak
    if fn(device, C.byref(out)) != 0:
        raise RtbError("rtb_measure_fp32_peak failed (no CUDA device?)")
    return out.value


class GLWrapper:
    """src/GLWrapper.h, same method names; `draw()` launches the CUDA ray-trace pass."""

    def __init__(self, width: int, height: int, fullScreen: bool = False, device: int = 0, n_gpus: int = 1, block_rows: int = 4):
        """n_gpus > 1: ONE process drives devices 0..n_gpus-1 (rtb_create_multi); every method then fans out inside the
        library and draw() ends with the frame gathered on device 0 (what the C++ host does under RT_GPUS=n)."""
        self._L = load_library()
        self.width, self.height = int(width), int(height)
        self.device = device
        self.n_gpus, self.block_rows = int(n_gpus), int(block_rows)
        self._ctx = None
        self._ubos = {}                 # handle -> binding (update_buffer is static and only gets the handle)
        self._next_handle = 1
        self._textures = {}             # handle -> unit
        self._skybox = None
        self.window = self              # GLFWwindow* stand-in

    # -- GLWrapper.h:21-23
    def getWidth(self):
        return self.width

    def getHeight(self):
        return self.height

    def getProgramId(self):
        return 1

    def _check(self, rc):
        if rc != 0:
            raise RtbError(self._L.rtb_last_error(self._ctx).decode() or f"rtb error {rc}")

    # -- GLWrapper.h:25  (GLWrapper.cpp:61-133: create context)
    def init_window(self) -> bool:
        if self.n_gpus > 1:
            self._ctx = self._L.rtb_create_multi(self.width, self.height, self.n_gpus, self.block_rows)
        else:
            self._ctx = self._L.rtb_create(self.width, self.height, self.device)
        if not self._ctx:
            raise RtbError(self._L.rtb_last_error(None).decode())
        return True

    # -- GLWrapper.h:30  (GLWrapper.cpp:149-153).  preset: SMAA_PRESET 0..3 = LOW, MEDIUM, HIGH, ULTRA; None / -1 switches it off.
    #    The lookup tables (src/AreaTex.h, src/SearchTex.h) must have been handed over with smaa_set_tables().
    def enable_SMAA(self, preset=SMAA_ULTRA):
        self._check(self._L.rtb_enable_smaa(self._ctx, -1 if preset is None else int(preset)))

    # -- SMAA_Builder::load_area_texture / load_search_texture (SMAA_Builder.h:45-79): RG8 [560,160,2] and R8 [16,64]
    def smaa_set_tables(self, area, search):
        a = np.ascontiguousarray(area, dtype=np.uint8)
        s = np.ascontiguousarray(search, dtype=np.uint8)
        if a.size != 160 * 560 * 2 or s.size != 64 * 16:
            raise RtbError("AreaTex is 160x560 RG8, SearchTex 64x16 R8")
        self._check(self._L.rtb_smaa_set_tables(self._ctx, a.ctypes.data, s.ctypes.data))

    def smaa_apply(self, rgba8):
        """The three SMAA passes alone on an RGBA8 image [H, W, 4] of the context's size: (out, edges [H,W,2], blend [H,W,4], ms)."""
        img = np.ascontiguousarray(rgba8, dtype=np.uint8)
        if img.shape != (self.height, self.width, 4):
            raise RtbError(f"image must be [{self.height}, {self.width}, 4]")
        out = np.empty_like(img)
        edges = np.empty((self.height, self.width, 2), dtype=np.uint8)
        blend = np.empty_like(img)
        ms = C.c_float(0)
        self._check(self._L.rtb_smaa_apply(self._ctx, img.ctypes.data, out.ctypes.data, edges.ctypes.data, blend.ctypes.data, C.byref(ms)))
        return out, edges, blend, float(ms.value)

    def smaa_last_ms(self) -> float:
        ms = C.c_float(0)
        self._check(self._L.rtb_smaa_last_ms(self._ctx, C.byref(ms)))
        return float(ms.value)

    # -- GLWrapper.h:26  (GLWrapper.cpp:232-247)
    def init_shaders(self, defines):
        d = np.ascontiguousarray(np.asarray(defines, dtype=rt_defines)).reshape(1)
        self._check(self._L.rtb_set_defines(self._ctx, d.ctypes.data))

    # -- GLWrapper.h:35,27  (GLWrapper.cpp:284-317, 135-141).  `faces`: 6 decoded uint8 arrays [h,w,ch]
    def load_cubemap(self, faces, genMipmap: bool = False):
        faces = [np.ascontiguousarray(f, dtype=np.uint8) for f in faces]
        if len(faces) != 6:
            raise RtbError("a cubemap has six faces")
        h, w, ch = faces[0].shape
        arr = (C.c_void_p * 6)(*[f.ctypes.data for f in faces])
        self._check(self._L.rtb_set_cubemap(self._ctx, arr, w, h, ch))
        self._skybox = self._next_handle
        self._next_handle += 1
        return self._skybox

    def set_skybox(self, textureId):
        self._skybox = textureId

    # -- GLWrapper.h:36  (GLWrapper.cpp:356-363).  `pixels`: decoded uint8 [h,w,ch]; texNum = texture unit 1..5
    def load_texture(self, texNum: int, pixels, uniformName: str = "", wrapMode=None):
        a = np.ascontiguousarray(pixels, dtype=np.uint8)
        h, w, ch = a.shape
        self._check(self._L.rtb_set_texture2d(self._ctx, texNum, a.ctypes.data, w, h, ch))
        handle = self._next_handle
        self._next_handle += 1
        self._textures[handle] = texNum
        return handle

    # -- GLWrapper.h:37  (GLWrapper.cpp:365-379).  Returns the ubo handle (the reference writes it through GLuint*).
    def init_buffer(self, name: str, bindingPoint: int, data) -> int:
        if name not in UBO_BLOCKS:
            raise RtbError(f"Invalid ubo block name '{name}'")           # GLWrapper.cpp:371-375
        if UBO_BLOCKS[name][0] != bindingPoint:
            raise RtbError(f"block {name} is bound at {UBO_BLOCKS[name][0]}, not {bindingPoint}")
        handle = self._next_handle
        self._next_handle += 1
        self._ubos[handle] = bindingPoint
        if data is None:                                                  # glBufferData(size, NULL): allocate, contents follow
            self._check(self._L.rtb_upload(self._ctx, bindingPoint, None, UBO_BLOCKS[name][1].itemsize if bindingPoint == 0 else 0))
        else:
            a = np.ascontiguousarray(data)
            self._check(self._L.rtb_upload(self._ctx, bindingPoint, a.ctypes.data if a.nbytes else None, a.nbytes))
        return handle

    # -- GLWrapper.h:38  (GLWrapper.cpp:381-386: glBufferSubData(0, size) — the block keeps its size)
    def update_buffer(self, ubo: int, data):
        a = np.ascontiguousarray(data)
        if a.nbytes:
            self._check(self._L.rtb_update(self._ctx, self._ubos[ubo], a.ctypes.data, a.nbytes))

    # -- GLWrapper.h:34  (GLWrapper.cpp:155-165)
    def draw(self):
        self._check(self._L.rtb_render(self._ctx))

    # -- GLWrapper.h:29
    def stop(self):
        if self._ctx:
            self._L.rtb_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.stop()
        except Exception:
            pass

    # ---- beyond the reference: read-back, options, multi-GPU partition, statistics ----
    def set_option(self, key: str, value: int):
        self._check(self._L.rtb_set_option(self._ctx, key.encode(), int(value)))

    def set_partition(self, rank: int, world: int, block_rows: int = 16):
        self._check(self._L.rtb_set_partition(self._ctx, rank, world, block_rows))

    def local_rows(self) -> int:
        return self._L.rtb_local_rows(self._ctx)

    def sync(self):
        self._check(self._L.rtb_sync(self._ctx))

    def _frame_rows(self) -> int:
        return self.height if self.n_gpus > 1 else self.local_rows()

    def read_pixels(self) -> np.ndarray:
        """RGBA32F [rows, W, 4], row 0 = bottom scanline: this rank's packed blocks, or the whole gathered frame of a multi-GPU context."""
        out = np.empty((self._frame_rows(), self.width, 4), dtype=np.float32)
        self._check(self._L.rtb_read_rgba32f(self._ctx, out.ctypes.data))
        return out

    def read_pixels_u8(self) -> np.ndarray:
        out = np.empty((self._frame_rows(), self.width, 4), dtype=np.uint8)
        self._check(self._L.rtb_read_rgba8(self._ctx, out.ctypes.data))
        return out

    def draw_to(self, device_ptr: int, stream: int = 0):
        """Render into a caller-owned device buffer (e.g. a torch tensor's data_ptr) on a caller stream."""
        self._check(self._L.rtb_render_to(self._ctx, device_ptr, stream or None))

    def tile_order(self):
        """(cost, order) of the cost-ordered tile hand-out (option "lpt") after the last frame, or None while there is none."""
        n = self._L.rtb_tile_order(self._ctx, None, None, 0)
        if n < 0:
            self._check(n)
        if n == 0:
            return None
        cost, order = np.empty(n, np.uint32), np.empty(n, np.uint32)
        self._check(min(0, self._L.rtb_tile_order(self._ctx, cost.ctypes.data, order.ctypes.data, n)))
        return cost, order

    def draw_counted(self) -> RtbStats:
        st = RtbStats()
        self._check(self._L.rtb_render_counted(self._ctx, C.byref(st)))
        return st

    def stats(self) -> RtbStats:
        st = RtbStats()
        self._check(self._L.rtb_get_stats(self._ctx, C.byref(st)))
        return st

    def device_framebuffer(self) -> int:
        return self._L.rtb_device_framebuffer(self._ctx)

    # ---- multi-GPU behind the C-ABI ----
    def rank_times(self):
        """(kernel ms of every rank, device-side frame ms incl. the gather) of the last draw()."""
        k = (C.c_float * max(1, self._L.rtb_n_gpus(self._ctx)))()
        f = C.c_float(0)
        self._check(self._L.rtb_rank_times(self._ctx, k, C.byref(f)))
        return [float(x) for x in k], float(f.value)

    def comm_unique_id(self) -> bytes:
        """rank 0 of a one-process-per-GPU job: the NCCL id the other ranks need (ship it with the launcher's own means)."""
        buf = (C.c_uint8 * COMM_ID_BYTES)()
        rc = self._L.rtb_comm_unique_id(buf)
        if rc != 0:
            raise RtbError(self._L.rtb_last_error(None).decode())
        return bytes(buf)

    def comm_init(self, comm_id: bytes, rank: int, world: int, block_rows: int = 4):
        buf = (C.c_uint8 * COMM_ID_BYTES).from_buffer_copy(comm_id)
        self._check(self._L.rtb_comm_init(self._ctx, buf, rank, world, block_rows))

    def gather(self, local_ptr: int = 0, full_ptr: int = 0, stream: int = 0):
        """the frame-end NCCL gather of a one-process-per-GPU job (rtb_gather): device pointers, ordered on `stream`"""
        self._check(self._L.rtb_gather(self._ctx, local_ptr or None, full_ptr or None, stream or None))


def setup_scene(gl: GLWrapper, scene, textures=None):
    """What main.cpp + SceneManager::init do with a scene_container: init_shaders, samplers, init_buffers
    (main.cpp:134-156, SceneManager.cpp:244-255) followed by one update_buffers (SceneManager.cpp:266-276)."""
    gl.init_shaders(scene.get_defines())
    if textures is not None:
        if textures.cube is not None:
            gl.set_skybox(gl.load_cubemap(textures.cube, False))
        for unit, px in sorted(textures.tex2d.items()):
            gl.load_texture(unit, px, "")
    handles = {}
    handles["scene_buf"] = gl.init_buffer("scene_buf", 0, None)
    for name, attr in (("spheres_buf", "spheres"), ("planes_buf", "planes"), ("surfaces_buf", "surfaces"), ("boxes_buf", "boxes"),
                       ("toruses_buf", "toruses"), ("rings_buf", "rings"), ("lights_point_buf", "lights_point"),
                       ("lights_direct_buf", "lights_direct")):
        handles[name] = gl.init_buffer(name, UBO_BLOCKS[name][0], scene.array(attr))
    gl.update_buffer(handles["scene_buf"], np.ascontiguousarray(scene.scene).reshape(1))
    return handles


def update_buffers(gl: GLWrapper, scene, handles):
    """SceneManager::update_buffers (SceneManager.cpp:266-276), what the frame loop calls every frame: the scene uniform and the
    seven primitive / point-light arrays, empty ones skipped (:257-264).  lights_direct_buf is NOT among them: the reference
    writes it once, in init_buffers (:254) — a host that re-sent it would change what an animated scene renders."""
    gl.update_buffer(handles["scene_buf"], np.ascontiguousarray(scene.scene).reshape(1))
    for name, attr in (("spheres_buf", "spheres"), ("planes_buf", "planes"), ("surfaces_buf", "surfaces"), ("boxes_buf", "boxes"),
                       ("toruses_buf", "toruses"), ("rings_buf", "rings"), ("lights_point_buf", "lights_point")):
        a = scene.array(attr)
        if len(a):
            gl.update_buffer(handles[name], a)


def gather_rows(parts, height: int, world: int, block_rows: int = 16) -> np.ndarray:
    """Re-interleave per-rank packed scanlines (list indexed by rank, each [local_rows, W, 4]) into the full frame."""
    w = parts[0].shape[1]
    out = np.empty((height, w, 4), dtype=parts[0].dtype)
    cursor = [0] * world
    b = 0
    while b * block_rows < height:
        r = b % world
        n = min(block_rows, height - b * block_rows)
        out[b * block_rows:b * block_rows + n] = parts[r][cursor[r]:cursor[r] + n]
        cursor[r] += n
        b += 1
    return out
