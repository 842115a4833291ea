"""raytracing-opengl_b200 — B200-native ray-trace pass behind the raytracing-opengl scene API.

The directory name contains a hyphen (it is the reference's name); import it with
`importlib.import_module("raytracing-opengl_b200")` or through the `rtb200`
alias module at the repository root.

  scene     rt_* structs + SceneManager / SurfaceFactory factories (src/scene.h, SceneManager.cpp, Surface.h)
  scenes    the reference's default scene and the synthetic BASELINE.json scenes
  textures  decoded sampler inputs
  dist      row-block partition + the single frame gather (torch.distributed: NCCL / gloo)
  api       GLWrapper (src/GLWrapper.h) over the C-ABI of librtb200.so (CUDA, sm_100a; no CPU path)
"""
from . import api, scene, scenes, textures  # noqa: F401
from .api import GLWrapper, RtbError, gather_rows, measure_fp32_peak, setup_scene, update_buffers  # noqa: F401
from .scene import SceneContainer, SceneManager, SurfaceFactory  # noqa: F401
