"""raytracing-opengl_b200 — B200-native ray-trace pass behind the raytracing-opengl scene API.

The directory name contains a hyphen (it is the reference's name); import it with
`importlib.import_module("raytracing-opengl_b200")` or through the `rtb200`
alias module at the repository root.
"""
from . import scene, scenes, textures  # noqa: F401
from .scene import SceneContainer, SceneManager, SurfaceFactory  # noqa: F401
