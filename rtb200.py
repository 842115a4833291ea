"""Import alias: `import rtb200` == the package in ./raytracing-opengl_b200/ (a hyphen is not importable by name)."""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
_pkg = importlib.import_module("raytracing-opengl_b200")
sys.modules[__name__] = _pkg
