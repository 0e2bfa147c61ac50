"""rtb200 — B200-native (sm_100a) implementation of the `ray_color` bounce loop of OrientalHorizon/Ray-Tracer-Archive.

The package holds the CUDA library (csrc/ -> librtb200.so, C ABI in include/rtb200.h) and the host-side mirror of the
reference's scene-construction API (scene.py).  Importing the package does not need a GPU; creating a Context does,
and fails loudly without one — there is no CPU fallback.
"""
from . import _ffi
from ._ffi import Camera, Params, RtbError, Stats
from .renderer import Context, Scene, make_params
from .scene import *  # noqa: F401,F403  (the reference's constructor names)
from .scene import compile_scene

__all__ = ["Camera", "Params", "Stats", "RtbError", "Context", "Scene", "make_params", "compile_scene", "_ffi"]
