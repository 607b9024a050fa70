"""B200-native path-tracing core for Smaug123/ray-tracing-fsharp: host-side mirror of the F# surface.

The compute path lives in csrc/ (CUDA, sm_100a) behind the C ABI of include/rtfs_b200.h and is
loaded by `native.py`; there is no CPU fallback — calls fail loudly if the library or a GPU is missing.
"""
from . import abi  # noqa: F401
from .domain import (Colour, Hittable, InfinitePlane, InfinitePlaneStyle, ParameterisedTexture, Pixel,  # noqa: F401
                     PlaneMapInverse, Sphere, SphereStyle, Texture, marshal)
