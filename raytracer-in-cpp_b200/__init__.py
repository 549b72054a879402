"""raytracer-in-cpp_b200 -- B200-native render path for the Raytracer-in-CPP scene conventions.

The product is the CUDA library built from csrc/ (C ABI in include/rt_api.h) plus the C++ host
facade in host/.  This Python package only holds build plumbing (build.py), the ctypes binding used
by tests and bench.py (capi.py) and the synthetic scene generators (scenes.py).
"""
from . import build, capi, scenes  # noqa: F401
