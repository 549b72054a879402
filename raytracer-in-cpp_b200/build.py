"""In-tree build of the CUDA library and the C++ host programs (explicit nvcc command lines).

    librt_b200.so   csrc/rt_api.cu + host/{obj_loader,bvh_builder}.cpp   -> the C ABI (include/rt_api.h)
    rt_cli          host/rt_cli.cpp + host/flyscene.cpp                   -> headless `main.cpp` replacement

Device code is compiled for sm_100a only, with -fmad=false (see csrc/rt_device.cuh: the parity-
critical arithmetic must not be contracted into FMAs) and -lineinfo for ncu source correlation.
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "librt_b200.so")
CLI = os.path.join(HERE, "rt_cli")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
              "-Xcompiler", "-fPIC,-ffp-contract=off,-O2"]

LIB_SOURCES = ["csrc/rt_api.cu", "host/obj_loader.cpp", "host/bvh_builder.cpp", "host/ref_octree.cpp"]
LIB_DEPS = LIB_SOURCES + ["csrc/rt_device.cuh", "csrc/rt_kernels.cuh", "csrc/rt_frame.cuh", "csrc/rt_multi.inl", "csrc/rt_build.cuh", "csrc/rt_gpu_build.inl", "host/obj_loader.hpp", "host/bvh_builder.hpp",
                          "host/vec3.hpp", "host/ref_octree.hpp", "../include/rt_api.h"]
CLI_SOURCES = ["host/rt_cli.cpp", "host/flyscene.cpp"]
CLI_DEPS = CLI_SOURCES + ["host/flyscene.hpp", "host/vec3.hpp", "../include/rt_api.h"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(os.path.join(HERE, d)) > t for d in deps if os.path.exists(os.path.join(HERE, d)))


def build_lib(force: bool = False, verbose: bool = False) -> str:
    if force or _stale(LIB, LIB_DEPS):
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-shared", "-o", LIB] + LIB_SOURCES
        subprocess.check_call(cmd, cwd=HERE)
    return LIB


def build_cli(force: bool = False) -> str:
    if not all(os.path.exists(os.path.join(HERE, s)) for s in CLI_SOURCES):
        return ""
    if force or _stale(CLI, CLI_DEPS) or _stale(CLI, ["librt_b200.so"]):
        cmd = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-Wall", "-o", CLI] + CLI_SOURCES + \
              ["-L" + HERE, "-lrt_b200", "-Wl,-rpath,$ORIGIN"]
        subprocess.check_call(cmd, cwd=HERE)
    return CLI


PROBE = os.path.join(HERE, "facade_probe")


def build_probe(force: bool = False) -> str:
    """Small C++ program that calls every render-path member of the facade once (used by the GPU tests)."""
    srcs = ["host/facade_probe.cpp", "host/flyscene.cpp"]
    if force or _stale(PROBE, srcs + ["host/flyscene.hpp", "librt_b200.so"]):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-Wall", "-o", PROBE] + srcs +
                              ["-L" + HERE, "-lrt_b200", "-Wl,-rpath,$ORIGIN"], cwd=HERE)
    return PROBE


def build_all(force: bool = False) -> None:
    build_lib(force)
    build_cli(force)
    build_probe(force)
