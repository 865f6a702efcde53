"""In-tree build of libb200rt.so for sm_100a (nvcc cross-compiles without a GPU).

    python homework-18-graphics-raytracer_b200/build.py [--force] [--verbose]

The kernels are compiled with -fmad=false: every float expression rounds like the reference's
non-fused f32 code; the fused arithmetic of the cast filter is written with explicit FMA intrinsics.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_DIR = os.path.join(PKG_DIR, "_lib")
OBJ_DIR = os.path.join(PKG_DIR, "_lib", "obj")
LIB_PATH = os.path.join(LIB_DIR, "libb200rt.so")
INCLUDE = os.path.join(ROOT, "include")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_COMMON = ["-O3", "-lineinfo", "-std=c++17", "-I", INCLUDE, "-I", CSRC]
if os.environ.get("B200RT_DIST_THREADS"):        # tuning knob: CTA size of the lockstep (distributed) tracer
    NVCC_COMMON.append("-DB200RT_DIST_THREADS=" + os.environ["B200RT_DIST_THREADS"])
if os.environ.get("B200RT_TRACE_MIN_BLOCKS"):   # tuning knob: resident CTAs per SM the tracer is compiled for
    NVCC_COMMON.append("-DB200RT_TRACE_MIN_BLOCKS=" + os.environ["B200RT_TRACE_MIN_BLOCKS"])


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libb200rt.so cannot be built (there is no CPU fallback)")


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd: list[str], verbose: bool) -> None:
    if verbose:
        print("+", " ".join(cmd), flush=True)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("build step failed: " + " ".join(cmd))
    if verbose and (res.stdout or res.stderr):
        print(res.stdout + res.stderr)


def build_all(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = _nvcc()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(INCLUDE, "b200rt.h"))
    headers.append(os.path.join(INCLUDE, "b200rt_dev.h"))
    headers.append(os.path.abspath(__file__))
    objs = []

    def obj(name: str) -> str:
        return os.path.join(OBJ_DIR, name + ".o")

    # kernels: no implicit FMA contraction
    src = os.path.join(CSRC, "rt_kernels.cu")
    o = obj("rt_kernels")
    if force or _newer(o, [src] + headers):
        _run([nvcc, *ARCH, *NVCC_COMMON, "-fmad=false", "-Xptxas", "-v", "-Xcompiler", "-fPIC", "-c", src, "-o", o],
             verbose)
    objs.append(o)

    src = os.path.join(CSRC, "rt_wavefront.cu")
    o = obj("rt_wavefront")
    if force or _newer(o, [src] + headers):
        _run([nvcc, *ARCH, *NVCC_COMMON, "-fmad=false", "-Xptxas", "-v", "-Xcompiler", "-fPIC", "-c", src, "-o", o],
             verbose)
    objs.append(o)

    src = os.path.join(CSRC, "rt_image.cu")
    o = obj("rt_image")
    if force or _newer(o, [src] + headers):
        _run([nvcc, *ARCH, *NVCC_COMMON, "-fmad=false", "-Xptxas", "-v", "-Xcompiler", "-fPIC", "-c", src, "-o", o],
             verbose)
    objs.append(o)

    src = os.path.join(CSRC, "rt_filter_bench.cu")
    o = obj("rt_filter_bench")
    if force or _newer(o, [src] + headers):
        _run([nvcc, *ARCH, *NVCC_COMMON, "-fmad=false", "-Xptxas", "-v", "-Xcompiler", "-fPIC", "-c", src, "-o", o],
             verbose)
    objs.append(o)

    src = os.path.join(CSRC, "b200rt_api.cu")
    o = obj("b200rt_api")
    if force or _newer(o, [src] + headers):
        _run([nvcc, *ARCH, *NVCC_COMMON, "-fmad=false", "-Xcompiler", "-fPIC,-ffp-contract=off", "-c", src, "-o", o],
             verbose)
    objs.append(o)

    src = os.path.join(CSRC, "b200rt_group.cu")   # device groups (NCCL through dlopen: no link-time dependency)
    o = obj("b200rt_group")
    if force or _newer(o, [src] + headers):
        _run([nvcc, *ARCH, *NVCC_COMMON, "-Xcompiler", "-fPIC", "-c", src, "-o", o], verbose)
    objs.append(o)

    src = os.path.join(CSRC, "host_world.cpp")
    o = obj("host_world")
    if force or _newer(o, [src] + headers):
        _run(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-I", INCLUDE, "-I", CSRC,
              "-c", src, "-o", o], verbose)
    objs.append(o)

    if force or _newer(LIB_PATH, objs):
        _run([nvcc, *ARCH, "-shared", "-o", LIB_PATH, *objs, "-ldl", "-lpthread"], verbose)
    return LIB_PATH


if __name__ == "__main__":
    path = build_all(force="--force" in sys.argv, verbose="--verbose" in sys.argv or "-v" in sys.argv)
    print(path)
