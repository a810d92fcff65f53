"""In-tree build of the native extension ``ray_shuffling_data_loader_b200._C``.

``nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo`` for the kernels, g++
for the pybind11 runtime, static cudart - the resulting ``.so`` sits next to the
sources so it travels to the GPU box with the repo snapshot (no JIT cache).

    python -m ray_shuffling_data_loader_b200._build [--force] [--verbose]
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig
from typing import List

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
BUILD_DIR = os.path.join(PKG_DIR, "csrc", "build")
CUDA_HOME = os.environ.get("CUDA_HOME", "/usr/local/cuda")
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
CU_SOURCES = ["shuffle_kernels.cu"]
CPP_SOURCES = ["bindings.cpp"]
HEADERS = ["common.cuh", "perm.cuh", "kernels.h"]


def ext_path() -> str:
    return os.path.join(PKG_DIR, "_C" + sysconfig.get_config_var("EXT_SUFFIX"))


def _newer(target: str, deps: List[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd: List[str], verbose: bool):
    if verbose:
        print(" ".join(cmd), flush=True)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"build step failed: {' '.join(cmd)}\n{res.stdout}\n{res.stderr}")
    if verbose and (res.stdout or res.stderr):
        print(res.stdout, res.stderr)


def build_variant(name: str, defines: List[str], verbose: bool = False) -> str:
    """Build an experimental variant of the extension (different -D tile
    geometry) into ``csrc/build/variants/<name>/_C*.so`` for A/B kernel runs
    (``tools/kernel_bench.py --ext``)."""
    import pybind11
    out_dir = os.path.join(BUILD_DIR, "variants", name)
    os.makedirs(out_dir, exist_ok=True)
    target = os.path.join(out_dir, "_C" + sysconfig.get_config_var("EXT_SUFFIX"))
    nvcc = os.path.join(CUDA_HOME, "bin", "nvcc")
    includes = ["-I", CSRC, "-I", pybind11.get_include(),
                "-I", sysconfig.get_paths()["include"],
                "-I", os.path.join(CUDA_HOME, "include")]
    objs = []
    for src in CU_SOURCES:
        obj = os.path.join(out_dir, src + ".o")
        _run([nvcc, "-std=c++17", "-O3", *ARCH_FLAGS, "-lineinfo", *[f"-D{d}" for d in defines],
              "-Xcompiler", "-fPIC", *includes, "-c", os.path.join(CSRC, src), "-o", obj], verbose)
        objs.append(obj)
    # the bindings object is geometry independent: reuse the main build's
    objs.append(os.path.join(BUILD_DIR, CPP_SOURCES[0] + ".o"))
    _run(["g++", "-shared", "-o", target, *objs, "-L", os.path.join(CUDA_HOME, "lib64"),
          "-lcudart_static", "-lpthread", "-ldl", "-lrt"], verbose)
    return target


def build(force: bool = False, verbose: bool = False) -> str:
    import pybind11
    os.makedirs(BUILD_DIR, exist_ok=True)
    target = ext_path()
    headers = [os.path.join(CSRC, h) for h in HEADERS]
    all_src = ([os.path.join(CSRC, s) for s in CU_SOURCES + CPP_SOURCES] + headers
               + [os.path.abspath(__file__)])
    if not force and not _newer(target, all_src):
        return target
    nvcc = os.path.join(CUDA_HOME, "bin", "nvcc")
    includes = ["-I", CSRC, "-I", pybind11.get_include(),
                "-I", sysconfig.get_paths()["include"],
                "-I", os.path.join(CUDA_HOME, "include")]
    objs = []
    for src in CU_SOURCES:
        obj = os.path.join(BUILD_DIR, src + ".o")
        if force or _newer(obj, [os.path.join(CSRC, src)] + headers):
            _run([nvcc, "-std=c++17", "-O3", *ARCH_FLAGS, "-lineinfo",
                  *(["-Xptxas", "-v"] if verbose else []),
                  "-Xcompiler", "-fPIC", *includes, "-c", os.path.join(CSRC, src),
                  "-o", obj], verbose)
        objs.append(obj)
    for src in CPP_SOURCES:
        obj = os.path.join(BUILD_DIR, src + ".o")
        if force or _newer(obj, [os.path.join(CSRC, src)] + headers):
            _run(["g++", "-std=c++17", "-O2", "-fPIC", "-fvisibility=hidden", *includes,
                  "-c", os.path.join(CSRC, src), "-o", obj], verbose)
        objs.append(obj)
    _run(["g++", "-shared", "-o", target, *objs,
          "-L", os.path.join(CUDA_HOME, "lib64"), "-lcudart_static",
          "-lpthread", "-ldl", "-lrt"], verbose)
    return target


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
