"""Builds libdint_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdint_b200.so")
DRIVER = os.path.join(HERE, "dint_b200_decode")
SOURCES = [os.path.join(CSRC, "dint_kernels.cu"), os.path.join(CSRC, "dint_discover.cu"), os.path.join(CSRC, "dint_host.cpp")]
HEADERS = [os.path.join(CSRC, h) for h in ("dint_internal.hpp", "dint_cuda_util.hpp", "dint_decoder.cuh", "dint_strips.cuh", "dint_discover.cuh",
                                            "dint_index_blocks.cuh", "dint_block_codecs.cuh")] + [os.path.join(ROOT, "include", "dint_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    """Compile the CUDA library (and the C++ driver) if sources are newer. Returns the .so path."""
    nvcc = _nvcc()
    if force or _stale(LIB, SOURCES + HEADERS):
        if nvcc is None:
            if os.path.exists(LIB):
                return LIB
            raise RuntimeError("nvcc not found and libdint_b200.so is not built")
        cmd = [nvcc] + NVCC_FLAGS + ["-shared", "-o", LIB] + SOURCES
        if verbose:
            print("+", " ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)
    main = os.path.join(CSRC, "vroom_decode_main.cpp")
    if os.path.exists(main) and (force or _stale(DRIVER, [main, LIB] + HEADERS)):
        cxx = os.environ.get("CXX") or shutil.which("g++") or "g++"
        cmd = [cxx, "-O2", "-std=c++17", "-Wall", "-o", DRIVER, main, "-L", HERE, "-ldint_b200", "-Wl,-rpath,$ORIGIN"]
        if verbose:
            print("+", " ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)
    return LIB
