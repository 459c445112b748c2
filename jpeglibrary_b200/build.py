"""Builds the native libraries in-tree (no JIT cache: the .so files travel with the repo snapshot).

  lib/libjpegb200.so       CUDA kernels + C-ABI shim   (nvcc, sm_100a only)
  lib/libjpegb200_host.so  host marker walk + C++ mirror of the reference API (g++)
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
# JB_LIBDIR / JB_BUILD_DEFINES: A/B builds of kernel variants for profiling (e.g. JB_BUILD_DEFINES="-DJB_K2_PACKED=0"
# JB_LIBDIR=.../lib_scalar); the package loads JB_LIBDIR when it is set, the default build is what ships
LIB = os.environ.get("JB_LIBDIR") or os.path.join(HERE, "lib")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CXX = os.environ.get("CXX", "g++")

CUDA_SRC = [os.path.join(HERE, "csrc", "jpegb200.cu")]
CUDA_DEPS = [os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc"))] + [
    os.path.join(HERE, "..", "include", "jpegb200.h")]
HOST_SRC = [os.path.join(HERE, "host", f) for f in sorted(os.listdir(os.path.join(HERE, "host"))) if f.endswith(".cpp")]
HOST_DEPS = HOST_SRC + [os.path.join(HERE, "..", "include", f) for f in os.listdir(os.path.join(HERE, "..", "include"))]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "-diag-suppress", "550",
] + os.environ.get("JB_BUILD_DEFINES", "").split()


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    os.makedirs(LIB, exist_ok=True)
    cuda_so = os.path.join(LIB, "libjpegb200.so")
    host_so = os.path.join(LIB, "libjpegb200_host.so")
    if force or _stale(cuda_so, CUDA_DEPS):
        cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", cuda_so] + CUDA_SRC
        subprocess.check_call(cmd)
    if force or _stale(host_so, HOST_DEPS):
        cmd = [CXX, "-O2", "-g", "-std=c++17", "-shared", "-fPIC", "-fvisibility=hidden", "-Wall", "-o", host_so] + HOST_SRC + [
            "-lpthread", "-L" + LIB, "-ljpegb200", "-Wl,-rpath,$ORIGIN"]
        subprocess.check_call(cmd)
    return cuda_so, host_so


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
