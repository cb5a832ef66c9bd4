"""In-tree build of the native pieces (no JIT cache: the built .so files travel to the GPU box with the snapshot).

  libmoolib_b200.so   CUDA kernels + C-ABI (include/moolib_b200.h), nvcc, sm_100a only
  _C*.so              pybind11 host layer mirroring moolib's Python API (moolib_b200/csrc/host), g++ over libtorch

`python moolib_b200/build.py [--force] [--only lib|host]`
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libmoolib_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

CUDA_SOURCES = ["mb_core.cu", "mb_copy.cu", "mb_allreduce.cu", "mb_learner.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared",
]


def _newer(srcs, target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in srcs)


def build_lib(force=False, verbose=False):
    srcs = [os.path.join(CSRC, s) for s in CUDA_SOURCES]
    deps = srcs + [os.path.join(CSRC, "mb_common.cuh"), os.path.join(ROOT, "include", "moolib_b200.h")]
    if not force and not _newer(deps, LIB):
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + srcs
    subprocess.run(cmd, check=True, cwd=CSRC)
    return LIB


def host_ext_path():
    return os.path.join(HERE, "_C" + sysconfig.get_config_var("EXT_SUFFIX"))


def host_sources():
    d = os.path.join(CSRC, "host")
    if not os.path.isdir(d):
        return []
    return sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cc"))


def build_host(force=False):
    """pybind11 module over libtorch; links libmoolib_b200.so by rpath ($ORIGIN/lib)."""
    srcs = host_sources()
    if not srcs:
        return None
    out = host_ext_path()
    hdrs = [os.path.join(CSRC, "host", f) for f in os.listdir(os.path.join(CSRC, "host")) if f.endswith(".h")]
    deps = srcs + hdrs + [os.path.join(ROOT, "include", "moolib_b200.h")]
    if not force and not _newer(deps, out):
        return out
    import pybind11
    import torch

    tdir = os.path.dirname(torch.__file__)
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    flags = [
        "-std=c++17", "-O2", "-fPIC", "-fvisibility=hidden", "-Wall", "-Wno-unused-function", "-DNDEBUG",
        f"-D_GLIBCXX_USE_CXX11_ABI={abi}", "-DTORCH_EXTENSION_NAME=_C",
        "-I" + os.path.join(ROOT, "include"), "-I" + pybind11.get_include(),
        "-I" + sysconfig.get_paths()["include"], "-I" + os.path.join(tdir, "include"),
        "-I" + os.path.join(tdir, "include", "torch", "csrc", "api", "include"), "-I/usr/local/cuda/include",
    ]
    procs = []
    objs = []
    for s in srcs:
        o = os.path.join(objdir, os.path.basename(s) + ".o")
        objs.append(o)
        if force or _newer([s] + hdrs, o):
            procs.append((s, subprocess.Popen(["g++"] + flags + ["-c", s, "-o", o])))
    for s, p in procs:
        if p.wait() != 0:
            raise RuntimeError(f"compiling {s} failed")
    link = ["g++", "-shared", "-o", out] + objs + [
        "-L" + os.path.join(tdir, "lib"), "-ltorch", "-ltorch_cpu", "-lc10", "-ltorch_python", "-lc10_cuda",
        "-ltorch_cuda", "-L" + LIBDIR, "-lmoolib_b200", "-lpthread", "-lrt",
        "-Wl,-rpath," + os.path.join(tdir, "lib"), "-Wl,-rpath,$ORIGIN/lib",
    ]
    subprocess.run(link, check=True)
    return out


def build_all(force=False, verbose=False):
    lib = build_lib(force, verbose)
    host = build_host(force)
    return lib, host


if __name__ == "__main__":
    force = "--force" in sys.argv
    only = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else None
    if only in (None, "lib"):
        print(build_lib(force, "-v" in sys.argv))
    if only in (None, "host"):
        print(build_host(force))
