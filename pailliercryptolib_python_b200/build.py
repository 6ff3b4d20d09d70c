"""Builds libphe_b200.so (CUDA kernels + C ABI) in-tree for sm_100a with nvcc, and the pybind11 shim.

python -m pailliercryptolib_python_b200.build [--force]
"""
import concurrent.futures
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(LIBDIR, "libphe_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--cudart", "static",
]
CU_SOURCES = ["phe_api.cu", "pipe_peak.cu", "pair_shapes.cu", "chacha20.cu"] + ["shape_%d_%d.cu" % s for s in ((20, 1), (20, 2), (20, 4), (20, 8), (15, 4), (15, 8))]
HEADERS = ["mont52.cuh", "paillier_items.cuh", "npair_items.cuh", "npair_kernels.cuh", "phe_kernels.cuh", "phe_launch.cuh", "phe_shapes.hpp", "hostbn.hpp",
           os.path.join("..", "..", "include", "phe_b200.h")]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s" % (" ".join(cmd), r.stdout))
    return r.stdout


def build_lib(force=False, verbose=False, variant=None, defines=()):
    """variant / defines: an A/B build of the kernels with extra -D flags into lib/libphe_b200_<variant>.so (selected at
    run time with PHE_B200_LIB, see capi.py); the default build has neither."""
    objdir = OBJDIR if not variant else OBJDIR + "_" + variant
    lib = LIB if not variant else os.path.join(LIBDIR, "libphe_b200_%s.so" % variant)
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(objdir, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    jobs = []
    objs = []
    for src in CU_SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _newer(o, [s] + hdrs):
            jobs.append([NVCC] + NVCC_FLAGS + ["-D" + d for d in defines] + ["-c", s, "-o", o])
    if jobs:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            for out in ex.map(_run, jobs):
                if verbose and out.strip():
                    print(out)
    if force or jobs or _newer(lib, objs):
        _run([NVCC, "-shared", "--cudart", "static", "-o", lib] + objs)
    return lib


def build_bindings(force=False):
    """pybind11 shim ipcl_bindings (thin: forwards to the C ABI)."""
    src = os.path.join(CSRC, "ipcl_bindings.cpp")
    if not os.path.exists(src):
        return None
    import pybind11

    ext = sysconfig.get_config_var("EXT_SUFFIX")
    out = os.path.join(HERE, "bindings", "ipcl_bindings" + ext)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if force or _newer(out, [src, os.path.join(HERE, "..", "include", "phe_b200.h")]):
        cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-fvisibility=hidden",
               "-I", pybind11.get_include(), "-I", sysconfig.get_paths()["include"],
               "-I", os.path.join(HERE, "..", "include"), src, "-o", out,
               "-L", LIBDIR, "-lphe_b200", "-Wl,-rpath,$ORIGIN/../lib"]
        _run(cmd)
    return out


def build_all(force=False, verbose=False):
    lib = build_lib(force, verbose)
    build_bindings(force)
    return lib


if __name__ == "__main__":
    if "--variant" in sys.argv:     # python -m pailliercryptolib_python_b200.build --variant u4 PHE52_U=4 ...
        i = sys.argv.index("--variant")
        print(build_lib(variant=sys.argv[i + 1], defines=sys.argv[i + 2:], verbose=True))
    else:
        print(build_all(force="--force" in sys.argv, verbose=True))
