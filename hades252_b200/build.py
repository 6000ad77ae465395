"""Builds hades252_b200/lib/libhades_b200.so with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib", "libhades_b200.so")
SOURCES = ["hades_engine.cu", "hades_w3.cu", "hades_w5.cu", "hades_w9.cu",
           "hades_w3_dense.cu", "hades_w5_dense.cu", "hades_w9_dense.cu",
           "hades_w3_ccf.cu", "hades_w5_ccf.cu", "hades_w9_ccf.cu", "hades_generic.cu", "hades_frtest.cu"]
HEADERS = ["fr.cuh", "hades.cuh", "coop.cuh", "width_impl.cuh", "width_ops.hpp", "util_kernels.cuh", "host_tables.hpp",
           os.path.join("..", "..", "include", "hades_cuda.h")]
_KERNEL_HDRS = ["fr.cuh", "hades.cuh", "width_impl.cuh", "width_ops.hpp"]


def _deps(src: str):
    """headers a translation unit includes (an object is rebuilt only when one of them is newer)"""
    if src == "hades_engine.cu":
        return ["fr.cuh", "host_tables.hpp", "util_kernels.cuh", "width_ops.hpp", os.path.join("..", "..", "include", "hades_cuda.h")]
    if src == "hades_frtest.cu":
        return ["fr.cuh", "width_ops.hpp"]
    if src == "hades_generic.cu":
        return ["fr.cuh", "hades.cuh", "width_ops.hpp"]
    if src == "hades_w5_ccf.cu":
        return _KERNEL_HDRS + ["coop.cuh"]
    return _KERNEL_HDRS
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    """One nvcc -c per translation unit (in parallel), then one link into the shared library."""
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    # HADES_BUILD_TAG=x: an experimental build (with HADES_NVCC_EXTRA flags) beside the default one:
    # lib/libhades_b200_x.so from lib/obj_x/, selected at run time with HADES_B200_LIB
    tag = os.environ.get("HADES_BUILD_TAG", "")
    lib_path = LIB.replace(".so", f"_{tag}.so") if tag else LIB
    objdir = os.path.join(HERE, "lib", "obj" + (f"_{tag}" if tag else ""))
    os.makedirs(objdir, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("HADES_NVCC_EXTRA", "").split()  # e.g. -DHADES_SYNC_MID for experiments

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        # HADES_NVCC_EXTRA_<unit> (e.g. HADES_NVCC_EXTRA_hades_w5_ccf): flags for ONE translation unit, so that a tagged
        # experimental build recompiles only that unit (the others are copied from the default build's objects)
        own = os.environ.get("HADES_NVCC_EXTRA_" + src.replace(".cu", ""), "").split()
        cmd = [nvcc, *NVCC_FLAGS, *extra, *own, "-c", "-o", obj, os.path.join(CSRC, src)]
        if tag and not own and not extra:
            base = os.path.join(HERE, "lib", "obj", src.replace(".cu", ".o"))
            if os.path.exists(base) and os.path.exists(base + ".cmd"):
                import shutil
                shutil.copy2(base, obj)
                res = subprocess.CompletedProcess(cmd, 0, stdout=open(base + ".cmd").read().split("\n", 1)[1])
                res.fresh = False
                return src, obj, cmd, res
        stamp = obj + ".cmd"   # the command line and ptxas report of the object on disk
        fresh = (not force and os.path.exists(obj) and os.path.exists(stamp)
                 and open(stamp).readline().rstrip("\n") == " ".join(cmd)
                 and all(os.path.getmtime(os.path.join(CSRC, f)) <= os.path.getmtime(obj) for f in [src] + _deps(src)))
        if fresh:
            res = subprocess.CompletedProcess(cmd, 0, stdout=open(stamp).read().split("\n", 1)[1])
            res.fresh = True
            return src, obj, cmd, res
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if res.returncode == 0:
            with open(stamp, "w") as f:
                f.write(" ".join(cmd) + "\n" + res.stdout)
        return src, obj, cmd, res

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        results = list(pool.map(compile_one, SOURCES))
    log = os.path.join(HERE, "lib", "build.log" if not tag else f"build_{tag}.log")
    with open(log, "w") as f:
        for src, obj, cmd, res in results:
            f.write(" ".join(cmd) + "\n" + res.stdout + "\n")
    for src, obj, cmd, res in results:
        if verbose or res.returncode:
            sys.stderr.write(res.stdout)
        if res.returncode:
            raise RuntimeError(f"nvcc failed on {src} ({res.returncode}); see {log}")
    if (os.path.exists(lib_path) and all(getattr(r[3], "fresh", False) for r in results)
            and all(os.path.getmtime(r[1]) <= os.path.getmtime(lib_path) for r in results)):
        return lib_path
    link = [nvcc, "-shared", "-cudart", "static", "-o", lib_path, *[r[1] for r in results], "-ldl", "-lpthread"]
    res = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    with open(log, "a") as f:
        f.write(" ".join(link) + "\n" + res.stdout)
    if res.returncode:
        sys.stderr.write(res.stdout)
        raise RuntimeError(f"link failed ({res.returncode}); see {log}")
    return lib_path


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
