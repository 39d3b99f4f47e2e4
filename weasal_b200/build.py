"""Builds libweasal_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["primitives.cu", "grid_subsample.cu", "radius_search.cu", "kpconv_simt.cu", "kpconv_tc.cu", "pool_ops.cu", "pyramid.cu", "sphere_vote.cu", "sm_partition.cu", "capi.cu"]
LIB = os.path.join(HERE, "libweasal_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def nvcc_path():
    return os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    deps.append(os.path.join(HERE, "..", "include", "weasal_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defines=(), lib=None):
    """``defines`` / ``lib``: an experiment build with extra -D flags into another file name (A/B runs on the GPU box
    select it with WEASAL_B200_LIB); the default build is the product."""
    lib = lib or LIB
    if not force and not defines and not needs_build():
        return lib
    import fcntl
    with open(os.path.join(HERE, ".build.lock"), "w") as lock:  # several ranks may arrive here at once
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and not defines and not needs_build():
            return lib
        return _build_locked(force, verbose, defines, lib)


def _build_locked(force, verbose, defines, lib):
    nvcc = nvcc_path()
    objs = []
    tag = "" if not defines else "." + "_".join(d.replace("=", "") for d in defines)
    for src in SOURCES:
        obj = os.path.join(CSRC, src.replace(".cu", tag + ".o"))
        cmd = ([nvcc] + NVCC_FLAGS + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if verbose else [])
               + ["-c", os.path.join(CSRC, src), "-o", obj])
        subprocess.check_call(cmd)
        objs.append(obj)
    subprocess.check_call([nvcc, "-shared", "-o", lib + ".tmp"] + objs + ["-lcudart"])
    os.replace(lib + ".tmp", lib)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
