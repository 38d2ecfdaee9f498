"""Build libpcr.so (the C-ABI extension of include/pcr.h) in-tree with nvcc for sm_100a.

`python -m pointcloud_render_b200.build` or build() from __graft_entry__.  nvcc
cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.environ.get("PCR_LIB_OUT") or os.path.join(HERE, "libpcr.so")           # PCR_LIB_OUT: where a diagnostics build goes
SOURCES = [os.path.join(CSRC, "pcr_api.cu")]
DEPS = SOURCES + [os.path.join(CSRC, "pcr_kernels.cuh"), os.path.join(CSRC, "pcr_droplets.cuh"), os.path.join(ROOT, "include", "pcr.h")]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found")


def command(verbose=False):
    cmd = [nvcc_path(), "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
           "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-Wall",
           "-shared", "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-o", LIB] + SOURCES + ["-ldl"]
    if os.path.isfile("/usr/bin/g++"):
        cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    for flag in os.environ.get("PCR_NVCC_FLAGS", "").split():      # e.g. -DPCR_RASTER_STATS for diagnostics builds
        cmd.insert(1, flag)
    return cmd


def up_to_date():
    return os.path.isfile(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in DEPS)


def build(force=False, verbose=False):
    if not force and not verbose and up_to_date():
        return LIB
    res = subprocess.run(command(verbose), capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
