"""Builds libaomarl.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libaomarl.so")
SOURCES = ["aomarl.cu"]
HEADERS = ["atmos_kernels.cuh", "gemm_kernels.cuh", "gemm_tc.cuh", "rtc_kernels.cuh", "wfs_kernels.cuh", "wfs_mma.cuh", "wfs_tma.cuh", "wfs_pipe.cuh", "wfs_tc.cuh", "geo_kernels.cuh", "pupil_sweep.cuh", "denoise_kernels.cuh", "fft16.cuh", "twiddles.cuh",
           "rng.cuh", "../../include/aomarl.h"]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(HERE, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-shared", "-Xcompiler", "-fPIC", "-o", LIB] + SOURCES
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    r = subprocess.run(cmd, cwd=HERE, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
