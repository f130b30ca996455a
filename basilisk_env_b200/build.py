"""In-tree build of the CUDA library (libbskenv.so) for sm_100a with plain nvcc.

nvcc cross-compiles without a GPU, so this runs on the CPU build box; the .so travels to the GPU box
with the repo snapshot (git-ignored, not gpurun-ignored)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.environ.get("BSKENV_LIB") or os.path.join(HERE, "libbskenv.so")   # BSKENV_LIB: load a tuning variant instead
SOURCES = ["bskenv.cu", "opnav.cu"]
DEPS = ["bskenv.cu", "leo_core.cuh", "leo_f32.cuh", "leo_params.h", "leo_host.h", "opnav.cu", "opnav_core.cuh", "opnav_params.h", "opnav_host.h",
        os.path.join("..", "..", "include", "bskenv.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found: cannot build libbskenv.so")
    return p


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force=False, extra_flags=(), verbose=False, out=None):
    if out is None and not force and not is_stale():
        return LIB
    out = out or LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + list(extra_flags) + ["-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd, cwd=CSRC)
    return out


if __name__ == "__main__":
    print(build(force=True, verbose=True))
