"""Build the CUDA library in-tree: ``marlon_b200/libcbx.so`` (sm_100a only, -lineinfo for ncu source pages)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libcbx.so")
SOURCES = ["cbx_kernels.cu", "cbx_api.cu"]
HEADERS = ["cbx_layout.h", "cbx_device.cuh", "cbx_pipe.cuh", "cbx_wide.cuh", os.path.join("..", "..", "include", "cbx.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--extended-lambda",
              "-Xcompiler", "-fPIC", "-shared", "-diag-suppress", "177"]


def _stale() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build_variant(name: str, defines: dict, verbose: bool = False) -> str:
    """Experiment builds: libcbx_<name>.so with -D overrides (CBX_TILE, CBX_MIN_CTAS, ...); load with CBX_LIB=<path>."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    out = os.path.join(HERE, f"libcbx_{name}.so")
    cmd = [nvcc] + NVCC_FLAGS + [f"-D{k}={v}" for k, v in defines.items()] + (["-Xptxas", "-v"] if verbose else []) \
        + ["-o", out] + [os.path.join(CSRC, f) for f in SOURCES]
    subprocess.check_call(cmd, cwd=CSRC)
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return SO
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build marlon_b200/libcbx.so (there is no CPU fallback)")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO] + [os.path.join(CSRC, f) for f in SOURCES]
    subprocess.check_call(cmd, cwd=CSRC)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
