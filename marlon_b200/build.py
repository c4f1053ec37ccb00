"""Build the CUDA library in-tree: ``marlon_b200/libcbx.so`` (sm_100a only, -lineinfo for ncu source pages).

The five translation units (API, fused kernel + sampler + GAE, pipelined kernel for each defender binding, warp-per-tile kernel) compile in parallel
to ``build/*.o`` and are linked into one shared library.  ``-DCBX_EXPERIMENTS`` (``build_variant``) adds the section-skip
timing switches and the plain-copy (non-TMA) staging variants; the release library has neither.
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libcbx.so")
OBJ = os.path.join(HERE, "build")
SOURCES = ["cbx_kernels.cu", "cbx_pipe.cu", "cbx_pipe_live.cu", "cbx_wide.cu", "cbx_api.cu"]
HEADERS = ["cbx_layout.h", "cbx_device.cuh", "cbx_shared.cuh", "cbx_pipe.cuh", "cbx_wide.cuh", os.path.join("..", "..", "include", "cbx.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--extended-lambda",
              "-Xcompiler", "-fPIC", "-diag-suppress", "177"]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build marlon_b200/libcbx.so (there is no CPU fallback)")
    return nvcc


def _stale(out: str) -> bool:
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def _compile_link(out: str, tag: str, defines: dict, verbose: bool, force: bool = False) -> str:
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    dflags = [f"-D{k}={v}" for k, v in defines.items()]
    vflags = ["-Xptxas", "-v"] if verbose else []

    def one(src):
        obj = os.path.join(OBJ, f"{tag}{src[:-3]}.o")
        srcp = os.path.join(CSRC, src)
        fresh = (not force and not verbose and os.path.exists(obj)
                 and os.path.getmtime(obj) > max(os.path.getmtime(p) for p in [srcp] + [os.path.join(CSRC, h) for h in HEADERS]))
        if not fresh:
            r = subprocess.run([nvcc] + NVCC_FLAGS + dflags + vflags + ["-c", srcp, "-o", obj], cwd=CSRC, capture_output=True, text=True)
            if r.returncode != 0 or verbose:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src}")
        return obj

    with ThreadPoolExecutor(len(SOURCES)) as ex:
        objs = list(ex.map(one, SOURCES))
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out] + objs, cwd=CSRC)
    return out


def build_variant(name: str, defines: dict, verbose: bool = False) -> str:
    """Experiment builds: libcbx_<name>.so with -D overrides (CBX_EXPERIMENTS, CBX_TILE, ...); load with CBX_LIB=<path>."""
    return _compile_link(os.path.join(HERE, f"libcbx_{name}.so"), f"{name}_", defines, verbose, force=True)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not verbose and not _stale(SO):
        return SO
    return _compile_link(SO, "", {}, verbose, force=force)


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
