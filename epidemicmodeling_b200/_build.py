"""In-tree build of libepi_b200.so (hand-written CUDA for sm_100a + the C ABI).

nvcc cross-compiles without a GPU.  Flags that are part of the arithmetic
contract: --fmad=false (no implicit FMA contraction; fma() appears
only where the contract names it), default IEEE division/sqrt (-prec-div/-prec-sqrt
are nvcc's FP64 defaults and FP64 has no fast variants).
"""
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libepi_b200.so")
OBJ_DIR = os.path.join(PKG, "build")
SOURCES = ["capi.cu", "ekf_forward.cu", "eks_gain.cu", "eks_backward.cu", "ekf_rows.cu", "ekf_pair.cu", "other_kernels.cu", "pareto_sorted.cu", "rt_expfit.cu", "preprocess.cu", "nnls.cu"]
HEADERS = [os.path.join(CSRC, h) for h in ("epi_device.cuh", "epi_internal.h", "epi_linalg.cuh", "epi_linalg_experiments.cuh", "ekf_common.cuh", "epi_async.cuh")] + \
          [os.path.join(ROOT, "include", "epi_b200.h")]
NVCC_FLAGS = ["-O3", "--fmad=false", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-std=c++17", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libepi_b200.so cannot be built (there is no CPU fallback)")
    return exe


def _digest():
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for f in [os.path.join(CSRC, s) for s in SOURCES] + HEADERS:
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def build(force=False, verbose=False, extra_flags=(), lib=LIB, obj_dir=OBJ_DIR):
    """Compile (if sources changed) and return the path of libepi_b200.so.
    extra_flags/lib/obj_dir build kernel-variant experiments (e.g. -DEPI_GAIN_MIN_BLOCKS=4)."""
    stamp = os.path.join(obj_dir, "stamp")
    dig = _digest() + " ".join(extra_flags)
    if not force and os.path.exists(lib) and os.path.exists(stamp) and open(stamp).read() == dig:
        return lib
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + list(extra_flags) + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        with open(obj + ".log", "w") as fh:
            fh.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as fh:
        fh.write(dig)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
