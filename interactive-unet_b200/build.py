"""Build `libiunet_b200.so` (the C-ABI library of include/iunet_b200.h) in-tree with nvcc for sm_100a.

nvcc cross-compiles without a GPU, so this runs in the CPU-only build container; the resulting
`.so` is git-ignored but travels to the GPU box with the repo snapshot.  The library links the CUDA
runtime statically and reaches the one driver symbol it needs (`cuTensorMapEncodeTiled`) through
`cudaGetDriverEntryPoint`, so it loads (for symbol checks) on machines without a driver as well.
"""
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
BUILD_DIR = os.path.join(PKG_DIR, "build")
LIB_PATH = os.path.join(PKG_DIR, "libiunet_b200.so")
SOURCES = ["conv_tc.cu", "conv_tc2.cu", "conv_halo.cu", "conv_row.cu", "conv_chain.cu", "conv_stem.cu", "aux_kernels.cu", "engine.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: cannot build libiunet_b200.so")
    return exe


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a and link the shared library.  Returns its path."""
    os.makedirs(BUILD_DIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(REPO_ROOT, "include", "iunet_b200.h"))
    nvcc = _nvcc()
    objects = []
    for src in SOURCES:
        src_path = os.path.join(CSRC, src)
        obj = os.path.join(BUILD_DIR, src.replace(".cu", ".o"))
        objects.append(obj)
        if force or _stale(obj, [src_path] + headers):
            cmd = [nvcc, *ARCH, "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC,-Wall", "-Xptxas", "-v",
                   "-I", os.path.join(REPO_ROOT, "include"), "-I", CSRC, "-c", src_path, "-o", obj]
            res = subprocess.run(cmd, capture_output=True, text=True)
            if verbose or res.returncode:
                sys.stderr.write(res.stdout + res.stderr)
            if res.returncode:
                raise RuntimeError(f"nvcc failed on {src}")
    if force or _stale(LIB_PATH, objects):
        cmd = [nvcc, *ARCH, "-shared", "-o", LIB_PATH, *objects, "-cudart", "static"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or res.returncode:
            sys.stderr.write(res.stdout + res.stderr)
        if res.returncode:
            raise RuntimeError("link of libiunet_b200.so failed")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
