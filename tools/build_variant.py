"""Build a kernel-variant copy of the library for A/B runs (development aid).

    python tools/build_variant.py <tag> -DIU_PAIR_B_STAGES=16 -DIU_PAIR_A_STAGES=2 ...
writes interactive-unet_b200/build/variants/libiunet_<tag>.so; select it with IU_LIB=<path>.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "interactive-unet_b200")
CSRC = os.path.join(PKG, "csrc")
SOURCES = ["conv_tc.cu", "conv_halo.cu", "conv_row.cu", "conv_stem.cu", "aux_kernels.cu", "engine.cu"]


def main():
    tag, defs = sys.argv[1], sys.argv[2:]
    out_dir = os.path.join(PKG, "build", "variants", tag)
    os.makedirs(out_dir, exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(out_dir, src.replace(".cu", ".o"))
        objs.append(obj)
        procs.append(subprocess.Popen(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
                                       "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", CSRC, *defs,
                                       "-c", os.path.join(CSRC, src), "-o", obj]))
    for p in procs:
        if p.wait():
            raise SystemExit("nvcc failed")
    lib = os.path.join(PKG, "build", "variants", f"libiunet_{tag}.so")
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", lib, *objs, "-cudart", "static"])
    print(lib)


if __name__ == "__main__":
    main()
