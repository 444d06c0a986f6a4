"""Build libfocr_b200.so in-tree with nvcc for sm_100a (and nothing else).

    python font-ocr_b200/build.py [--force]

The library is the product: hand-written CUDA kernels + the C ABI of include/focr_b200.h.
It is built in-tree (font-ocr_b200/libfocr_b200.so, git-ignored) so that it travels to the GPU
box with the gpurun snapshot.  cudart is linked statically and libcuda is only reached through
cudaGetDriverEntryPoint, so the .so also loads on a machine without a driver (the CPU test box),
where every compute entry then returns FOCR_ERR_CUDA.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libfocr_b200.so")
STAMP = OUT + ".stamp"
SOURCES = ["api.cu", "stats.cu", "scan_simt.cu", "scan_tc.cu", "finalize.cu", "postprocess.cu", "focr_decode.cu", "multi.cpp", "../host/focr_host.cpp", "../host/focr_raster.cpp"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "--fmad=false",              # no implicit contraction: the only fused ops are the explicit __fma_rn
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off",
    "-shared", "-cudart", "static",
]


def _digest() -> str:
    h = hashlib.sha256()
    files = sorted(os.listdir(CSRC)) + ["../../include/focr_b200.h", "../build.py", "../host/focr_host.cpp",
                                        "../host/focr_host.hpp", "../host/focr_raster.cpp", "../host/focr_cli.cpp"]
    for f in files:
        p = os.path.join(CSRC, f)
        if os.path.isfile(p):
            h.update(f.encode())
            h.update(open(p, "rb").read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Build the product library.  With FOCR_TC_EXPERIMENTS=1 in the environment the kernel's experiment hooks are compiled
    in and the result goes to libfocr_b200_exp.so (tools/ select it with FOCR_B200_LIB); the product library is untouched."""
    exp = bool(os.environ.get("FOCR_TC_EXPERIMENTS"))
    variant = os.environ.get("FOCR_BUILD_VARIANT", "")   # tools/: "name:-DFLAG ..." builds libfocr_b200_<name>.so with extra flags
    out = OUT.replace(".so", "_exp.so") if exp else OUT
    if variant:
        out = OUT.replace(".so", "_" + variant.split(":")[0] + ".so")
    stamp = out + ".stamp"
    dg = _digest() + ("+exp" if exp else "") + variant
    if not force and os.path.exists(out) and os.path.exists(stamp) and open(stamp).read() == dg:
        if not exp and not variant and not os.path.exists(os.path.join(HERE, "bin", "focr_cli")):
            build_cli(out)
        return out
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = ["-DFOCR_TC_EXPERIMENTS"] if exp else []  # tools/tc_trace.py, tools/tc_modes.py, tools/tc_timeline.py
    if variant and ":" in variant:
        extra += variant.split(":", 1)[1].split()
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + \
          [os.path.join(CSRC, s) for s in SOURCES] + ["-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libfocr_b200.so")
    if verbose:
        sys.stderr.write(r.stderr)
    open(stamp, "w").write(dg)
    if not exp and not variant:
        build_cli(out)
    return out


def build_cli(lib: str) -> str:
    """The C++ front-end font-ocr_b200/bin/focr_cli (host/focr_cli.cpp): plain g++, linked against the library next to it."""
    cli = os.path.join(HERE, "bin", "focr_cli")
    os.makedirs(os.path.dirname(cli), exist_ok=True)
    cmd = [os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-Wall", os.path.join(HERE, "host", "focr_cli.cpp"), "-o", cli,
           "-L" + HERE, "-l:" + os.path.basename(lib), "-lz", "-Wl,-rpath,$ORIGIN/.."]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("g++ failed building focr_cli")
    return cli


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
