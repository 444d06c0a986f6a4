"""ctypes loader of tools/microbench/libfocr_microbench.so (focr_microbench.h): tcgen05 / TMEM / mbarrier
micro-benchmarks.  Measurement aids only -- not part of the product library or its header."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libfocr_microbench.so")
SRC = os.path.join(HERE, "umma_bench.cu")
SYMBOLS = ["focr_microbench_last_error", "focr_bench_umma_i8", "focr_bench_umma_i8_2cta", "focr_bench_umma_issue_cycles", "focr_bench_pingpong",
           "focr_bench_tmem"]


def build(force: bool = False) -> str:
    """nvcc -> libfocr_microbench.so, sm_100a only (cross-compiles without a GPU)."""
    deps = [SRC, os.path.join(HERE, "focr_microbench.h")]
    if not force and os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(d) for d in deps):
        return SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
           "-shared", "-cudart", "static", SRC, "-o", SO]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libfocr_microbench.so\n" + r.stdout + r.stderr)
    return SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            raise ImportError(f"{SO} is missing: run __graft_entry__.build()")
        l = C.CDLL(SO)
        vp, i = C.c_void_p, C.c_int
        l.focr_microbench_last_error.restype = C.c_char_p
        l.focr_bench_umma_i8.argtypes = [i, i, i, i, i, vp, vp]
        l.focr_bench_umma_i8_2cta.argtypes = [i, i, i, i, i, vp, vp]
        l.focr_bench_umma_issue_cycles.restype = C.c_double
        l.focr_bench_pingpong.argtypes = [i, i, i, i, i, vp]
        l.focr_bench_tmem.argtypes = [i, i, i, i, i, i, vp, vp]
        _lib = l
    return _lib


def check(rc: int):
    if rc != 0:
        raise RuntimeError(f"libfocr_microbench error {rc}: {lib().focr_microbench_last_error().decode()}")


def ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def int8_peak_tops(device: int = 0, n: int = 256, ksteps: int = 7, iters: int = 4000, sm_count: int = 148) -> float:
    """Measured dense int8 tensor peak: M128 x N x K32 tcgen05.mma kind::i8 back to back on every SM."""
    cyc, ms = np.zeros(1), np.zeros(1)
    check(lib().focr_bench_umma_i8(device, n, ksteps, iters, 1, ptr(cyc), ptr(ms)))
    return 2.0 * 128 * n * 32 * ksteps * iters * sm_count / (float(ms[0]) * 1e-3) / 1e12
