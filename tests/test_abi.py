"""The drop-in boundary: the library builds for sm_100a, loads without a GPU, exports every symbol
include/focr_b200.h declares, and refuses to compute without a device (no CPU fallback)."""
import numpy as np
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "focr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b([a-z_0-9]+)\s*\([^;{]*\)\s*;", src)
    return sorted(set(n for n in names if n.startswith(("focr_", "ncc_"))))


def test_header_symbols_are_exported(built_lib):
    declared = _declared_symbols()
    assert "ncc_8_u8" in declared and "ncc_16_u8" in declared and "focr_ncc_scan" in declared
    lib = C.CDLL(built_lib)
    missing = [n for n in declared if not hasattr(lib, n)]
    assert not missing, missing
    from font_ocr_b200 import native

    assert sorted(native.SYMBOLS) == declared


def test_built_for_sm100a_only(built_lib):
    out = subprocess.run(["cuobjdump", "-lelf", built_lib], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_match_struct_layout(pkg):
    from font_ocr_b200 import native

    assert native.MATCH_DTYPE.itemsize == 8  # ncc.cpp:7-10
    assert native.MATCH_DTYPE.fields["similarity"][1] == 4
    assert native.RASTER_DTYPE.itemsize == 16


def test_no_cpu_fallback(built_lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from font_ocr_b200 import native, ncc

    with pytest.raises(native.FocrError) as e:
        ncc.Context(0)
    assert e.value.code == native.FOCR_ERR_CUDA
    assert b"no CPU fallback" in native.lib().focr_last_error()


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: nothing under font-ocr_b200/ may import, link or load it."""
    pkg_dir = os.path.join(ROOT, "font-ocr_b200")
    for dp, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                s = open(os.path.join(dp, f)).read()
                assert "from oracle" not in s and "import oracle" not in s and "libncc_oracle" not in s \
                    and "libncc_ref" not in s, os.path.join(dp, f)


def test_staging_pool_copies(built_lib):
    """The staging copies (pageable callers: api.cu `parallel_copy` on its persistent per-thread pool) need no GPU: contiguous
    and strided copies of many sizes, repeated (the pool is reused), and from several host threads at once (one pool each,
    like the devices of a focr_multi call) must move every byte."""
    import ctypes as C
    import threading

    lib = C.CDLL(built_lib)
    copy = getattr(lib, "_Z27focr_internal_parallel_copyPhmPKhmmm")   # (dst, dst_stride, src, src_stride, row_bytes, rows), C++ linkage
    copy.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t]
    copy.restype = None
    rng = np.random.default_rng(3)

    def work(seed, errors):
        r = np.random.default_rng(seed)
        for it in range(12):
            n = int(r.integers(1, 24 << 20))
            src = r.integers(0, 256, n, dtype=np.uint8)
            dst = np.zeros(n, np.uint8)
            copy(dst.ctypes.data, 0, src.ctypes.data, 0, n, 1)                       # one contiguous block
            if not np.array_equal(src, dst):
                errors.append(("contiguous", seed, it))
            rows, rb = int(r.integers(1, 3000)), int(r.integers(1, 5000))
            ss, ds = rb + int(r.integers(0, 64)), rb + int(r.integers(0, 64))
            src2 = r.integers(0, 256, rows * ss, dtype=np.uint8)
            dst2 = np.full(rows * ds, 7, np.uint8)
            copy(dst2.ctypes.data, ds, src2.ctypes.data, ss, rb, rows)               # strided rows
            a = src2.reshape(rows, ss)[:, :rb]
            b = dst2.reshape(rows, ds)
            if not (np.array_equal(a, b[:, :rb]) and (b[:, rb:] == 7).all()):
                errors.append(("strided", seed, it))

    errors = []
    work(int(rng.integers(1 << 30)), errors)
    threads = [threading.Thread(target=work, args=(100 + i, errors)) for i in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors

