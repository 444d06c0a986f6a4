"""ctypes binding of libfocr_b200.so (include/focr_b200.h).  No fallback: if the library is missing
or there is no CUDA device the calls raise."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# FOCR_B200_LIB: tools/ load the experiments build (libfocr_b200_exp.so, see build.py) through this; the default is the product
LIB_PATH = os.environ.get("FOCR_B200_LIB") or os.path.join(HERE, "libfocr_b200.so")

MATCH_DTYPE = np.dtype([("x", np.uint16), ("y", np.uint16), ("similarity", np.float32)])  # focr_match
RASTER_DTYPE = np.dtype([("offset", np.uint64), ("left", np.int16), ("top", np.int16),
                         ("w", np.uint16), ("h", np.uint16)])                                # focr_glyph_raster
FOCR_OK, FOCR_ERR_CUDA, FOCR_ERR_ARG, FOCR_ERR_UNSUPPORTED, FOCR_ERR_NOMEM = range(5)
KERNEL_AUTO, KERNEL_SIMT, KERNEL_TCGEN05 = 0, 1, 2

# every symbol include/focr_b200.h declares (tests/test_abi.py checks the .so exports them all)
SYMBOLS = [
    "ncc_8_u8", "ncc_16_u8",
    "focr_version", "focr_last_error", "focr_get_limits",
    "focr_ctx_create", "focr_ctx_destroy", "focr_ctx_set_kernel", "focr_ctx_stream", "focr_ctx_sync",
    "focr_ctx_launch_count", "focr_ctx_profile", "focr_ctx_profile_read",
    "focr_pin_register", "focr_pin_unregister", "focr_pin_alloc", "focr_pin_free",
    "focr_bank_create", "focr_bank_destroy", "focr_bank_size",
    "focr_multi_create", "focr_multi_destroy", "focr_multi_size", "focr_multi_ctx", "focr_multi_page_block",
    "focr_multi_bank_create", "focr_multi_bank_destroy", "focr_multi_ncc_scan",
    "focr_multi_glyph_bank_create", "focr_multi_glyph_bank_destroy", "focr_multi_decode_pages",
    "focr_ncc_scan", "focr_ncc_scan_device", "focr_process_hits_device", "focr_window_stats", "focr_ncc_numerators",
    "focr_glyph_bank_create", "focr_glyph_bank_destroy", "focr_decode_pages", "focr_sum_of_squares",
    "focr_host_process_hits", "focr_host_search_c_u8", "focr_host_line_text_with_spaces",
    "focr_host_font_open", "focr_host_font_close", "focr_host_font_set_hinting", "focr_host_font_glyph_metrics", "focr_host_tbank_render", "focr_host_tbank_count", "focr_host_tbank_pixel_bytes",
    "focr_host_tbank_get", "focr_host_tbank_free", "focr_host_gbank_render", "focr_host_gbank_pixel_bytes", "focr_host_gbank_get",
    "focr_host_gbank_free",
]


class FocrError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libfocr_b200 error {code}: {msg}")
        self.code = code


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python font-ocr_b200/build.py` "
                          "(there is no CPU fallback)")
    l = C.CDLL(LIB_PATH)
    vp, sz, u32, u64 = C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint64
    shim = [vp, sz, sz, vp, sz, sz, vp, sz, vp, vp, vp, C.c_float, vp, sz]
    for n in ("ncc_8_u8", "ncc_16_u8"):
        getattr(l, n).restype, getattr(l, n).argtypes = sz, shim
    l.focr_version.restype = C.c_char_p
    l.focr_last_error.restype = C.c_char_p
    l.focr_get_limits.argtypes = [vp]
    l.focr_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    l.focr_ctx_destroy.argtypes = [vp]
    l.focr_ctx_destroy.restype = None
    l.focr_ctx_set_kernel.argtypes = [vp, C.c_int]
    l.focr_ctx_stream.argtypes = [vp]
    l.focr_ctx_stream.restype = vp
    l.focr_ctx_sync.argtypes = [vp]
    l.focr_ctx_launch_count.argtypes = [vp]
    l.focr_ctx_launch_count.restype = u64
    l.focr_ctx_profile.argtypes = [vp, C.c_int]
    l.focr_ctx_profile_read.argtypes = [vp, vp, vp]
    l.focr_pin_register.argtypes = [vp, vp, sz]
    l.focr_pin_unregister.argtypes = [vp, vp]
    l.focr_pin_alloc.argtypes = [vp, sz, C.POINTER(vp)]
    l.focr_pin_free.argtypes = [vp, vp]
    l.focr_bank_create.argtypes = [vp, vp, vp, vp, vp, u32, C.POINTER(vp)]
    l.focr_bank_destroy.argtypes = [vp]
    l.focr_bank_destroy.restype = None
    l.focr_bank_size.argtypes = [vp]
    l.focr_bank_size.restype = u32
    l.focr_ncc_scan.argtypes = [vp, vp, vp, sz, u32, u32, u32, C.c_float, u32, vp, vp]
    l.focr_ncc_scan_device.argtypes = [vp, vp, vp, sz, sz, u32, u32, u32, C.c_float, u32, vp, vp]
    l.focr_multi_create.argtypes = [vp, u32, C.POINTER(vp)]
    l.focr_multi_destroy.argtypes = [vp]
    l.focr_multi_destroy.restype = None
    l.focr_multi_size.argtypes = [vp]
    l.focr_multi_size.restype = u32
    l.focr_multi_ctx.argtypes = [vp, u32]
    l.focr_multi_ctx.restype = vp
    l.focr_multi_page_block.argtypes = [vp, u32, u32, vp, vp]
    l.focr_multi_page_block.restype = None
    l.focr_multi_bank_create.argtypes = [vp, vp, vp, vp, vp, u32, C.POINTER(vp)]
    l.focr_multi_bank_destroy.argtypes = [vp]
    l.focr_multi_bank_destroy.restype = None
    l.focr_multi_ncc_scan.argtypes = [vp, vp, vp, sz, u32, u32, u32, C.c_float, u32, vp, vp]
    l.focr_multi_glyph_bank_create.argtypes = [vp, vp, sz, vp, vp, u32, C.c_int32, C.POINTER(vp)]
    l.focr_multi_glyph_bank_destroy.argtypes = [vp]
    l.focr_multi_glyph_bank_destroy.restype = None
    l.focr_multi_decode_pages.argtypes = [vp, vp, vp, sz] + [u32] * 10 + [vp, vp, vp, vp]
    l.focr_window_stats.argtypes = [vp, vp, u32, u32, u32, u32, vp, vp, vp]
    l.focr_ncc_numerators.argtypes = [vp, vp, u32, vp, u32, u32, vp]
    l.focr_glyph_bank_create.argtypes = [vp, vp, sz, vp, vp, u32, C.c_int32, C.POINTER(vp)]
    l.focr_glyph_bank_destroy.argtypes = [vp]
    l.focr_glyph_bank_destroy.restype = None
    l.focr_decode_pages.argtypes = [vp, vp, vp, sz] + [u32] * 10 + [vp, vp, vp, vp]
    l.focr_sum_of_squares.argtypes = [vp, vp, vp, sz, u32, vp]
    l.focr_process_hits_device.argtypes = [vp, vp, vp, u32, u32, u32, C.c_float, C.c_int32, u32, u32, vp, vp, vp, vp, vp, vp]
    l.focr_host_process_hits.argtypes = [vp, vp, vp, vp, u32, C.c_float, C.c_int32, vp, vp, vp]
    l.focr_host_line_text_with_spaces.argtypes = [vp, vp, u32, vp, vp, u32, C.c_float, vp, u32, vp]
    l.focr_host_font_open.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(vp)]
    l.focr_host_font_glyph_metrics.argtypes = [vp, u32, C.c_float, vp, vp]
    l.focr_host_font_set_hinting.argtypes = [vp, C.c_int]
    l.focr_host_font_set_hinting.restype = None
    l.focr_host_font_close.argtypes = [vp]
    l.focr_host_font_close.restype = None
    l.focr_host_tbank_render.argtypes = [vp, C.c_float, vp, u32, u32, u32, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    l.focr_host_tbank_count.argtypes = [vp]
    l.focr_host_tbank_count.restype = u32
    l.focr_host_tbank_pixel_bytes.argtypes = [vp]
    l.focr_host_tbank_pixel_bytes.restype = u64
    l.focr_host_tbank_get.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    l.focr_host_tbank_free.argtypes = [vp]
    l.focr_host_tbank_free.restype = None
    l.focr_host_gbank_render.argtypes = [vp, C.c_float, vp, u32, C.c_float, C.POINTER(vp)]
    l.focr_host_gbank_pixel_bytes.argtypes = [vp]
    l.focr_host_gbank_pixel_bytes.restype = u64
    l.focr_host_gbank_get.argtypes = [vp, vp, vp, vp, vp]
    l.focr_host_gbank_free.argtypes = [vp]
    l.focr_host_gbank_free.restype = None
    l.focr_host_search_c_u8.argtypes = [vp, u32, u32, vp, u32, u32, C.c_float, vp, vp]
    _lib = l
    return l


def check(rc: int):
    if rc != FOCR_OK:
        raise FocrError(rc, lib().focr_last_error().decode())


def ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(int(a))  # raw address (e.g. torch.Tensor.data_ptr())
