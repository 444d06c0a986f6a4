"""NCC template scan: Python mirror of the reference's host interface over the C ABI.

Names follow ncc.rs: `Searcher` (ncc.rs:128-141,230-404), `get_hits` (ncc.rs:544-721),
`process_hits` / `partition_by` (ncc.rs:723-786, 1036-1052).  The compute always happens in
libfocr_b200.so (CUDA, sm_100a); there is no CPU path here.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import native
from .native import MATCH_DTYPE, check, lib, ptr

MAX_MATCHES = 1024  # ncc.rs:31


class Pinned:
    """A numpy array page-locked through the library (include/focr_b200.h: focr_pin_register)."""

    def __init__(self, ctx: "Context", a: np.ndarray):
        if not a.flags.c_contiguous:
            raise ValueError("pin: the array must be C-contiguous")
        self.ctx, self.array = ctx, a
        check(lib().focr_pin_register(ctx._h, ptr(a), a.nbytes))
        self._live = True

    def release(self):
        if self._live:
            self._live = False
            check(lib().focr_pin_unregister(self.ctx._h, ptr(self.array)))

    def __enter__(self):
        return self.array

    def __exit__(self, *exc):
        self.release()


class Context:
    """focr_ctx: one per GPU."""

    def __init__(self, device: int = 0, kernel: int = native.KERNEL_AUTO):
        self._h = C.c_void_p()
        check(lib().focr_ctx_create(device, C.byref(self._h)))
        self.device = device
        if kernel != native.KERNEL_AUTO:
            self.set_kernel(kernel)

    def set_kernel(self, kernel: int):
        check(lib().focr_ctx_set_kernel(self._h, kernel))

    @property
    def stream(self) -> int:
        return int(lib().focr_ctx_stream(self._h) or 0)

    def sync(self):
        check(lib().focr_ctx_sync(self._h))

    @property
    def launch_count(self) -> int:
        return int(lib().focr_ctx_launch_count(self._h))

    def profile(self, enable: bool = True):
        check(lib().focr_ctx_profile(self._h, int(enable)))

    def profile_read(self):
        """{stage: (total ms, launches)} since the last read; synchronises."""
        ms = np.zeros(6, np.float64)
        n = np.zeros(6, np.uint64)
        check(lib().focr_ctx_profile_read(self._h, ptr(ms), ptr(n)))
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(("invert", "stats", "scan", "finalize", "exact", "decode"))}

    def pin(self, a: np.ndarray) -> "Pinned":
        """Page-lock a caller-owned array in place (focr_pin_register): scans then DMA from / into it directly instead
        of staging it.  Use as a context manager, or call .release() when the array is no longer scanned."""
        return Pinned(self, a)

    def close(self):
        if self._h:
            lib().focr_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Bank:
    """focr_bank: the device-resident (glyph, subpixel shift) raster cache."""

    def __init__(self, ctx: Context, templates):
        """templates: sequence of u8 [n_h, n_w] arrays in the reference's (offset, letter) order."""
        tpls = [np.ascontiguousarray(t, np.uint8) for t in templates]
        self.ctx = ctx
        self.sizes = [(t.shape[1], t.shape[0]) for t in tpls]
        offsets = np.zeros(len(tpls), np.uint64)
        o = 0
        for i, t in enumerate(tpls):
            offsets[i] = o
            o += t.size
        pixels = np.concatenate([t.ravel() for t in tpls]) if tpls else np.zeros(0, np.uint8)
        n_w = np.array([s[0] for s in self.sizes], np.uint16)
        n_h = np.array([s[1] for s in self.sizes], np.uint16)
        self._h = C.c_void_p()
        check(lib().focr_bank_create(ctx._h, ptr(pixels), ptr(offsets), ptr(n_w), ptr(n_h), len(tpls),
                                     C.byref(self._h)))
        self.T = len(tpls)

    def close(self):
        if self._h:
            lib().focr_bank_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def scan_pages(ctx: Context, bank: Bank, pages: np.ndarray, threshold: float = 0.8, n_out: int = MAX_MATCHES,
               out: np.ndarray | None = None, counts: np.ndarray | None = None):
    """focr_ncc_scan: pages u8 [P, r_h, r_w] gray (host).  Returns (matches [P, T, n_out], counts [P, T])."""
    pages = np.ascontiguousarray(pages, np.uint8)
    if pages.ndim == 2:
        pages = pages[None]
    P, r_h, r_w = pages.shape
    if out is None:
        out = np.zeros((P, bank.T, n_out), MATCH_DTYPE)
    if counts is None:
        counts = np.zeros((P, bank.T), np.uint32)
    check(lib().focr_ncc_scan(ctx._h, bank._h, ptr(pages), r_w * r_h, r_w, r_h, P, C.c_float(threshold), n_out,
                              ptr(out), ptr(counts)))
    return out, counts


class MultiContext:
    """focr_multi: one process, one context per GPU (the reference's page-parallel driver, ncc.rs:839-847)."""

    def __init__(self, devices=None, n_devices: int = 0):
        """devices: explicit list of device indices; else devices 0..n_devices-1 (0 = every visible device)."""
        self._h = C.c_void_p()
        if devices is not None:
            arr = np.asarray(list(devices), np.int32)
            check(lib().focr_multi_create(ptr(arr), len(arr), C.byref(self._h)))
        else:
            check(lib().focr_multi_create(None, n_devices, C.byref(self._h)))
        self.size = int(lib().focr_multi_size(self._h))

    def page_block(self, n_pages: int, i: int):
        """(first page, pages) of device slot i for a batch of n_pages (== shard.shard_range)."""
        a, b = np.zeros(1, np.uint32), np.zeros(1, np.uint32)
        lib().focr_multi_page_block(self._h, n_pages, i, ptr(a), ptr(b))
        return int(a[0]), int(b[0])

    def sync(self):
        for i in range(self.size):
            check(lib().focr_ctx_sync(lib().focr_multi_ctx(self._h, i)))

    def close(self):
        if self._h:
            lib().focr_multi_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _pack_templates(templates):
    tpls = [np.ascontiguousarray(t, np.uint8) for t in templates]
    offsets = np.zeros(len(tpls), np.uint64)
    o = 0
    for i, t in enumerate(tpls):
        offsets[i] = o
        o += t.size
    pixels = np.concatenate([t.ravel() for t in tpls]) if tpls else np.zeros(0, np.uint8)
    n_w = np.array([t.shape[1] for t in tpls], np.uint16)
    n_h = np.array([t.shape[0] for t in tpls], np.uint16)
    return pixels, offsets, n_w, n_h


class MultiBank:
    """focr_multi_bank: the template bank replicated to every device of a MultiContext."""

    def __init__(self, mctx: MultiContext, templates):
        self.mctx = mctx
        pixels, offsets, n_w, n_h = _pack_templates(templates)
        self.T = len(offsets)
        self._h = C.c_void_p()
        check(lib().focr_multi_bank_create(mctx._h, ptr(pixels), ptr(offsets), ptr(n_w), ptr(n_h), self.T, C.byref(self._h)))

    def close(self):
        if self._h:
            lib().focr_multi_bank_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def scan_pages_multi(mctx: MultiContext, bank: MultiBank, pages: np.ndarray, threshold: float = 0.8, n_out: int = MAX_MATCHES,
                     out: np.ndarray | None = None, counts: np.ndarray | None = None):
    """focr_multi_ncc_scan: ONE batch sharded by page over the devices of `mctx`; same result layout as scan_pages."""
    pages = np.ascontiguousarray(pages, np.uint8)
    if pages.ndim == 2:
        pages = pages[None]
    P, r_h, r_w = pages.shape
    if out is None:
        out = np.zeros((P, bank.T, n_out), MATCH_DTYPE)
    if counts is None:
        counts = np.zeros((P, bank.T), np.uint32)
    check(lib().focr_multi_ncc_scan(mctx._h, bank._h, ptr(pages), r_w * r_h, r_w, r_h, P, C.c_float(threshold), n_out,
                                    ptr(out), ptr(counts)))
    return out, counts


def scan_pages_device(ctx: Context, bank: Bank, pages_ptr: int, page_stride: int, pitch: int, r_w: int, r_h: int,
                      n_pages: int, threshold: float, n_out: int, out_ptr: int, counts_ptr: int):
    """focr_ncc_scan_device on raw device addresses (e.g. torch tensors' data_ptr())."""
    check(lib().focr_ncc_scan_device(ctx._h, bank._h, ptr(pages_ptr), page_stride, pitch, r_w, r_h, n_pages,
                                     C.c_float(threshold), n_out, ptr(out_ptr), ptr(counts_ptr)))


def window_stats(ctx: Context, page: np.ndarray, n_w: int, n_h: int):
    page = np.ascontiguousarray(page, np.uint8)
    r_h, r_w = page.shape
    sp = np.zeros((r_h, r_w), np.uint32)
    s2 = np.zeros((r_h, r_w), np.uint64)
    rn = np.zeros((r_h, r_w), np.float64)
    check(lib().focr_window_stats(ctx._h, ptr(page), r_w, r_h, n_w, n_h, ptr(sp), ptr(s2), ptr(rn)))
    return sp, s2, rn


def numerators(ctx: Context, bank: Bank, t: int, page: np.ndarray):
    page = np.ascontiguousarray(page, np.uint8)
    r_h, r_w = page.shape
    acc = np.zeros((r_h, r_w), np.uint32)
    check(lib().focr_ncc_numerators(ctx._h, bank._h, t, ptr(page), r_w, r_h, ptr(acc)))
    return acc


class Searcher:
    """ncc.rs:128-141 `Searcher` through the compat shim: `search_c_u8` marshals exactly like
    ncc.rs:332-404 and calls the library's `ncc_8_u8` / `ncc_16_u8` symbols."""

    def __init__(self, gray: np.ndarray):
        gray = np.ascontiguousarray(gray, np.uint8)
        self.r_h, self.r_w = gray.shape
        self.reference_u8 = (255 - gray).astype(np.uint8)  # image_to_u8, ncc.rs:887-892
        self.matches_c = np.zeros(MAX_MATCHES, MATCH_DTYPE)

    def search_c_u8(self, needle: np.ndarray, threshold: float):
        n_h, n_w = needle.shape
        if n_w <= 8:
            N, fn = 8, lib().ncc_8_u8
        elif n_w <= 16:
            N, fn = 16, lib().ncc_16_u8
        else:
            raise NotImplementedError("not handled")  # ncc.rs:392
        padded = np.zeros((n_h, N), np.uint8)          # copy_needle_n_u8, ncc.rs:925-935
        padded[:, :n_w] = needle
        n = fn(ptr(self.reference_u8), self.r_w, self.r_h, ptr(padded), n_w, n_h, None, 0, None, None, None,
               C.c_float(threshold), ptr(self.matches_c), MAX_MATCHES)
        return self.matches_c[:n].copy()


def partition_by(xs, pred):
    """ncc.rs:1036-1052: groups are anchored to their FIRST element; panics on empty input."""
    if len(xs) == 0:
        raise IndexError("partition_by: empty input (the reference panics here, ncc.rs:1040)")
    i = j = 0
    last = xs[0]
    out = []
    for nxt in xs[1:]:
        j += 1
        if not pred(last, nxt):
            out.append((i, j))
            i = j
            last = nxt
    out.append((i, j + 1))
    return out


def get_hits(matches: np.ndarray, counts: np.ndarray, letters):
    """Flatten one page's scan result into the reference's all_hits order (ncc.rs:675-681):
    (template order, y, x) -> list of (letter, x, y, similarity)."""
    out = []
    for t, letter in enumerate(letters):
        for m in matches[t, : counts[t]]:
            out.append((letter, int(m["x"]), int(m["y"]), np.float32(m["similarity"])))
    return out


def process_hits(all_hits, anchor_threshold: float = 0.95, overlap: int = 5):
    """ncc.rs:723-786 on (letter, x, y, similarity) tuples; returns lines of the same tuples."""
    at = np.float32(anchor_threshold)
    keep_y = {h[2] for h in all_hits if h[3] >= at}
    hits = [h for h in all_hits if h[2] in keep_y]
    hits.sort(key=lambda h: h[2])
    slices = partition_by(hits, lambda a, b: a[2] == b[2])
    for i, j in slices:
        hits[i:j] = sorted(hits[i:j], key=lambda h: h[1])
    lines = []
    for i, j in slices:
        sl = hits[i:j]
        dedup = []
        for a, b in partition_by(sl, lambda p, q: abs(p[1] - q[1]) <= overlap):
            best = sl[a]
            for h in sl[a + 1:b]:
                if h[3] >= best[3]:  # Iterator::max_by keeps the LAST maximum
                    best = h
            dedup.append(best)
        lines.append(dedup)
    return lines


def process_hits_device(ctx, matches_dev, counts_dev, T: int, n_out: int, n_pages: int, letters,
                        anchor_threshold: float = 0.95, overlap: int = 5, raw: bool = False, only_page=None):
    """process_hits (ncc.rs:723-786) on the GPU, on match lists that are still resident in HBM (the buffers
    scan_pages_device filled).  Returns, per page, the same lines of (letter, x, y, similarity) tuples as
    process_hits(get_hits(...)); a page without any anchor line gets [] (the reference panics there, ncc.rs:1040).
    raw=True returns the C arrays (line_page, line_start, sel_tpl, sel) instead; only_page unpacks one page only."""
    n_lines, n_sel = np.zeros(1, np.uint32), np.zeros(1, np.uint32)
    line_cap, sel_cap = 4096 * n_pages, 65536 * n_pages
    while True:
        line_page = np.zeros(line_cap, np.uint32)
        line_start = np.zeros(line_cap + 1, np.uint32)
        sel_tpl = np.zeros(sel_cap, np.uint32)
        sel = np.zeros(sel_cap, MATCH_DTYPE)
        rc = lib().focr_process_hits_device(ctx._h, ptr(matches_dev), ptr(counts_dev), T, n_out, n_pages,
                                            C.c_float(anchor_threshold), overlap, line_cap, sel_cap, ptr(n_lines), ptr(n_sel),
                                            ptr(line_page), ptr(line_start), ptr(sel_tpl), ptr(sel))
        if rc == native.FOCR_ERR_NOMEM and (n_lines[0] > line_cap or n_sel[0] > sel_cap):
            line_cap, sel_cap = max(line_cap, int(n_lines[0])), max(sel_cap, int(n_sel[0]))
            continue
        check(rc)
        break
    nl, ns = int(n_lines[0]), int(n_sel[0])
    if raw:
        return line_page[:nl], line_start[:nl + 1], sel_tpl[:ns], sel[:ns]
    pages = [[] for _ in range(n_pages)]
    for l in range(nl):
        if only_page is not None and int(line_page[l]) != only_page:
            continue
        a, b = int(line_start[l]), int(line_start[l + 1])
        pages[int(line_page[l])].append([(letters[int(sel_tpl[k])], int(sel[k]["x"]), int(sel[k]["y"]),
                                          np.float32(sel[k]["similarity"])) for k in range(a, b)])
    return pages


def lines_to_text(lines):
    return ["".join(h[0] for h in line) for line in lines]


def lines_to_text_with_spaces(lines, advance_px: dict, space_px: float):
    """EXTENSION (opt-in; README.md:46 "does not currently detect spaces"): like lines_to_text, with spaces inserted where
    the pen travel between two kept hits exceeds the left glyph's advance, through the C++ host mirror
    (focr_host::line_text_with_spaces).  advance_px: {letter: advance in pixels}; space_px: advance of ' '."""
    letters = np.array([ord(c) for c in advance_px], np.uint32)
    adv = np.array([advance_px[c] for c in advance_px], np.float32)
    out = []
    for line in lines:
        xs = np.array([h[1] for h in line], np.int32)
        ls = np.array([ord(h[0]) for h in line], np.uint32)
        cap = 8 * len(line) + 64
        while True:
            buf, n_out = np.zeros(cap, np.uint32), np.zeros(1, np.uint32)
            rc = lib().focr_host_line_text_with_spaces(ptr(xs), ptr(ls), len(line), ptr(letters), ptr(adv), len(letters),
                                                       C.c_float(space_px), ptr(buf), cap, ptr(n_out))
            if rc == native.FOCR_ERR_NOMEM and int(n_out[0]) > cap:
                cap = int(n_out[0])
                continue
            check(rc)
            break
        out.append("".join(chr(c) for c in buf[:int(n_out[0])]))
    return out


def host_process_hits(all_hits, anchor_threshold: float = 0.95, overlap: int = 5):
    """process_hits through the C++ host mirror (host/focr_host.cpp); same input/output as process_hits.
    Raises IndexError where the reference panics (no anchor line, ncc.rs:1040)."""
    n = len(all_hits)
    xs = np.array([h[1] for h in all_hits], np.int32)
    ys = np.array([h[2] for h in all_hits], np.int32)
    sims = np.array([h[3] for h in all_hits], np.float32)
    letters = np.array([ord(h[0]) for h in all_hits], np.uint32)
    out_index = np.zeros(max(n, 1), np.uint32)
    line_offsets = np.zeros(n + 2, np.uint32)
    n_lines = np.zeros(1, np.uint32)
    rc = lib().focr_host_process_hits(ptr(xs), ptr(ys), ptr(sims), ptr(letters), n, C.c_float(anchor_threshold),
                                      overlap, ptr(out_index), ptr(line_offsets), ptr(n_lines))
    if rc != native.FOCR_OK:
        msg = lib().focr_last_error().decode()
        if msg.startswith("panic:"):
            raise IndexError(msg)
        raise native.FocrError(rc, msg)
    return [[all_hits[i] for i in out_index[line_offsets[l]:line_offsets[l + 1]]] for l in range(int(n_lines[0]))]


def host_search_c_u8(gray: np.ndarray, needle: np.ndarray, threshold: float):
    """Searcher::new + search_c_u8 through the C++ host mirror."""
    gray = np.ascontiguousarray(gray, np.uint8)
    needle = np.ascontiguousarray(needle, np.uint8)
    out = np.zeros(MAX_MATCHES, MATCH_DTYPE)
    n = np.zeros(1, np.uint32)
    rc = lib().focr_host_search_c_u8(ptr(gray), gray.shape[1], gray.shape[0], ptr(needle), needle.shape[1],
                                     needle.shape[0], C.c_float(threshold), ptr(out), ptr(n))
    if rc == native.FOCR_ERR_UNSUPPORTED:
        raise NotImplementedError(lib().focr_last_error().decode())
    check(rc)
    return out[:n[0]].copy()
