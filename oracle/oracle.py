"""oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT.

CPU oracle for the font-ocr hot path.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; nothing under font-ocr_b200/ does.

Three layers:
  * `impl="reference"`: the reference's own AVX2 kernel compiled from /root/reference/src/ncc.cpp
    into oracle/_ref/libncc_ref.so (oracle/Makefile), driven through the FFI signature of
    ncc.rs:92-126 with the window statistics prepared by the restated `prepare_for_size`.
  * `impl="port"`: oracle/ncc_oracle.c, our plain-C restatement of the same algorithm.
  * `brute_force`: an independent numpy/f64 evaluation of the NCC definition.
Plus the restated Rust host logic the kernel sits in: get_hits ordering (ncc.rs:587-702),
process_hits / partition_by (ncc.rs:723-786, 1036-1052), and focr's decode_line / decode_image
(main.rs:87-218).

Parity pin: the reference has no tests or golden vectors.  The port is pinned to the
compiled reference kernel in tests/test_oracle.py (bit-identical match lists) and to the fixtures
in tests/golden/ which that compiled kernel generated (tests/golden/make_golden.py).  The Rust host
logic and everything produced by font-kit/FreeType cannot be executed here (no Rust toolchain):
for those rows parity is UNPINNED and rests on the line-by-line restatement cited below.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
MAX_MATCHES = 1024  # ncc.rs:31

MATCH_DTYPE = np.dtype([("x", np.uint16), ("y", np.uint16), ("similarity", np.float32)])  # ncc.cpp:7-10
PAGE_PAD = 64  # the reference over-reads up to 16-n_w bytes past the page

_port = None
_ref = None


def build(quiet: bool = True):
    """Compile the C restatement (and, when /root/reference is present, the reference kernel)."""
    subprocess.run(["make", "-C", _HERE], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def port_lib():
    global _port
    if _port is None:
        path = os.path.join(_HERE, "_build", "libncc_oracle.so")
        if not os.path.exists(path):
            build()
        lib = C.CDLL(path)
        lib.orc_ncc_scan.restype = C.c_size_t
        lib.orc_ncc_scan.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t,
                                     C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p,
                                     C.c_size_t, C.c_void_p]
        lib.orc_prepare_for_size.restype = C.c_int
        lib.orc_prepare_for_size.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t,
                                             C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.orc_sum_table.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
        lib.orc_sumsqr_table.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
        lib.orc_copy_needle.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p]
        lib.orc_image_to_u8.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        lib.orc_sum_of_squares.restype = C.c_int64
        lib.orc_sum_of_squares.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        lib.orc_decode_line_cached.restype = C.c_size_t
        lib.orc_decode_line_cached.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p,
                                               C.c_void_p, C.c_size_t, C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p]
        lib.orc_decode_image_cached.restype = C.c_size_t
        lib.orc_decode_image_cached.argtypes = [C.c_void_p] + [C.c_size_t] * 7 + [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                                                                  C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                                                                  C.c_size_t, C.c_size_t, C.c_void_p]
        _port = lib
    return _port


def ref_lib():
    """The compiled, unmodified reference kernel, or None when it was never built."""
    global _ref
    if _ref is None:
        path = os.path.join(_HERE, "_ref", "libncc_ref.so")
        if not os.path.exists(path):
            return None
        lib = C.CDLL(path)
        sig = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t,
               C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_size_t]  # ncc.rs:93-125
        for name in ("ncc_8_u8", "ncc_16_u8"):
            getattr(lib, name).restype = C.c_size_t
            getattr(lib, name).argtypes = sig
        _ref = lib
    return _ref


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class Searcher:
    """ncc.rs:128-141,230-404 `Searcher`, over the C restatement; `impl` picks which kernel scans."""

    def __init__(self, gray: np.ndarray, impl: str = "port"):
        assert gray.dtype == np.uint8 and gray.ndim == 2
        self.impl = impl
        self.r_h, self.r_w = gray.shape
        lib = port_lib()
        n = self.r_w * self.r_h
        self._buf = np.zeros(n + PAGE_PAD, np.uint8)
        g = np.ascontiguousarray(gray)
        lib.orc_image_to_u8(_p(g), n, _p(self._buf))               # ncc.rs:232
        self.reference_u8 = self._buf[:n].reshape(self.r_h, self.r_w)
        self.sum_table = np.zeros(n, np.uint32)
        self.sumsqr_table = np.zeros(n, np.uint64)
        lib.orc_sum_table(_p(self._buf), self.r_h, self.r_w, _p(self.sum_table))        # ncc.rs:233
        lib.orc_sumsqr_table(_p(self._buf), self.r_h, self.r_w, _p(self.sumsqr_table))  # ncc.rs:234
        self.patch_sum = np.zeros(n, np.uint32)
        self.patch_rnorm = np.zeros(n, np.float64)
        self.start_end = np.zeros(self.r_h * 2, np.uint16)
        self.acc_u32 = np.zeros(self.r_w * 8 + 8, np.uint32)       # ncc.rs:242
        self.matches_c = np.zeros(MAX_MATCHES, MATCH_DTYPE)        # ncc.rs:240
        self.last_patch_size = None

    def prepare_for_size(self, n_w: int, n_h: int):
        if self.last_patch_size == (n_w, n_h):                     # ncc.rs:264-268
            return
        rc = port_lib().orc_prepare_for_size(_p(self.sum_table), _p(self.sumsqr_table), self.r_w, self.r_h,
                                             n_w, n_h, _p(self.patch_sum), _p(self.patch_rnorm),
                                             _p(self.start_end))
        if rc != 0:
            raise OverflowError("start/end does not fit u16 (ncc.rs:313-314 panics)")
        self.last_patch_size = (n_w, n_h)

    def search_c_u8(self, needle: np.ndarray, threshold: float, n_out: int = MAX_MATCHES,
                    want_acc: bool = False, allow_wide: bool = False):
        """ncc.rs:332-404.  needle: u8 [n_h, n_w].  Returns a MATCH_DTYPE array (and the acc plane)."""
        n_h, n_w = needle.shape
        self.prepare_for_size(n_w, n_h)
        if n_w <= 8:
            N = 8
        elif n_w <= 16:
            N = 16
        elif allow_wide:
            N = (n_w + 15) // 16 * 16
        else:
            raise NotImplementedError("not handled")               # ncc.rs:392 panic!
        lib = port_lib()
        padded = np.zeros((n_h, N), np.uint8)
        lib.orc_copy_needle(_p(np.ascontiguousarray(needle)), n_w, n_h, N, _p(padded))
        out = self.matches_c if n_out == MAX_MATCHES else np.zeros(n_out, MATCH_DTYPE)
        acc_plane = np.zeros(self.r_w * self.r_h, np.uint32) if want_acc else None
        if self.impl == "reference" and N <= 16 and not want_acc:
            rl = ref_lib()
            if rl is None:
                raise FileNotFoundError("oracle/_ref/libncc_ref.so missing (run make -C oracle where /root/reference exists)")
            fn = rl.ncc_8_u8 if N == 8 else rl.ncc_16_u8
            cnt = fn(_p(self._buf), self.r_w, self.r_h, _p(padded), n_w, n_h, _p(self.acc_u32),
                     self.acc_u32.size, _p(self.patch_sum), _p(self.patch_rnorm), _p(self.start_end),
                     C.c_float(threshold), _p(out), n_out)
        else:
            cnt = lib.orc_ncc_scan(_p(self._buf), self.r_w, self.r_h, _p(padded), N, n_w, n_h,
                                   _p(self.patch_sum), _p(self.patch_rnorm), _p(self.start_end),
                                   C.c_float(threshold), _p(out), n_out,
                                   _p(acc_plane) if want_acc else None)
        res = out[:cnt].copy()
        if want_acc:
            return res, acc_plane.reshape(self.r_h, self.r_w)
        return res


def get_hits(gray: np.ndarray, templates, threshold: float, impl: str = "port", n_out: int = MAX_MATCHES,
             allow_wide: bool = False):
    """ncc.rs:544-721 restricted to the scan: templates in bank order (offset index, alphabet index);
    returns a list (one MATCH_DTYPE array per template) -- all_hits order is (template, y, x)."""
    s = Searcher(gray, impl)
    return [s.search_c_u8(np.asarray(t), threshold, n_out, allow_wide=allow_wide) for t in templates]


# --------------------------------------------------------------------------- independent check
def window_sums(inv: np.ndarray, n_w: int, n_h: int):
    """Exact window sum and sum of squares for every (y, x) (int64 [y_searches, x_searches])."""
    p = inv.astype(np.int64)
    def box(a):
        c = np.zeros((a.shape[0] + 1, a.shape[1] + 1), np.int64)
        c[1:, 1:] = a.cumsum(0).cumsum(1)
        return c[n_h:, n_w:] - c[:-n_h, n_w:] - c[n_h:, :-n_w] + c[:-n_h, :-n_w]
    return box(p), box(p * p)


def brute_force(gray: np.ndarray, needle: np.ndarray, threshold: float, n_out: int = MAX_MATCHES):
    """NCC from its definition in numpy int64/f64: (matches, acc, s_p, s2_p).  Independent of both
    the reference's SAT bookkeeping and its SIMD layout; rows/cols 0 are masked like the reference."""
    inv = (255 - gray).astype(np.uint8)
    n_h, n_w = needle.shape
    n = n_w * n_h
    win = np.lib.stride_tricks.sliding_window_view(inv, (n_h, n_w))
    acc = np.einsum("yxij,ij->yx", win.astype(np.int64), needle.astype(np.int64))
    s_p, s2_p = window_sums(inv, n_w, n_h)
    s_n = int(needle.astype(np.int64).sum())
    s2_n = int((needle.astype(np.int64) ** 2).sum())
    with np.errstate(all="ignore"):
        num = acc - (s_n * s_p) * (1.0 / n)
        den = (1.0 / np.sqrt(s2_n - s_n * s_n / n)) * (1.0 / np.sqrt(s2_p - (s_p * s_p) / n))
        sim = num * den
    ok = np.isfinite(sim) & (sim > np.float64(np.float32(threshold)))
    ok[0, :] = False
    ok[:, 0] = False
    # the reference scans y in [1, y_searches) and x in [1, x_searches): the last row/col of `win`
    # (index y_searches-1 / x_searches-1) are included, index 0 is not.
    ys, xs = np.nonzero(ok)
    m = np.zeros(min(len(ys), n_out), MATCH_DTYPE)
    m["x"], m["y"], m["similarity"] = xs[:n_out], ys[:n_out], sim[ys, xs][:n_out].astype(np.float32)
    return m, acc, s_p, s2_p


# --------------------------------------------------------------------------- K12 post-processing
def partition_by(xs, pred):
    """ncc.rs:1036-1052.  `last` is the FIRST element of the current group; panics on empty input."""
    if len(xs) == 0:
        raise IndexError("partition_by on empty input (ncc.rs:1040 unwrap on None)")
    i = j = 0
    last = xs[0]
    slices = []
    for nxt in xs[1:]:
        j += 1
        if not pred(last, nxt):
            slices.append((i, j))
            i = j
            last = nxt
    slices.append((i, j + 1))
    return slices


def process_hits(all_hits, anchor_threshold: float = 0.95, overlap: int = 5):
    """ncc.rs:723-786.  all_hits: sequence of (letter, x, y, similarity_f32) in get_hits order.
    Returns lines: list of lists of the same tuples."""
    at = np.float32(anchor_threshold)
    keep_y = {h[2] for h in all_hits if np.float32(h[3]) >= at}                  # ncc.rs:727-731
    hits = [h for h in all_hits if h[2] in keep_y]                               # ncc.rs:732-738
    hits.sort(key=lambda h: h[2])                                                # stable, ncc.rs:741
    line_slices = partition_by(hits, lambda a, b: a[2] == b[2])                  # ncc.rs:747
    for i, j in line_slices:
        hits[i:j] = sorted(hits[i:j], key=lambda h: h[1])                        # stable, ncc.rs:749-752
    lines = []
    for i, j in line_slices:
        sl = hits[i:j]
        dedup = []
        for a, b in partition_by(sl, lambda p, q: abs(p[1] - q[1]) <= overlap):  # ncc.rs:755-757
            best = sl[a]
            for h in sl[a + 1:b]:                                                # max_by: LAST max wins
                if np.float32(h[3]) >= np.float32(best[3]):
                    best = h
            dedup.append(best)
        lines.append(dedup)
    return lines


def hits_with_letters(per_template, letters):
    """Flatten get_hits output into process_hits input, in the reference's push order (ncc.rs:675-681)."""
    out = []
    for ms, letter in zip(per_template, letters):
        for m in ms:
            out.append((letter, int(m["x"]), int(m["y"]), np.float32(m["similarity"])))
    return out


def lines_to_text(lines):
    return ["".join(h[0] for h in line) for line in lines]                       # ncc.rs:869-876


def lines_to_text_with_spaces(lines, advance_px: dict, space_px: float):
    """Restatement of the opt-in space-detection EXTENSION (not in the reference, README.md:46): between consecutive kept
    hits of a line, round((x_i - x_{i-1} - advance(letter_{i-1})) / space_px) spaces when that excess is positive."""
    f32 = np.float32
    out = []
    for line in lines:
        s = []
        for i, h in enumerate(line):
            if i and space_px > 0:
                excess = f32(f32(h[1] - line[i - 1][1]) - f32(advance_px.get(line[i - 1][0], 0.0)))
                if excess > 0:
                    s.append(" " * int(np.floor(f32(f32(excess / f32(space_px)) + f32(0.5)))))
            s.append(h[0])
        out.append("".join(s))
    return out


# --------------------------------------------------------------------------- focr (main.rs)
def sum_of_squares(xs: np.ndarray, ys: np.ndarray) -> int:
    """main.rs:510-516 via the C restatement."""
    xs = np.ascontiguousarray(xs, np.uint8).ravel()
    ys = np.ascontiguousarray(ys, np.uint8).ravel()
    assert xs.size == ys.size
    return int(port_lib().orc_sum_of_squares(_p(xs), _p(ys), xs.size))


def decode_line(strip_gray: np.ndarray, font, alphabet: str, size: float, kern_x: float = 1.0):
    """main.rs:112-181 with score_glyph (main.rs:87-110): `font` is a font_ocr_b200.raster.Font (the
    rasteriser is an input producer; the oracle calls it per cell exactly like the reference)."""
    f32 = np.float32
    h, w = strip_gray.shape
    canvas = np.zeros((h, w), np.uint8)
    upem = f32(font.units_per_em)
    gids = [(c, font.glyph_for_char(c)) for c in alphabet]
    x0 = y0 = 0                                                                  # RectF::default()
    for _, gid in gids:                                                          # main.rs:133-146
        a, b, _c, _d = font.raster_bounds(gid, size, 0.0, 0.0)
        x0, y0 = min(x0, a), min(y0, b)
    ox, oy = f32(-x0), f32(-y0)                                                  # main.rs:147
    ref = (255 - strip_gray).astype(np.uint8)                                    # main.rs:150
    pos = f32(0.0)
    s = []
    while pos < f32(w):                                                          # main.rs:158
        best, best_score = None, None
        for c, gid in gids:                                                      # min_by_key: FIRST min
            canvas[:] = 0
            font.rasterize_glyph(canvas, gid, size, f32(ox + pos), f32(oy + f32(0.0)))
            sc = sum_of_squares(ref, canvas)
            if best_score is None or sc < best_score:
                best, best_score = (c, gid), sc
        s.append(best[0])
        pos = f32(pos + f32(f32(f32(font.advance(best[1])[0] / upem) * f32(size)) * f32(kern_x)))
    return "".join(s)


def decode_image(gray: np.ndarray, font, alphabet: str, size: float, x_start: int, y_start: int, width: int,
                 line_height: int, line_advance: int, kern_x: float = 1.0, max_lines: int | None = None):
    """main.rs:183-218.  crop_imm clamps the rectangle to the image (image crate)."""
    H, W = gray.shape
    out = []
    i = 0
    while True:
        y = y_start + i * line_advance
        i += 1
        xs, ys = min(x_start, W), min(y, H)
        strip = gray[ys:min(ys + line_height, H), xs:min(xs + width, W)]
        if strip.shape[0] == 0:                                                  # main.rs:205-207
            break
        if (strip == 255).all():                                  # main.rs:208-211
            continue
        text = decode_line(strip, font, alphabet, size, kern_x)
        if text == "":                                                           # main.rs:213-215
            break
        out.append((text, y))
        if max_lines is not None and len(out) >= max_lines:
            break
    return out


# --------------------------------------------------------------------------- focr with cached glyph rasters
RASTER_DTYPE = np.dtype([("offset", np.uint64), ("left", np.int16), ("top", np.int16), ("w", np.uint16), ("h", np.uint16)])


class GlyphCache:
    """The (glyph, 26.6 phase) raster cache BASELINE config 4 names, built with the same producer calls decode_line makes
    per cell (main.rs:98-106): 64 horizontal phases per alphabet glyph, the vertical delta fixed by the origin
    (main.rs:147), and the f32 advances of main.rs:176-178."""

    def __init__(self, font, alphabet: str, size: float, kern_x: float = 1.0):
        f32 = np.float32
        self.alphabet = alphabet
        gids = [font.glyph_for_char(c) for c in alphabet]                       # main.rs:125-128
        x0 = y0 = 0                                                             # RectF::default()
        for gid in gids:                                                        # main.rs:133-146
            a, b, _c, _d = font.raster_bounds(gid, size, 0.0, 0.0)
            x0, y0 = min(x0, a), min(y0, b)
        self.origin_x, origin_y = -x0, -y0                                      # main.rs:147
        upem = f32(font.units_per_em)
        self.advance_px = np.array([f32(f32(f32(font.advance(g)[0] / upem) * f32(size)) * f32(kern_x)) for g in gids],
                                   np.float32)
        self.rasters = np.zeros((len(gids), 64), RASTER_DTYPE)
        chunks, off = [], 0
        for gi, gid in enumerate(gids):
            for ph in range(64):
                bmp, left, top = font.glyph_bitmap(gid, size, ph, -int(f32(f32(origin_y) * f32(64.0))))
                h, w = bmp.shape if bmp.size else (0, 0)
                self.rasters[gi, ph] = (off, left, -top, w, h)
                if bmp.size:
                    chunks.append(bmp.ravel())
                    off += bmp.size
        self.pixels = np.concatenate(chunks) if chunks else np.zeros(1, np.uint8)


def decode_image_cached(gray: np.ndarray, cache: GlyphCache, x_start: int, y_start: int, width: int, line_height: int,
                        line_advance: int, max_cells: int = 512):
    """main.rs:183-218 with score_glyph's rasterisation served from `cache` (C restatement): [(text, y)]."""
    gray = np.ascontiguousarray(gray, np.uint8)
    H, W = gray.shape
    max_lines = max((max(H - y_start, 0) + line_advance - 1) // line_advance, 1)
    glyphs = np.zeros((max_lines, max_cells), np.uint16)
    n_cells = np.zeros(max_lines, np.uint32)
    line_y = np.zeros(max_lines, np.uint32)
    scratch = np.zeros(2 * width * line_height + 64, np.uint8)
    n = port_lib().orc_decode_image_cached(_p(gray), W, H, x_start, y_start, width, line_height, line_advance,
                                           _p(cache.pixels), _p(cache.rasters), _p(cache.advance_px), len(cache.alphabet),
                                           int(cache.origin_x), _p(glyphs), _p(n_cells), _p(line_y), max_lines, max_cells,
                                           _p(scratch))
    if n == 2 ** 64 - 1:
        raise OverflowError("a line needs more than max_cells cells")
    return [("".join(cache.alphabet[g] for g in glyphs[l, :n_cells[l]]), int(line_y[l])) for l in range(n)]
