"""focr least-squared-distance line decode: Python mirror of the reference's host interface
(main.rs:87-239) over the C ABI.  `GlyphBank` is the (glyph, subpixel shift) raster cache the
README asks for (README.md:44): the reference re-rasterises every alphabet glyph for every cell."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import native
from .native import RASTER_DTYPE, check, lib, ptr
from .raster import FOCR_DEFAULT_ALPHABET, Font, f32


class GlyphBank:
    """focr_glyph_bank: 64 horizontal 26.6 phases of every alphabet glyph + f32 advances.
    `ctx` is an ncc.Context (one GPU) or an ncc.MultiContext (the bank is replicated to every device)."""

    def __init__(self, ctx, font: Font, size: float, alphabet: str = FOCR_DEFAULT_ALPHABET, kern_x: float = 1.0):
        self.ctx, self.font, self.alphabet, self.size = ctx, font, alphabet, size
        self.multi = hasattr(ctx, "page_block")
        gids = [font.glyph_for_char(c) for c in alphabet]                       # main.rs:125-128
        x0 = y0 = 0                                                             # RectF::default() seeds the union
        for gid in gids:                                                        # main.rs:133-146
            a, b, _, _ = font.raster_bounds(gid, size, 0.0, 0.0)
            x0, y0 = min(x0, a), min(y0, b)
        self.origin = (-x0, -y0)                                                # main.rs:147
        upem = f32(font.units_per_em)
        self.advance_px = np.array(
            [f32(f32(f32(font.advance(g)[0] / upem) * f32(size)) * f32(kern_x)) for g in gids], np.float32)
        rasters = np.zeros((len(gids), 64), RASTER_DTYPE)
        chunks, off = [], 0
        for gi, gid in enumerate(gids):
            for s in range(64):
                bmp, left, top = font.glyph_bitmap(gid, size, s, -self.origin[1] * 64)
                h, w = bmp.shape if bmp.size else (0, 0)
                rasters[gi, s] = (off, left, -top, w, h)
                if bmp.size:
                    chunks.append(bmp.ravel())
                    off += bmp.size
        pixels = np.concatenate(chunks) if chunks else np.zeros(1, np.uint8)
        self._h = C.c_void_p()
        create = lib().focr_multi_glyph_bank_create if self.multi else lib().focr_glyph_bank_create
        check(create(ctx._h, ptr(pixels), pixels.size, ptr(rasters), ptr(self.advance_px),
                     len(gids), int(self.origin[0]), C.byref(self._h)))

    def close(self):
        if self._h:
            (lib().focr_multi_glyph_bank_destroy if self.multi else lib().focr_glyph_bank_destroy)(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def decode_images(ctx, bank: GlyphBank, pages: np.ndarray, x_start: int, y_start: int, width: int,
                  line_height: int, line_advance: int, max_cells: int = 512):
    """decode_image_vec (main.rs:220-239) for a batch: returns per page a list of (text, y)."""
    pages = np.ascontiguousarray(pages, np.uint8)
    if pages.ndim == 2:
        pages = pages[None]
    P, r_h, r_w = pages.shape
    max_lines = max((max(r_h - y_start, 0) + line_advance - 1) // line_advance, 1)
    glyphs = np.zeros((P, max_lines, max_cells), np.uint16)
    n_cells = np.zeros((P, max_lines), np.uint32)
    line_y = np.zeros((P, max_lines), np.uint32)
    n_lines = np.zeros(P, np.uint32)
    decode = lib().focr_multi_decode_pages if bank.multi else lib().focr_decode_pages   # pages sharded over the GPUs
    check(decode(ctx._h, bank._h, ptr(pages), r_w * r_h, r_w, r_h, P, x_start, y_start, width,
                 line_height, line_advance, max_lines, max_cells, ptr(glyphs), ptr(n_cells),
                 ptr(line_y), ptr(n_lines)))
    out = []
    for p in range(P):
        out.append([("".join(bank.alphabet[g] for g in glyphs[p, l, :n_cells[p, l]]), int(line_y[p, l]))
                    for l in range(n_lines[p])])
    return out


def sum_of_squares(ctx, xs: np.ndarray, ys: np.ndarray):
    """main.rs:510-516 for a batch of equal-length u8 strips [n, len] -> i64 [n]."""
    xs = np.ascontiguousarray(xs, np.uint8)
    ys = np.ascontiguousarray(ys, np.uint8)
    if xs.ndim == 1:
        xs, ys = xs[None], ys[None]
    assert xs.shape == ys.shape
    out = np.zeros(xs.shape[0], np.int64)
    check(lib().focr_sum_of_squares(ctx._h, ptr(xs), ptr(ys), xs.shape[1], xs.shape[0], ptr(out)))
    return out
