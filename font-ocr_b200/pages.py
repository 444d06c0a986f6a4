"""Synthetic page generator for the BASELINE.json configs.

There is no network (no Courier New, no scans), and the image only has proportional web fonts,
so pages are synthesised: base64 text of seeded random bytes, black ink on a white page.

* NCC pages (`make_ncc_page`) are laid out FIXED-PITCH: each character is the bank's own template
  for a (glyph, subpixel offset) stamped at an integer pen position, pitch = box width + 1, line
  advance = box height + 3.  `BoxSize::Alphabet` templates (ncc.rs:600-626) assume neighbours do
  not intrude into the box, which only holds for a monospace face; the fixed pitch restores that
  for Lato.  `shifts="bank"` draws a random subpixel offset per character from the bank's offset
  grid (text that a renderer grid-snapped, README.md "NCC" section); `shifts="zero"` uses offset
  (0,0) only.
* focr pages (`make_focr_page`) replay the reference's own text layout (main.rs:40-85 `render`):
  natural advances accumulated in f32, glyphs rasterised at the fractional pen position.
"""
from __future__ import annotations

import base64

import numpy as np

from .raster import Font, TemplateBank, _blit, f32


def base64_text(seed: int, n_chars: int) -> str:
    rng = np.random.default_rng(seed)
    raw = rng.bytes((n_chars * 3) // 4 + 3)
    return base64.b64encode(raw).decode()[:n_chars]


def make_ncc_page(bank: TemplateBank, width: int, height: int, seed: int = 0, margin_x: int = 45,
                  margin_y: int = 39, shifts: str = "zero", fill: float = 1.0):
    """Returns (gray u8 [height, width], lines: list[str], placements: list[(x, y, template_index)])."""
    A = len(bank.alphabet)
    n_off = len(bank) // A
    bw = max(t.pixels.shape[1] for t in bank.templates)
    bh = max(t.pixels.shape[0] for t in bank.templates)
    pitch, adv = bw + 1, bh + 3
    cols = max((width - 2 * margin_x) // pitch, 0)
    rows = max((height - 2 * margin_y) // adv, 0)
    n_lines = max(int(rows * fill), 1 if rows else 0)
    text = base64_text(seed, cols * n_lines)
    index = {c: i for i, c in enumerate(bank.alphabet)}
    rng = np.random.default_rng(seed + 0x5EED)
    ink = np.zeros((height, width), np.uint8)
    lines, placements = [], []
    for li in range(n_lines):
        line = text[li * cols:(li + 1) * cols]
        lines.append(line)
        y = margin_y + li * adv
        offs = rng.integers(0, n_off, size=len(line)) if shifts == "bank" else np.zeros(len(line), int)
        for ci, ch in enumerate(line):
            ti = int(offs[ci]) * A + index[ch]
            px = bank.templates[ti].pixels
            x = margin_x + ci * pitch
            h, w = px.shape
            np.maximum(ink[y:y + h, x:x + w], px, out=ink[y:y + h, x:x + w])
            placements.append((x, y, ti))
    return (255 - ink).astype(np.uint8), lines, placements


def focr_render_line(font: Font, text: str, size: float, kern_x: float = 1.0):
    """main.rs:40-85 `render`: (A8 canvas u8 [h, w], bounds origin (x, y) as ints)."""
    upem = f32(font.units_per_em)
    pos = f32(0.0)
    gp = []
    for ch in text:
        gid = font.glyph_for_char(ch)
        gp.append((gid, pos))
        pos = f32(pos + f32(f32(f32(font.advance(gid)[0] / upem) * f32(size)) * f32(kern_x)))
    x0 = y0 = x1 = y1 = 0  # RectF::default() seeds the union with (0,0)
    for gid, p in gp:
        a, b, c, d = font.raster_bounds(gid, size, p, 0.0)
        x0, y0, x1, y1 = min(x0, a), min(y0, b), max(x1, c), max(y1, d)
    canvas = np.zeros((y1 - y0, x1 - x0), np.uint8)
    for gid, p in gp:
        font.rasterize_glyph(canvas, gid, size, f32(f32(-x0) + p), f32(-y0))
    return canvas, (x0, y0)


def make_focr_page(font: Font, size: float, width: int, height: int, seed: int = 0, x_start: int = 45,
                   y_start: int = 39, line_width: int = 608, line_height: int = 12, line_advance: int = 15,
                   fill: float = 1.0, alphabet: str | None = None):
    """A page of base64 lines placed where focr's rectangles look (main.rs:199-203).  Line i is laid
    out with the reference's own pen arithmetic (main.rs:46-54: f32 advances) and rasterised with the
    origin decode_line uses (main.rs:133-147: minus the alphabet's raster-bounds origin), clipped to
    the (line_width x line_height) rectangle at (x_start, y_start + i*line_advance).
    Returns (gray u8, lines)."""
    from .raster import FOCR_DEFAULT_ALPHABET

    alphabet = alphabet or FOCR_DEFAULT_ALPHABET
    ink = np.zeros((height, width), np.uint8)
    n_lines = max(int(((height - y_start - line_height) // line_advance) * fill), 0)
    gids = {c: font.glyph_for_char(c) for c in alphabet}
    x0 = y0 = 0
    for gid in gids.values():
        a, b, _, _ = font.raster_bounds(gid, size, 0.0, 0.0)
        x0, y0 = min(x0, a), min(y0, b)
    ox, oy = f32(-x0), f32(-y0)
    upem = f32(font.units_per_em)
    body = [c for c in alphabet if c in "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/"]
    avg = float(np.mean([font.advance(gids[c])[0] for c in body])) / font.units_per_em * size
    per_line = max(int((line_width - 8) / avg) - 4, 1)
    rng = np.random.default_rng(seed)
    lines = []
    for li in range(n_lines):
        line = "> " + "".join(body[i] for i in rng.integers(0, len(body), per_line))
        canvas = np.zeros((line_height, line_width), np.uint8)
        pos = f32(0.0)
        for ch in line:
            g = np.zeros_like(canvas)
            font.rasterize_glyph(g, gids[ch], size, f32(ox + pos), oy)
            np.maximum(canvas, g, out=canvas)
            pos = f32(pos + f32(f32(font.advance(gids[ch])[0] / upem) * f32(size)))
        y = y_start + li * line_advance
        h, w = min(line_height, height - y), min(line_width, width - x_start)
        if h <= 0 or w <= 0:
            break
        np.maximum(ink[y:y + h, x_start:x_start + w], canvas[:h, :w], out=ink[y:y + h, x_start:x_start + w])
        lines.append(line)
    return (255 - ink).astype(np.uint8), lines


__all__ = ["base64_text", "make_ncc_page", "make_focr_page", "focr_render_line", "_blit"]
