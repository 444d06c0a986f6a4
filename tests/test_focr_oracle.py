"""CPU checks of the focr oracle restatement and of the (glyph, shift) bank assumption."""
import numpy as np


def test_sum_of_squares_definition(oracle):
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, 5000, dtype=np.uint8)
    b = rng.integers(0, 256, 5000, dtype=np.uint8)
    assert oracle.sum_of_squares(a, b) == int(((a.astype(np.int64) - b.astype(np.int64)) ** 2).sum())


def test_bitmap_depends_only_on_the_26_6_phase(font):
    """The raster cache is keyed by (glyph, d & 63): an integer-pixel change of the FreeType delta must
    move the bitmap by whole pixels and leave its bytes unchanged."""
    for ch in "AgW/":
        gid = font.glyph_for_char(ch)
        for frac in (0, 17, 63):
            b0, l0, t0 = font.glyph_bitmap(gid, 13, frac, -10 * 64)
            b1, l1, t1 = font.glyph_bitmap(gid, 13, frac + 64 * 37, -10 * 64)
            assert np.array_equal(b0, b1) and l1 == l0 + 37 and t1 == t0


def test_oracle_decodes_rendered_line(oracle, font, pkg):
    page, lines = pkg.pages.make_focr_page(font, 13, 700, 39 + 15 * 2 + 14, seed=3, fill=1.0)
    out = oracle.decode_image(page, font, pkg.raster.FOCR_DEFAULT_ALPHABET, 13, 45, 39, 608, 12, 15, max_lines=1)
    assert out and out[0][1] == 39
    # the greedy walk derails on a proportional face (README.md: only tested with a monospace font), but
    # it must at least lock on to the start of the line it was rendered from
    assert out[0][0][:4] == lines[0][:4]


def test_cached_decode_equals_per_cell_rasterisation(oracle, font, pkg):
    """BASELINE config 4 runs focr "with cached glyph rasters".  The cached C restatement (oracle.decode_image_cached:
    the reference's whole-canvas sum_of_squares, rasters from a (glyph, 26.6 phase) cache) must decode exactly what the
    per-cell restatement (oracle.decode_image: one rasterisation per glyph per cell like main.rs:98-106) decodes --
    including a skipped all-white rectangle and a last rectangle clamped by crop_imm."""
    alphabet = "> =ABCDEFGHabcdefgh0123+/"
    H = 39 + 15 * 3 + 7
    page = pkg.pages.make_focr_page(font, 13, 400, H, seed=4, line_width=300, fill=1.0)[0]
    page[39 + 15:39 + 15 + 12, :] = 255     # an all-white rectangle: skipped (main.rs:208-211)
    page[H - 6:H - 2, 60:200] = 0           # ink inside the clamped last strip
    cache = oracle.GlyphCache(font, alphabet, 13)
    got = oracle.decode_image_cached(page, cache, 45, 39, 300, 12, 15)
    exp = oracle.decode_image(page, font, alphabet, 13, 45, 39, 300, 12, 15)
    assert got == exp and len(got) >= 2 and got[-1][1] == 39 + 15 * 3
