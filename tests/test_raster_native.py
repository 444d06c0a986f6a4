"""The C++ FreeType driver (font-ocr_b200/host/focr_raster.cpp, ABI section 5) against the Python producer
(font-ocr_b200/raster.py) that every other test and bench.py feed to the oracle and the GPU: byte-identical template
banks (all box modes, subpixel offsets, padding) and glyph-raster banks.  CPU only: FreeType runs on the host."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def nfont(built_lib, pkg, font):
    f = pkg.raster.NativeFont(font.path)
    yield f
    f.close()


@pytest.mark.parametrize("size,x_bits,y_bits,box,pad", [(13, 0, 0, "alphabet", (0, 0)), (13, 2, 0, "alphabet", (0, 0)),
                                                        (7, 2, 2, "alphabet", (0, 0)), (24, 1, 1, "alphabet", (1, 2)),
                                                        (13, 1, 0, "font", (0, 0)), (13, 0, 1, "char", (2, 1))])
def test_template_bank_identical_to_python(nfont, pkg, font, size, x_bits, y_bits, box, pad):
    ref = pkg.raster.TemplateBank(font, size, x_bits=x_bits, y_bits=y_bits, box_size=box, padding=pad)
    tpls, letters, cy = nfont.template_bank(size, pkg.raster.NCC_DEFAULT_ALPHABET, x_bits, y_bits, box, pad)
    assert len(tpls) == len(ref) and letters == ref.letters()
    for i, (a, t) in enumerate(zip(tpls, ref.templates)):
        assert a.shape == t.pixels.shape and np.array_equal(a, t.pixels), (i, t.letter, t.offset)
        assert np.float32(cy[i]) == np.float32(t.corrected_y)
    assert any(t.any() for t in tpls)


def test_config5_bank_and_missing_glyph(nfont, pkg, font):
    alphabet = "".join(chr(c) for c in range(32, 127))
    ref = pkg.raster.TemplateBank(font, 24, x_bits=1, y_bits=0, alphabet=alphabet)
    tpls, letters, _ = nfont.template_bank(24, alphabet, 1, 0)
    assert all(np.array_equal(a, t.pixels) for a, t in zip(tpls, ref.templates)) and len(tpls) == 2 * 95
    from font_ocr_b200 import native

    with pytest.raises(native.FocrError) as e:   # the reference .unwrap()s a missing glyph
        nfont.template_bank(13, "A\U0010ffff")
    assert "panic" in str(e.value)


def test_glyph_bank_identical_to_python(nfont, pkg, font, oracle):
    alphabet = pkg.raster.FOCR_DEFAULT_ALPHABET
    px, ras, adv, origin = nfont.glyph_bank(13, alphabet)
    cache = oracle.GlyphCache(font, alphabet, 13)   # built with raster.Font's calls, the layout focr_glyph_bank_create takes
    assert origin[0] == cache.origin_x
    assert np.array_equal(adv, cache.advance_px)
    assert ras.shape == cache.rasters.shape
    for g in range(len(alphabet)):
        for ph in range(64):
            a, b = ras[g, ph], cache.rasters[g, ph]
            assert (a["left"], a["top"], a["w"], a["h"]) == (b["left"], b["top"], b["w"], b["h"]), (g, ph)
            n = int(a["w"]) * int(a["h"])
            assert np.array_equal(px[int(a["offset"]):int(a["offset"]) + n], cache.pixels[int(b["offset"]):int(b["offset"]) + n])


def test_hinted_banks_identical_and_phase_property(built_lib, pkg, font):
    """--hinting (HintingOptions::Full, ncc.rs:547-551): both producers rasterise with FT_LOAD_TARGET_NORMAL; the banks stay
    byte-identical to each other, and a hinted bitmap still only depends on the 26.6 phase of
    the pen delta (FreeType applies the transform after hinting) -- what the (glyph, phase) cache relies on."""
    hfont = pkg.raster.Font(font.path, hinting=True)
    nfont = pkg.raster.NativeFont(font.path, hinting=True)
    ref = pkg.raster.TemplateBank(hfont, 13, x_bits=1)
    tpls, letters, _ = nfont.template_bank(13, pkg.raster.NCC_DEFAULT_ALPHABET, 1, 0)
    assert all(np.array_equal(a, t.pixels) for a, t in zip(tpls, ref.templates)) and len(tpls) == len(ref)
    # (whether hinted rasters differ from unhinted ones depends on the font: the Lato web font of this image carries no
    # instructions and its rasters come out the same)
    for ch in "AgW/":
        gid = hfont.glyph_for_char(ch)
        for frac in (0, 17, 63):
            b0, l0, t0 = hfont.glyph_bitmap(gid, 13, frac, -10 * 64)
            b1, l1, t1 = hfont.glyph_bitmap(gid, 13, frac + 64 * 37, -10 * 64)
            assert np.array_equal(b0, b1) and l1 == l0 + 37 and t1 == t0
    nfont.close()
