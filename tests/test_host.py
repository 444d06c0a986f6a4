"""Host-side logic on CPU: the template producer's geometry, the page generator, and the two
mirrors of process_hits (Python in ncc.py, C++ in host/focr_host.cpp) against the oracle."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st


def test_offset_grid_is_x_major(pkg):
    g = pkg.raster.offset_grid(1, 2)  # ncc.rs:563-573: for x { for y }
    assert [(float(a), float(b)) for a, b in g] == [(0, 0), (0, .25), (0, .5), (0, .75), (.5, 0), (.5, .25), (.5, .5), (.5, .75)]


def test_bank_order_and_box_sizes(font, pkg):
    bank = pkg.raster.TemplateBank(font, 13, x_bits=1)
    assert len(bank) == 2 * 74 and bank.letters()[:3] == ["A", "B", "C"] and bank.letters()[74] == "A"
    assert bank.templates[0].offset_index == 0 and bank.templates[74].offset_index == 1
    # BoxSize::Alphabet: all letters of one offset share one canvas size (ncc.rs:600-626)
    for oi in (0, 1):
        assert len({t.pixels.shape for t in bank.templates if t.offset_index == oi}) == 1
    char_bank = pkg.raster.TemplateBank(font, 13, alphabet="Ail", box_size="char")
    assert len({t.pixels.shape for t in char_bank.templates}) > 1  # tight per-glyph boxes (ncc.rs:627)
    with pytest.raises(ValueError):
        pkg.raster.TemplateBank(font, 13, box_size="bogus")  # ncc.rs:559 unwraps the TryFrom error


def test_page_generator_is_deterministic(font, pkg):
    bank = pkg.raster.TemplateBank(font, 13)
    a = pkg.pages.make_ncc_page(bank, 608, 300, seed=4, shifts="bank")
    b = pkg.pages.make_ncc_page(bank, 608, 300, seed=4, shifts="bank")
    c = pkg.pages.make_ncc_page(bank, 608, 300, seed=5, shifts="bank")
    assert np.array_equal(a[0], b[0]) and a[1] == b[1] and not np.array_equal(a[0], c[0])
    assert a[0].dtype == np.uint8 and a[0].max() == 255 and a[0].min() < 128


hit = st.tuples(st.sampled_from("ABCabc+/"), st.integers(0, 60), st.integers(0, 6),
                st.sampled_from([0.5, 0.81, 0.9, 0.95, 0.96, 0.99]).map(np.float32))


@settings(max_examples=150, deadline=None)
@given(st.lists(hit, min_size=0, max_size=40), st.integers(0, 8))
def test_process_hits_mirrors_agree_with_oracle(built_lib, oracle, hits, overlap):
    from font_ocr_b200 import ncc

    try:
        exp = oracle.process_hits(hits, 0.95, overlap)
    except IndexError:
        exp = None  # the reference panics: no hit reaches the anchor threshold (ncc.rs:1040)
    for fn in (ncc.process_hits, ncc.host_process_hits):
        if exp is None:
            with pytest.raises(IndexError):
                fn(hits, 0.95, overlap)
        else:
            got = fn(hits, 0.95, overlap)
            assert [[(h[0], h[1], h[2], float(h[3])) for h in line] for line in got] == \
                   [[(h[0], h[1], h[2], float(h[3])) for h in line] for line in exp]


def test_cpp_partition_by_anchoring(built_lib):
    """Groups are anchored to their first element: x = 0, 4, 8 with overlap 5 -> {0, 4} {8}; ties on the
    similarity go to the LAST element (Iterator::max_by)."""
    from font_ocr_b200 import ncc

    f = np.float32
    hits = [("a", 0, 3, f(0.96)), ("b", 4, 3, f(0.96)), ("c", 8, 3, f(0.5))]
    assert ncc.lines_to_text(ncc.host_process_hits(hits, 0.95, 5)) == ["bc"]


def test_space_detection_extension_matches_restatement(built_lib, oracle, font):
    """Opt-in extension (README.md:46: the reference does not detect spaces): the C++ host mirror's
    line_text_with_spaces against the oracle's restatement, on lines laid out with the reference's own pen arithmetic
    (f32 advances, main.rs:176-178, hit x = floor of the pen position like a grid-snapping renderer), and the text with
    its spaces must come back; without the flag's space advance nothing is inserted."""
    from font_ocr_b200 import ncc

    f32 = np.float32
    upem = f32(font.units_per_em)
    px = lambda ch: float(f32(f32(font.advance(font.glyph_for_char(ch))[0] / upem) * f32(13)))
    alphabet = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/="
    adv = {ch: px(ch) for ch in alphabet}
    space = px(" ")
    rng = np.random.default_rng(12)
    lines, texts = [], []
    for _ in range(20):
        words = ["".join(rng.choice(list(alphabet), int(rng.integers(1, 9)))) for _ in range(int(rng.integers(1, 7)))]
        text = (" " * int(rng.integers(1, 4))).join(words) if rng.random() < 0.5 else " ".join(words)
        pen, line = f32(45.0), []
        for ch in text:
            if ch != " ":
                line.append((ch, int(np.floor(pen)), 39, f32(0.97)))
            pen = f32(pen + f32(adv.get(ch, space)))
        lines.append(line)
        texts.append(text)
    got = ncc.lines_to_text_with_spaces(lines, adv, space)
    assert got == oracle.lines_to_text_with_spaces(lines, adv, space)
    assert got == texts
    assert ncc.lines_to_text_with_spaces(lines, adv, 0.0) == [t.replace(" ", "") for t in texts] == ncc.lines_to_text(lines)
