"""Parity of the CUDA focr path (ABI section 3) with the oracle's restatement of main.rs:87-218.
Integer scores -> the chosen glyph sequence must be identical, character for character."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_sum_of_squares_exact(ctx, oracle):
    from font_ocr_b200 import focr

    rng = np.random.default_rng(1)
    xs = rng.integers(0, 256, (7, 608 * 12), dtype=np.uint8)
    ys = rng.integers(0, 256, (7, 608 * 12), dtype=np.uint8)
    ys[3] = xs[3]
    got = focr.sum_of_squares(ctx, xs, ys)
    exp = [oracle.sum_of_squares(a, b) for a, b in zip(xs, ys)]
    assert got.tolist() == exp and got[3] == 0


def test_decode_matches_oracle(ctx, oracle, font, pkg):
    """BASELINE config 4 geometry on a small page: -x 45 -y 39 -w 608 --line-height 12 --line-advance 15."""
    from font_ocr_b200 import focr

    alphabet = pkg.raster.FOCR_DEFAULT_ALPHABET
    page, lines = pkg.pages.make_focr_page(font, 13, 700, 39 + 15 * 6 + 20, seed=5, fill=1.0)
    page[39 + 15 * 2:39 + 15 * 2 + 12, :] = 255  # an all-white rectangle in the middle: must be skipped
    bank = focr.GlyphBank(ctx, font, 13, alphabet)
    got = focr.decode_images(ctx, bank, page, 45, 39, 608, 12, 15)[0]
    exp = oracle.decode_image(page, font, alphabet, 13, 45, 39, 608, 12, 15)
    assert [y for _, y in got] == [y for _, y in exp]
    assert [t for t, _ in got] == [t for t, _ in exp]
    assert len(got) >= 4 and (39 + 30) not in [y for _, y in got]
    assert all(t.startswith(">") for t, _ in got)  # every line was rendered with the "> " prefix
    bank.close()


def test_decode_batch_and_clamped_last_strip(ctx, oracle, font, pkg):
    """Two pages per call; the page height is chosen so that the last rectangle is clamped by
    crop_imm to fewer than line_height rows (main.rs:201-203)."""
    from font_ocr_b200 import focr

    alphabet = "> =ABCDEFGHabcdefgh0123+/"
    H = 39 + 15 * 3 + 7
    pages = np.stack([pkg.pages.make_focr_page(font, 13, 400, H, seed=s, line_width=300, fill=1.0)[0] for s in (1, 2)])
    pages[:, H - 6:H - 2, 60:200] = 0  # ink inside the clamped last strip
    bank = focr.GlyphBank(ctx, font, 13, alphabet)
    got = focr.decode_images(ctx, bank, pages, 45, 39, 300, 12, 15)
    for p in range(2):
        exp = oracle.decode_image(pages[p], font, alphabet, 13, 45, 39, 300, 12, 15)
        assert got[p] == exp
    bank.close()


def test_config4_full_page_and_multi_device(ctx, oracle, font, pkg):
    """BASELINE config 4 at its named shape: -x 45 -y 39 -w 608 --line-height 12 --line-advance 15 on full 2480x3508 pages
    with the cached glyph rasters, against the cached C restatement (pinned to the per-cell restatement in
    tests/test_focr_oracle.py); then the same batch sharded by page over a focr_multi (every visible GPU, or two contexts
    on device 0) must give the same lines."""
    import torch
    from font_ocr_b200 import focr, ncc

    alphabet = pkg.raster.FOCR_DEFAULT_ALPHABET
    pages = np.stack([pkg.pages.make_focr_page(font, 13, 2480, 3508, seed=7100 + i)[0] for i in range(3)])
    bank = focr.GlyphBank(ctx, font, 13, alphabet)
    got = focr.decode_images(ctx, bank, pages, 45, 39, 608, 12, 15)
    bank.close()
    cache = oracle.GlyphCache(font, alphabet, 13)
    exp0 = oracle.decode_image_cached(pages[0], cache, 45, 39, 608, 12, 15)
    assert got[0] == exp0 and len(exp0) > 200
    n_gpu = torch.cuda.device_count()
    mctx = ncc.MultiContext(devices=list(range(n_gpu)) if n_gpu > 1 else [0, 0])
    try:
        mbank = focr.GlyphBank(mctx, font, 13, alphabet)
        got_m = focr.decode_images(mctx, mbank, pages, 45, 39, 608, 12, 15)
        mbank.close()
    finally:
        mctx.close()
    assert got_m == got


def test_full_size_batch_invariances(ctx, font, pkg, monkeypatch):
    """focr_decode_pages over 40 full-size pages (three chunks: both slots are reused) from pageable and from pinned host
    memory, against the same pages decoded one per call; the row-task kernel (FOCR_DECODE_LEGACY) and the band layout
    (FOCR_DECODE_BAND) must decode the same text as the tile kernel on the compact layout."""
    import torch
    from font_ocr_b200 import focr

    alphabet = pkg.raster.FOCR_DEFAULT_ALPHABET
    distinct = [pkg.pages.make_focr_page(font, 13, 2480, 3508, seed=7300 + i)[0] for i in range(4)]
    order = [(7 * i + i // 4) % 4 for i in range(40)]
    pages = np.stack([distinct[k] for k in order])
    bank = focr.GlyphBank(ctx, font, 13, alphabet)
    try:
        single = [focr.decode_images(ctx, bank, d[None], 45, 39, 608, 12, 15)[0] for d in distinct]
        assert all(len(s) > 200 for s in single)
        got = focr.decode_images(ctx, bank, pages, 45, 39, 608, 12, 15)                       # pageable: staged gather
        assert [got[i] == single[k] for i, k in enumerate(order)] == [True] * 40
        pinned = torch.from_numpy(pages).pin_memory()
        got_p = focr.decode_images(ctx, bank, pinned.numpy(), 45, 39, 608, 12, 15)             # pinned: strided DMA
        assert got_p == got
        monkeypatch.setenv("FOCR_DECODE_LEGACY", "1")
        assert focr.decode_images(ctx, bank, pages[:5], 45, 39, 608, 12, 15) == got[:5]
        monkeypatch.delenv("FOCR_DECODE_LEGACY")
        monkeypatch.setenv("FOCR_DECODE_BAND", "1")
        assert focr.decode_images(ctx, bank, pinned.numpy()[:5], 45, 39, 608, 12, 15) == got[:5]
    finally:
        bank.close()

