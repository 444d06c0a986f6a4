"""Parity of the CUDA NCC path with the oracle, through the C ABI (run on the B200: -m gpu).

Bars (north_star): window sums and raw numerators bit-exact; match lists identical including order,
the 1024 truncation and the f32 score bits (the 1e-5 tolerance north_star allows is not needed);
post-processed text identical."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["simt", "tcgen05"])
def kctx(request, ctx):
    from font_ocr_b200 import native

    ctx.set_kernel(native.KERNEL_SIMT if request.param == "simt" else native.KERNEL_TCGEN05)
    yield ctx
    ctx.set_kernel(native.KERNEL_AUTO)


class Unsupported(Exception):
    """The forced kernel declines the shape (FOCR_ERR_UNSUPPORTED).  Never a silent skip: tests that loop over trials
    catch it per trial and assert how many trials ran."""


def _scan(kctx, templates, pages, thr, n_out=1024):
    from font_ocr_b200 import native, ncc

    bank = ncc.Bank(kctx, templates)
    try:
        return ncc.scan_pages(kctx, bank, pages, thr, n_out)
    except native.FocrError as e:
        if e.code == native.FOCR_ERR_UNSUPPORTED:
            raise Unsupported(str(e)) from e
        raise
    finally:
        bank.close()


def _assert_same(matches, counts, expected_lists, tag=""):
    for t, exp in enumerate(expected_lists):
        c = int(counts[t])
        assert c == len(exp), f"{tag} template {t}: count {c} != {len(exp)}"
        got = matches[t, :c]
        if got.tobytes() != exp.tobytes():
            bad = [i for i in range(c) if got[i].tobytes() != exp[i].tobytes()]
            raise AssertionError(f"{tag} template {t}: {len(bad)} of {c} differ, first {got[bad[0]]} vs {exp[bad[0]]}")


def test_window_stats_bit_exact(ctx, oracle, golden):
    from font_ocr_b200 import ncc

    rng = np.random.default_rng(11)
    cases = [(rng.integers(0, 256, (61, 97), dtype=np.uint8), [(5, 4), (16, 16), (1, 1), (32, 7)]),
             (golden["text16_page"], [(15, 14), (8, 7), (27, 26)]),
             (rng.integers(0, 256, (300, 700), dtype=np.uint8), [(32, 64), (13, 33)])]
    for page, sizes in cases:
        inv = (255 - page).astype(np.uint8)
        for n_w, n_h in sizes:
            sp, s2, rn = ncc.window_stats(ctx, page, n_w, n_h)
            esp, es2 = oracle.window_sums(inv, n_w, n_h)
            ys, xs = esp.shape
            assert np.array_equal(sp[:ys, :xs], esp), (n_w, n_h)
            assert np.array_equal(s2[:ys, :xs], es2), (n_w, n_h)
            with np.errstate(all="ignore"):  # ncc.rs:309-311 in IEEE f64
                ern = 1.0 / np.sqrt(es2.astype(np.float64) - (esp * esp).astype(np.float64) / float(n_w * n_h))
            assert np.array_equal(rn[:ys, :xs], ern), (n_w, n_h)


def test_numerators_bit_exact(ctx, oracle, golden):
    from font_ocr_b200 import ncc

    page, tpl = golden["text16_page"], golden["text16_tpl"]
    rng = np.random.default_rng(5)
    noise = rng.integers(0, 256, (90, 200), dtype=np.uint8)
    wide = rng.integers(0, 256, (26, 27), dtype=np.uint8)
    for pg, templates in ((page, [tpl[0], tpl[40]]), (noise, [tpl[3], wide, wide[:3, :20]])):
        bank = ncc.Bank(ctx, templates)
        inv = (255 - pg).astype(np.uint8)
        for t, tp in enumerate(templates):
            acc = ncc.numerators(ctx, bank, t, pg)
            win = np.lib.stride_tricks.sliding_window_view(inv, tp.shape)
            exp = np.einsum("yxij,ij->yx", win.astype(np.int64), tp.astype(np.int64))
            ys, xs = exp.shape
            assert np.array_equal(acc[1:ys, 1:xs], exp[1:, 1:]), t  # row/col 0 are never searched
        bank.close()


def test_scan_matches_golden_text(kctx, golden):
    for name in ("text16", "text8"):
        m, c = _scan(kctx, list(golden[f"{name}_tpl"]), golden[f"{name}_page"], float(golden[f"{name}_thr"]))
        exp, off = [], 0
        for n in golden[f"{name}_counts"]:
            exp.append(golden[f"{name}_hits"][off:off + n])
            off += n
        _assert_same(m[0], c[0], exp, name)


def test_scan_matches_golden_truncation(kctx, golden):
    n_out = int(golden["noise_n_out"])
    tpls = [golden[f"noise_tpl{i}"] for i in range(int(golden["noise_n"]))]
    m, c = _scan(kctx, tpls, golden["noise_page"], float(golden["noise_thr"]), n_out)
    _assert_same(m[0], c[0], [golden[f"noise_hits{i}"] for i in range(len(tpls))], "noise")
    m, c = _scan(kctx, [golden["dense_tpl"]], golden["dense_page"], float(golden["dense_thr"]))
    assert c[0, 0] == 1024  # ncc.cpp:225-227: full -> returns n_out
    _assert_same(m[0], c[0], [golden["dense_hits"]], "dense")


def test_scan_random_vs_oracle(kctx, oracle):
    rng = np.random.default_rng(2024)
    ran = 0
    for trial in range(10):
        h, w = int(rng.integers(40, 200)), int(rng.integers(40, 300))
        page = rng.integers(0, 256, (h, w), dtype=np.uint8)
        if trial % 2 == 0:
            page[rng.random((h, w)) < 0.8] = 255
        if trial % 3 == 0:
            page[10:30, 5:35] = 0  # a saturated block: constant non-zero windows
        n_w, n_h = int(rng.integers(1, 33)), int(rng.integers(1, 33))
        tpls = [rng.integers(0, 256, (n_h, n_w), dtype=np.uint8) for _ in range(int(rng.integers(1, 40)))]
        tpls[0][:] = 0  # an all-zero template (e.g. space): NaN similarity everywhere -> no hits
        if len(tpls) > 2:
            tpls[1][:] = 77  # a constant template: rnorm_n = inf
        thr = float(rng.choice([0.1, 0.2, 0.5]))
        n_out = int(rng.choice([16, 1024]))
        try:
            m, c = _scan(kctx, tpls, page, thr, n_out)
        except Unsupported:
            continue   # a shape the forced tcgen05 kernel declines (AUTO routes it to the SIMT kernel)
        ran += 1
        s = oracle.Searcher(page, "port")
        exp = [s.search_c_u8(t, thr, n_out=n_out, allow_wide=True) for t in tpls]
        _assert_same(m[0], c[0], exp, f"trial {trial} box {n_w}x{n_h}")
    assert ran >= 9, f"only {ran} of 10 random shapes ran"


def test_mixed_size_bank_and_batch(kctx, oracle, font, pkg):
    """--x-bits 2 --y-bits 2 at -t 7: four box sizes in one bank (ncc.rs:600-626); three pages per call."""
    bank = pkg.raster.TemplateBank(font, 7, x_bits=2, y_bits=2)
    assert len(bank.sizes()) > 1
    tpls = [t.pixels for t in bank.templates]
    pages = np.stack([pkg.pages.make_ncc_page(bank, 304, 200, seed=s, margin_x=11, margin_y=9, shifts="bank")[0]
                      for s in range(3)])
    m, c = _scan(kctx, tpls, pages, 0.8)
    for p in range(3):
        exp = oracle.get_hits(pages[p], tpls, 0.8, "port")
        _assert_same(m[p], c[p], exp, f"page {p}")


def test_shim_is_the_reference_ffi(ctx, golden):
    """Searcher.search_c_u8 marshals like ncc.rs:332-404 and calls the exported ncc_8_u8/ncc_16_u8."""
    from font_ocr_b200 import ncc

    s = ncc.Searcher(golden["text16_page"])
    off = 0
    for t in range(0, 74, 9):
        n = int(golden["text16_counts"][t])
        off = int(golden["text16_counts"][:t].sum())
        got = s.search_c_u8(golden["text16_tpl"][t], 0.8)
        assert got.tobytes() == golden["text16_hits"][off:off + n].tobytes(), t
    s8 = ncc.Searcher(golden["text8_page"])
    got = s8.search_c_u8(golden["text8_tpl"][5], 0.8)
    off = int(golden["text8_counts"][:5].sum())
    assert got.tobytes() == golden["text8_hits"][off:off + int(golden["text8_counts"][5])].tobytes()


def test_decoded_text_identical(kctx, oracle, golden):
    from font_ocr_b200 import ncc

    letters = list(golden["text16_letters"])
    m, c = _scan(kctx, list(golden["text16_tpl"]), golden["text16_page"], 0.8)
    # the C++ host mirror (independent of the oracle's Python restatement; ncc.process_hits is its character-level twin)
    ours = ncc.lines_to_text(ncc.host_process_hits(ncc.get_hits(m[0], c[0], letters)))
    assert ours == ncc.lines_to_text(ncc.process_hits(ncc.get_hits(m[0], c[0], letters)))
    ref_hits = oracle.get_hits(golden["text16_page"], list(golden["text16_tpl"]), 0.8, "port")
    theirs = oracle.lines_to_text(oracle.process_hits(oracle.hits_with_letters(ref_hits, letters)))
    assert ours == theirs
    truth = list(golden["text16_lines"])
    assert sum(a == b for a, b in zip(ours, truth)) >= len(truth) - 3


def test_full_size_page_properties_and_sample(kctx, oracle, font, pkg):
    """BASELINE config 3 shape: 2480x3508, --x-bits 2 (296 templates).  The oracle checks a sample of
    templates exactly (the compiled reference when present); size-independent properties cover the rest:
    raster order, the 1024 cap, every score above the threshold, idempotence."""
    bank = pkg.raster.TemplateBank(font, 13, x_bits=2)
    tpls = [t.pixels for t in bank.templates]
    page = pkg.pages.make_ncc_page(bank, 2480, 3508, seed=3, shifts="bank")[0]
    m, c = _scan(kctx, tpls, page, 0.8)
    m2, c2 = _scan(kctx, tpls, page, 0.8)
    assert np.array_equal(c, c2) and m.tobytes() == m2.tobytes()  # deterministic despite atomics
    assert c.max() == 1024 and (c <= 1024).all()
    for t in range(len(tpls)):
        g = m[0, t, :c[0, t]]
        key = g["y"].astype(np.int64) * 65536 + g["x"]
        assert (np.diff(key) > 0).all(), t
        assert (g["similarity"] > np.float32(0.8) - 1e-6).all()
        n_h, n_w = tpls[t].shape
        assert (g["x"] >= 1).all() and (g["y"] >= 1).all()
        assert (g["x"] <= 2480 - n_w).all() and (g["y"] <= 3508 - n_h).all()
    impl = "reference" if oracle.ref_lib() is not None else "port"
    s = oracle.Searcher(page, impl)
    sample = list(range(0, len(tpls), 37)) if impl == "reference" else [0, 150]
    for t in sample:
        exp = s.search_c_u8(tpls[t], 0.8)
        _assert_same(m[0, t:t + 1], c[0, t:t + 1], [exp], f"template {t}")


def test_full_size_invariances(ctx, font, pkg):
    """Properties at BASELINE config 3's full size that need no oracle: a page's match lists do not depend on the batch it
    is scanned in (alone, in a batch through the chunked host pipeline, in another position through the device-resident entry),
    a permuted bank permutes the lists (the tcgen05 path orders its columns by template similarity -- the bank order must not
    leak), and the hits at a higher threshold are exactly the higher-scoring hits of the lower one (a prefix of them where the lower list was cut at n_out)."""
    import torch
    from font_ocr_b200 import native, ncc

    tb = pkg.raster.TemplateBank(font, 13, x_bits=2)
    tpls = [t.pixels for t in tb.templates]
    T, n_out = len(tpls), 4096
    pages = np.stack([pkg.pages.make_ncc_page(tb, 2480, 3508, seed=40 + i, shifts="bank")[0] for i in range(5)])
    bank = ncc.Bank(ctx, tpls)
    m_all, c_all = ncc.scan_pages(ctx, bank, pages, 0.8, n_out)                 # chunks of 2 + 3 pages
    for p in (0, 3, 4):
        m1, c1 = ncc.scan_pages(ctx, bank, pages[p], 0.8, n_out)
        assert np.array_equal(c1[0], c_all[p]) and all(m1[0, t, :c1[0, t]].tobytes() == m_all[p, t, :c1[0, t]].tobytes() for t in range(T)), p
    rev = torch.from_numpy(pages[::-1].copy()).cuda()
    out_dev = torch.zeros(5 * T * n_out * 8, dtype=torch.uint8, device="cuda")
    cnt_dev = torch.zeros(5 * T, dtype=torch.int32, device="cuda")
    ncc.scan_pages_device(ctx, bank, rev.data_ptr(), 2480 * 3508, 2480, 2480, 3508, 5, 0.8, n_out, out_dev.data_ptr(), cnt_dev.data_ptr())
    m_rev = out_dev.cpu().numpy().view(native.MATCH_DTYPE).reshape(5, T, n_out)
    c_rev = cnt_dev.cpu().numpy().view(np.uint32).reshape(5, T)
    for p in range(5):
        assert np.array_equal(c_rev[4 - p], c_all[p])
        assert all(m_rev[4 - p, t, :c_all[p, t]].tobytes() == m_all[p, t, :c_all[p, t]].tobytes() for t in range(T)), p
    bank.close()
    # a permuted bank
    perm = np.random.default_rng(4).permutation(T)
    bank_p = ncc.Bank(ctx, [tpls[i] for i in perm])
    m_p, c_p = ncc.scan_pages(ctx, bank_p, pages[1], 0.8, n_out)
    for k, t in enumerate(perm):
        assert c_p[0, k] == c_all[1, t] and m_p[0, k, :c_p[0, k]].tobytes() == m_all[1, t, :c_all[1, t]].tobytes(), (k, t)
    # threshold monotonicity
    m_hi, c_hi = ncc.scan_pages(ctx, bank_p, pages[1], 0.9, n_out)
    bank_p.close()
    for k in range(T):
        lo = m_p[0, k, :c_p[0, k]]
        keep = lo[lo["similarity"] > np.float32(0.9)]
        if c_p[0, k] < n_out:   # not truncated: exactly the higher-scoring hits
            assert keep.tobytes() == m_hi[0, k, :c_hi[0, k]].tobytes(), k
        else:                   # truncated in raster order: they are the first of them
            assert c_hi[0, k] >= len(keep) and keep.tobytes() == m_hi[0, k, :len(keep)].tobytes(), k
    assert 0 < c_hi.sum() < c_p.sum()


def test_cpp_host_searcher_matches_golden(ctx, golden):
    """The C++ Searcher mirror (host/focr_host.cpp) marshals like ncc.rs:332-404 and reproduces the
    reference's match lists; widths above 16 'panic' like ncc.rs:392."""
    from font_ocr_b200 import ncc

    for t in (0, 33, 70):
        n = int(golden["text16_counts"][t])
        off = int(golden["text16_counts"][:t].sum())
        got = ncc.host_search_c_u8(golden["text16_page"], golden["text16_tpl"][t], 0.8)
        assert got.tobytes() == golden["text16_hits"][off:off + n].tobytes(), t
    with pytest.raises(NotImplementedError):
        ncc.host_search_c_u8(golden["text16_page"], np.zeros((5, 17), np.uint8), 0.8)


def test_process_hits_on_device(ctx, oracle, font, pkg):
    """process_hits (ncc.rs:723-786) on the GPU, on match lists that never leave HBM, against
    the oracle's restatement and the host mirrors: identical lines (letters, coordinates, f32 scores, order),
    including duplicate lines at y+-1, ties, and a page without any anchor line."""
    import torch
    from font_ocr_b200 import native, ncc

    bank_h = pkg.raster.TemplateBank(font, 13, x_bits=1)
    tpls = [t.pixels for t in bank_h.templates]
    letters = bank_h.letters()
    T, n_out = len(tpls), 1024
    pages = np.stack([pkg.pages.make_ncc_page(bank_h, 608, 400, seed=s, shifts="bank")[0] for s in range(3)] +
                     [np.full((400, 608), 255, np.uint8)])  # the last page is blank: no hits, no lines
    P = len(pages)
    bank = ncc.Bank(ctx, tpls)
    dev = torch.from_numpy(pages).cuda()
    out_dev = torch.zeros(P * T * n_out * 8, dtype=torch.uint8, device="cuda")
    cnt_dev = torch.zeros(P * T, dtype=torch.int32, device="cuda")
    ncc.scan_pages_device(ctx, bank, dev.data_ptr(), 608 * 400, 608, 608, 400, P, 0.8, n_out, out_dev.data_ptr(),
                          cnt_dev.data_ptr())
    ctx.sync()
    m = out_dev.cpu().numpy().view(native.MATCH_DTYPE).reshape(P, T, n_out)
    c = cnt_dev.cpu().numpy().view(np.uint32).reshape(P, T)
    for anchor, overlap in ((0.95, 5), (0.9, 0), (0.99, 40)):
        got = ncc.process_hits_device(ctx, out_dev.data_ptr(), cnt_dev.data_ptr(), T, n_out, P, letters, anchor, overlap)
        assert len(got) == P
        for p in range(P):
            hits = ncc.get_hits(m[p], c[p], letters)
            try:
                exp = oracle.process_hits(hits, anchor, overlap)
            except IndexError:  # the reference panics when no line has an anchor (ncc.rs:1040)
                exp = []
            assert len(got[p]) == len(exp), (anchor, overlap, p, len(got[p]), len(exp))
            for gl, el in zip(got[p], exp):
                assert [(h[0], h[1], h[2], np.float32(h[3]).tobytes()) for h in gl] == \
                       [(h[0], h[1], h[2], np.float32(h[3]).tobytes()) for h in el], (anchor, overlap, p)
        assert got[P - 1] == []
    assert any(len(lines) > 0 for lines in got)
    bank.close()


def test_scan_edge_cases_vs_oracle(kctx, oracle):
    """Edge cases of the screen + exact pass design, each against the oracle: more templates of one size than one
    launch holds (several launches, column remapping), thresholds that make every window a hit (candidate-list
    overflow and retry, the 1024/n_out cut), saturated pages where b*S + a*P leaves the linear range of the fp32
    trick ("always a candidate" windows), constant pages, n_out = 1, a page smaller than the template."""
    rng = np.random.default_rng(77)

    def check(page, tpls, thr, n_out, tag):
        m, c = _scan(kctx, tpls, page, thr, n_out)
        s = oracle.Searcher(page, "port")
        exp = [s.search_c_u8(t, thr, n_out=n_out, allow_wide=True) for t in tpls]
        _assert_same(m[0], c[0], exp, tag)
        return c

    # 600 templates of one box size: three launches of the tcgen05 kernel
    page = rng.integers(0, 256, (70, 180), dtype=np.uint8)
    page[rng.random(page.shape) < 0.7] = 255
    tpls = [rng.integers(0, 256, (9, 11), dtype=np.uint8) for _ in range(600)]
    check(page, tpls, 0.25, 1024, "600 templates")

    # every window is a hit: thr < -1
    page = rng.integers(0, 256, (120, 260), dtype=np.uint8)
    tpls = [rng.integers(0, 256, (9, 9), dtype=np.uint8) for _ in range(40)]
    c = check(page, tpls, -1.5, 64, "thr -1.5")
    assert (c == 64).all()
    c = check(page, tpls, -0.2, 1024, "thr -0.2")
    assert c.max() == 1024

    # saturated (nearly black) page and bright 16x16 templates: s_p and b = s_n/n are both near their maxima
    page = rng.integers(0, 12, (90, 200), dtype=np.uint8)
    page[30:50, 40:90] = rng.integers(0, 256, (20, 50), dtype=np.uint8)
    tpls = [np.clip(rng.integers(200, 256, (16, 16)) - (rng.random((16, 16)) < 0.1) * 150, 0, 255).astype(np.uint8) for _ in range(12)]
    check(page, tpls, 0.05, 1024, "saturated")

    # constant pages: every window is constant -> rnorm_p = inf -> no hit
    for v in (0, 255, 131):
        c = check(np.full((64, 100), v, np.uint8), tpls[:3], 0.1, 1024, f"constant {v}")
        assert c.sum() == 0

    # n_out = 1 and a page that is smaller than the template in one direction (no window at all)
    page = rng.integers(0, 256, (40, 60), dtype=np.uint8)
    tpls = [rng.integers(0, 256, (7, 8), dtype=np.uint8) for _ in range(5)]
    check(page, tpls, 0.0, 1, "n_out 1")
    m, c = _scan(kctx, [rng.integers(0, 256, (12, 20), dtype=np.uint8)], rng.integers(0, 256, (10, 64), dtype=np.uint8), 0.1)
    assert c.sum() == 0


def test_maximum_page_sizes(kctx, oracle):
    """The largest pages focr_get_limits advertises, against the oracle: the tallest page (u16 y coordinates, the grid.y limit of
    the staging kernel) and the widest one (finalize's shared-memory sort holds a whole row of keys); one pixel more is
    refused with FOCR_ERR_UNSUPPORTED instead of a CUDA error."""
    import ctypes as C
    from font_ocr_b200 import native, ncc

    lim = (C.c_uint32 * 5)()
    native.lib().focr_get_limits(lim)
    max_w, max_h = int(lim[2]), int(lim[3])
    assert max_w >= 4096 and max_h >= 32768
    rng = np.random.default_rng(31)
    tpls = [rng.integers(0, 256, (9, 9), dtype=np.uint8) for _ in range(6)] + [rng.integers(0, 256, (14, 15), dtype=np.uint8) for _ in range(3)]
    for shape in ((max_h, 48), (40, max_w)):
        page = rng.integers(0, 256, shape, dtype=np.uint8)
        page[rng.random(shape) < 0.6] = 255
        for t in tpls[:3]:   # plant each template near the far corner so that the last rows / columns hold hits
            h, w = t.shape
            page[shape[0] - h - 1:shape[0] - 1, shape[1] - w - 2:shape[1] - 2] = 255 - t   # (the library scans 255 - p)
        m, c = _scan(kctx, tpls, page, 0.3, 64)
        s = oracle.Searcher(page, "port")
        _assert_same(m[0], c[0], [s.search_c_u8(t, 0.3, n_out=64) for t in tpls], f"page {shape}")
        assert c[0].sum() > 0
    m, c = _scan(kctx, tpls[:3], rng.integers(0, 256, (max_h, 48), dtype=np.uint8), 0.999, 4)   # far-corner hit only
    for shape in ((max_h + 1, 48), (40, max_w + 1)):
        with pytest.raises(Unsupported):
            _scan(kctx, tpls[:2], np.zeros(shape, np.uint8), 0.5, 1024)   # (the width limit is quoted for the reference's n_out)


def test_sub_block_launch_matches_oracle(ctx, oracle, monkeypatch):
    """FOCR_TC_SPLIT=1: a launch covers its columns as two accumulator-sized sub-blocks per row (two jobs per row that
    share the row's operands).  Same results as the oracle."""
    from font_ocr_b200 import native

    monkeypatch.setenv("FOCR_TC_SPLIT", "1")
    ctx.set_kernel(native.KERNEL_TCGEN05)
    try:
        rng = np.random.default_rng(5)
        page = rng.integers(0, 256, (150, 300), dtype=np.uint8)
        page[rng.random(page.shape) < 0.75] = 255
        tpls = [rng.integers(0, 256, (14, 15), dtype=np.uint8) for _ in range(200)]
        m, c = _scan(ctx, tpls, page, 0.2)
        s = oracle.Searcher(page, "port")
        _assert_same(m[0], c[0], [s.search_c_u8(t, 0.2) for t in tpls], "split")
        assert c.sum() > 0
    finally:
        ctx.set_kernel(native.KERNEL_AUTO)


def test_large_box_bank_on_the_tensor_core(ctx, oracle, font, pkg):
    """BASELINE config 5 shape (-t 24, box about 27x26 > 256 pixels, wider than the AVX2 kernel's 16): the tcgen05
    kernel screens with templates scaled by 2^-sshift and the exact pass restores the reference arithmetic; the oracle
    here is the C restatement with the width limit lifted (the compiled reference panics for n_w > 16, ncc.rs:392)."""
    from font_ocr_b200 import native

    bank = pkg.raster.TemplateBank(font, 24, x_bits=1)
    tpls = [t.pixels for t in bank.templates]
    assert max(t.shape[0] * t.shape[1] for t in tpls) > 256 and max(t.shape[1] for t in tpls) > 16
    page, lines, _ = pkg.pages.make_ncc_page(bank, 900, 260, seed=2, shifts="bank")
    ctx.set_kernel(native.KERNEL_TCGEN05)
    try:
        m, c = _scan(ctx, tpls, page, 0.8)
    finally:
        ctx.set_kernel(native.KERNEL_AUTO)
    s = oracle.Searcher(page, "port")
    sample = list(range(0, len(tpls), 7))
    for t in sample:
        exp = s.search_c_u8(tpls[t], 0.8, allow_wide=True)
        _assert_same(m[0, t:t + 1], c[0, t:t + 1], [exp], f"template {t}")
    assert c.sum() > len(lines)


def _noise_batch(rng, n_pages, h, w):
    pages = rng.integers(0, 256, (n_pages, h, w), dtype=np.uint8)
    pages[rng.random(pages.shape) < 0.6] = 255
    return pages


def test_host_scan_slot_reuse_overflow_and_staging(built_lib, oracle):
    """focr_ncc_scan over 45 pages: the ramped two-slot pipeline reuses its slots (chunks of 2, 4, 8, 16, 15 pages), a fresh
    context's candidate and hit lists overflow (every window is a hit at thr -1.5) so the scan is repeated with larger
    lists, and the caller's buffers are pageable (library-owned pinned staging) or pinned (direct DMA).  All three
    must be byte-identical to focr_ncc_scan_device, and a sample of (page, template) lists to the oracle."""
    import torch
    from font_ocr_b200 import native, ncc

    rng = np.random.default_rng(123)
    P, H, W, n_out = 45, 200, 300, 64
    pages = _noise_batch(rng, P, H, W)
    tpls = [rng.integers(0, 256, (9, 9), dtype=np.uint8) for _ in range(40)]
    T = len(tpls)
    c = ncc.Context(0)   # fresh: default list capacities
    try:
        bank = ncc.Bank(c, tpls)
        m_page, c_page = ncc.scan_pages(c, bank, pages, -1.5, n_out)          # pageable numpy buffers, overflow + redo
        assert (c_page == n_out).all()
        m_page2, c_page2 = ncc.scan_pages(c, bank, pages, 0.3, n_out)         # lists already grown: no redo
        pin = torch.from_numpy(pages).pin_memory()
        out_pin = torch.zeros(P * T * n_out * 8, dtype=torch.uint8).pin_memory()
        cnt_pin = torch.zeros(P * T, dtype=torch.int32).pin_memory()
        m_pin = out_pin.numpy().view(native.MATCH_DTYPE).reshape(P, T, n_out)
        c_pin = cnt_pin.numpy().view(np.uint32).reshape(P, T)
        ncc.scan_pages(c, bank, pin.numpy(), 0.3, n_out, out=m_pin, counts=c_pin)
        dev = pin.cuda()
        out_dev = torch.zeros(P * T * n_out * 8, dtype=torch.uint8, device="cuda")
        cnt_dev = torch.zeros(P * T, dtype=torch.int32, device="cuda")
        ncc.scan_pages_device(c, bank, dev.data_ptr(), H * W, W, W, H, P, 0.3, n_out, out_dev.data_ptr(), cnt_dev.data_ptr())
        m_dev = out_dev.cpu().numpy().view(native.MATCH_DTYPE).reshape(P, T, n_out)
        c_dev = cnt_dev.cpu().numpy().view(np.uint32).reshape(P, T)
        assert np.array_equal(c_dev, c_page2) and np.array_equal(c_dev, c_pin)
        for p in range(P):
            for t in range(T):
                n = int(c_dev[p, t])
                assert m_dev[p, t, :n].tobytes() == m_page2[p, t, :n].tobytes() == m_pin[p, t, :n].tobytes(), (p, t)
        for p in (0, 1, 7, 30, 44):   # first chunk, slot reuse, last (partial) chunk
            s = oracle.Searcher(pages[p], "port")
            for t in (0, 17, 39):
                _assert_same(m_page2[p, t:t + 1], c_page2[p, t:t + 1], [s.search_c_u8(tpls[t], 0.3, n_out=n_out)], f"page {p} tpl {t}")
                _assert_same(m_page[p, t:t + 1], c_page[p, t:t + 1], [s.search_c_u8(tpls[t], -1.5, n_out=n_out)], f"page {p} tpl {t} thr -1.5")
        bank.close()
    finally:
        c.close()


def test_host_register_pins_caller_buffers(built_lib):
    """focr_pin_register / focr_pin_alloc (the page-locking a Rust caller has no CUDA binding for): a registered numpy
    buffer is seen as pinned by the staging decision, scans of registered buffers are byte-identical to staged ones, and
    unregistering restores the pageable path."""
    import ctypes as C
    from font_ocr_b200 import native, ncc

    rng = np.random.default_rng(5)
    P, H, W, n_out = 6, 120, 200, 32
    pages = _noise_batch(rng, P, H, W)
    tpls = [rng.integers(0, 256, (9, 9), dtype=np.uint8) for _ in range(12)]
    c = ncc.Context(0)
    try:
        lib = native.lib()
        is_pinned = getattr(lib, "_Z25focr_internal_host_pinnedPKv")   # the staging decision's own test (C++ linkage, internal)
        is_pinned.argtypes, is_pinned.restype = [C.c_void_p], C.c_bool
        bank = ncc.Bank(c, tpls)
        m0, c0 = ncc.scan_pages(c, bank, pages, 0.3, n_out)                     # pageable: staged
        out = np.zeros((P, len(tpls), n_out), native.MATCH_DTYPE)
        cnt = np.zeros((P, len(tpls)), np.uint32)
        assert not is_pinned(native.ptr(pages))
        with c.pin(pages), c.pin(out), c.pin(cnt):
            assert is_pinned(native.ptr(pages)) and is_pinned(native.ptr(out))
            ncc.scan_pages(c, bank, pages, 0.3, n_out, out=out, counts=cnt)       # direct DMA
        assert not is_pinned(native.ptr(pages))
        assert np.array_equal(cnt, c0)
        for p in range(P):
            for t in range(len(tpls)):
                n = int(c0[p, t])
                assert out[p, t, :n].tobytes() == m0[p, t, :n].tobytes(), (p, t)
        # pinned memory straight from the library
        buf = C.c_void_p()
        native.check(lib.focr_pin_alloc(c._h, pages.nbytes, C.byref(buf)))
        assert is_pinned(buf)
        C.memmove(buf, native.ptr(pages), pages.nbytes)
        view = np.ctypeslib.as_array(C.cast(buf, C.POINTER(C.c_uint8)), shape=(pages.nbytes,)).reshape(pages.shape)
        m1, c1 = ncc.scan_pages(c, bank, view, 0.3, n_out)
        assert np.array_equal(c1, c0) and all(m1[p, t, :int(c0[p, t])].tobytes() == m0[p, t, :int(c0[p, t])].tobytes()
                                              for p in range(P) for t in range(len(tpls)))
        del view
        native.check(lib.focr_pin_free(c._h, buf))
        assert lib.focr_pin_register(c._h, None, 16) != 0 and lib.focr_pin_unregister(c._h, native.ptr(pages)) != 0
        assert b"unregister" in lib.focr_last_error().lower() or b"registered" in lib.focr_last_error().lower()
        # a reported CUDA error must not resurface in the next call's launch checks
        m2, c2 = ncc.scan_pages(c, bank, pages, 0.3, n_out)
        assert np.array_equal(c2, c0)
        bank.close()
    finally:
        c.close()


def test_multi_device_scan_identical_to_single(ctx, oracle):
    """focr_multi_ncc_scan shards ONE batch by page over the contexts of a focr_multi (every visible GPU; on a one-GPU box two
    contexts on device 0, which exercises the same host threads, page blocks and gather) and must return exactly what the
    single-device call returns.  The batch has a tall box (window_stats needs its > 48 KB shared-memory opt-in on every
    device) and fewer pages than twice the device count in one case (blocks of 1 page, blocks of 0 pages)."""
    import torch
    from font_ocr_b200 import ncc

    n_gpu = torch.cuda.device_count()
    devices = list(range(n_gpu)) if n_gpu > 1 else [0, 0]
    rng = np.random.default_rng(9)
    tpls = [rng.integers(0, 256, (40, 20), dtype=np.uint8) for _ in range(5)] + \
           [rng.integers(0, 256, (14, 15), dtype=np.uint8) for _ in range(70)]
    bank1 = ncc.Bank(ctx, tpls)
    every = ncc.MultiContext()   # devices == NULL, n_devices == 0: one context per visible GPU
    assert every.size == n_gpu
    every.close()
    mctx = ncc.MultiContext(devices=devices)
    try:
        assert mctx.size == len(devices)
        mbank = ncc.MultiBank(mctx, tpls)
        for P in (len(devices) * 9 + 1, 1, len(devices) + 1):
            pages = _noise_batch(rng, P, 150, 260)
            blocks = [mctx.page_block(P, i) for i in range(mctx.size)]
            assert sum(b[1] for b in blocks) == P and blocks[0][0] == 0
            assert all(blocks[i][0] + blocks[i][1] == blocks[i + 1][0] for i in range(len(blocks) - 1))
            m1, c1 = ncc.scan_pages(ctx, bank1, pages, 0.25)
            m2, c2 = ncc.scan_pages_multi(mctx, mbank, pages, 0.25)
            assert np.array_equal(c1, c2) and c1.sum() > 0
            for p in range(P):
                for t in range(len(tpls)):
                    n = int(c1[p, t])
                    assert m1[p, t, :n].tobytes() == m2[p, t, :n].tobytes(), (P, p, t)
        s = oracle.Searcher(pages[0], "port")
        for t in (0, 4, 5, 74):
            _assert_same(m2[0, t:t + 1], c2[0, t:t + 1], [s.search_c_u8(tpls[t], 0.25, allow_wide=True)], f"tpl {t}")
        mbank.close()
    finally:
        mctx.close()
        bank1.close()


def test_config2_exact(ctx, oracle, font, pkg):
    """BASELINE config 2 exactly: one 608x800 page, -t 13, --x-bits 2 --y-bits 2 (16 offsets x 74 letters = 1184 templates),
    threshold 0.8: every match list against the compiled reference kernel (oracle/_ref; the C port when it is absent),
    and the decoded text against the oracle's process_hits on those lists (anchor 0.95, overlap 5)."""
    from font_ocr_b200 import ncc

    bank_h = pkg.raster.TemplateBank(font, 13, x_bits=2, y_bits=2)
    tpls = [t.pixels for t in bank_h.templates]
    assert len(tpls) == 1184
    page, truth, _ = pkg.pages.make_ncc_page(bank_h, 608, 800, seed=0, shifts="bank")
    bank = ncc.Bank(ctx, tpls)
    m, c = ncc.scan_pages(ctx, bank, page, 0.8)
    bank.close()
    impl = "reference" if oracle.ref_lib() is not None else "port"
    exp = oracle.get_hits(page, tpls, 0.8, impl)
    _assert_same(m[0], c[0], exp, "config 2")
    letters = bank_h.letters()
    ours = ncc.lines_to_text(ncc.host_process_hits(ncc.get_hits(m[0], c[0], letters), 0.95, 5))
    theirs = oracle.lines_to_text(oracle.process_hits(oracle.hits_with_letters(exp, letters), 0.95, 5))
    assert ours == theirs and len(ours) > 0


def test_config5_full_page_sample(ctx, oracle, font, pkg):
    """BASELINE config 5 at its named shape: 95 printable-ASCII glyphs, -t 24, --x-bits 3 --y-bits 2 (3040 templates, boxes
    wider than the AVX2 kernel's 16) on a full 2480x3508 page; a template sample against the wide C port (the compiled
    reference panics for n_w > 16, ncc.rs:392), size-independent properties for the rest."""
    from concurrent.futures import ThreadPoolExecutor
    from font_ocr_b200 import ncc

    alphabet = "".join(chr(ch) for ch in range(32, 127))
    bank_h = pkg.raster.TemplateBank(font, 24, x_bits=3, y_bits=2, alphabet=alphabet)
    tpls = [t.pixels for t in bank_h.templates]
    assert len(tpls) == 3040
    page = pkg.pages.make_ncc_page(bank_h, 2480, 3508, seed=11, shifts="bank")[0]
    bank = ncc.Bank(ctx, tpls)
    m, c = ncc.scan_pages(ctx, bank, page, 0.8)
    bank.close()
    assert (c <= 1024).all() and c.sum() > 0
    for t in range(0, len(tpls), 13):
        g = m[0, t, :c[0, t]]
        key = g["y"].astype(np.int64) * 65536 + g["x"]
        assert (np.diff(key) > 0).all() and (g["similarity"] > np.float32(0.8) - 1e-6).all(), t
    space = [t for t in range(len(tpls)) if not tpls[t].any()]
    assert space and all(c[0, t] == 0 for t in space)   # the all-zero space template: NaN similarity, no hits
    sample = [1, 700, 1519, 2222, 3039]

    def one(t):   # one Searcher per thread: prepare_for_size caches per instance
        so = oracle.Searcher(page, "port")
        return so.search_c_u8(tpls[t], 0.8, allow_wide=True)

    with ThreadPoolExecutor(len(sample)) as ex:
        exps = list(ex.map(one, sample))
    for t, exp in zip(sample, exps):
        _assert_same(m[0, t:t + 1], c[0, t:t + 1], [exp], f"template {t}")
