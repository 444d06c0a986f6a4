"""Generate tests/golden/ncc_golden.npz from the reference's OWN compiled kernel.

Run in the build container (needs /root/reference to have produced oracle/_ref/libncc_ref.so via
`make -C oracle`):   python tests/golden/make_golden.py

The reference repository has no tests, fixtures or golden vectors, so these
vectors are outputs of the unmodified `ncc_8_u8` / `ncc_16_u8` (src/ncc.cpp) driven through its FFI
signature (ncc.rs:92-126).  Inputs (pages, templates) are stored next to the expected outputs so the
fixtures do not depend on the FreeType version that rendered them.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import font_ocr_b200 as pkg  # noqa: E402
from oracle import oracle as O  # noqa: E402


def pack(hits):
    counts = np.array([len(h) for h in hits], np.uint32)
    flat = np.concatenate(hits) if hits else np.zeros(0, O.MATCH_DTYPE)
    return counts, flat


def main():
    assert O.ref_lib() is not None, "oracle/_ref/libncc_ref.so missing: run `make -C oracle` first"
    out = {}
    font = pkg.raster.Font()
    rng = np.random.default_rng(1234)

    # case text16: BASELINE config 1 shape -- 608x800 base64 page, -t 13 (15x14 box -> ncc_16_u8)
    bank = pkg.raster.TemplateBank(font, 13)
    page, lines, _ = pkg.pages.make_ncc_page(bank, 608, 800, seed=0)
    tpl = np.stack([t.pixels for t in bank.templates])
    c, f = pack(O.get_hits(page, list(tpl), 0.8, "reference"))
    out.update(text16_page=page, text16_tpl=tpl, text16_thr=np.float32(0.8), text16_counts=c, text16_hits=f,
               text16_lines=np.array(lines), text16_letters=np.array(bank.letters()))

    # case text8: -t 6 (8x7 box -> ncc_8_u8) on a smaller page
    bank8 = pkg.raster.TemplateBank(font, 6)
    page8, lines8, _ = pkg.pages.make_ncc_page(bank8, 304, 200, seed=1, margin_x=11, margin_y=9)
    tpl8 = np.stack([t.pixels for t in bank8.templates])
    c, f = pack(O.get_hits(page8, list(tpl8), 0.8, "reference"))
    out.update(text8_page=page8, text8_tpl=tpl8, text8_thr=np.float32(0.8), text8_counts=c, text8_hits=f)

    # case noise: random page, random templates of assorted sizes, low threshold, small n_out so the
    # early return (ncc.cpp:225-227) fires on some templates and not on others
    pagen = rng.integers(0, 256, (61, 97), dtype=np.uint8)
    pagen[20:40, 30:70] = 255  # a blank region: constant windows -> rnorm = inf -> never a hit
    sizes = [(5, 4), (8, 8), (13, 9), (16, 16), (3, 1), (16, 3), (9, 7)]
    n_out = 48
    tpls = [rng.integers(0, 256, (h, w), dtype=np.uint8) for (w, h) in sizes]
    pagen[5:5 + 9, 50:50 + 13] = 255 - tpls[2]  # plant one template so there is a perfect match
    for i, t in enumerate(tpls):
        s = O.Searcher(pagen, "reference")
        hits = s.search_c_u8(t, 0.25, n_out=n_out)
        out[f"noise_tpl{i}"] = t
        out[f"noise_hits{i}"] = hits
    out.update(noise_page=pagen, noise_thr=np.float32(0.25), noise_n_out=np.uint32(n_out),
               noise_n=np.uint32(len(sizes)))

    # case dense: a page that is a tiling of one glyph -> far more than 1024 hits -> truncation at 1024
    g = bank.templates[bank.alphabet.index("H")].pixels
    ink = np.zeros((300, 400), np.uint8)
    for y in range(3, 300 - 17, 17):
        for x in range(2, 400 - 16, 16):
            ink[y:y + g.shape[0], x:x + g.shape[1]] = g
    paged = (255 - ink).astype(np.uint8)
    s = O.Searcher(paged, "reference")
    hits = s.search_c_u8(g, 0.3)
    assert len(hits) == 1024
    out.update(dense_page=paged, dense_tpl=g, dense_thr=np.float32(0.3), dense_hits=hits)

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncc_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;",
          {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items() if "hits" in k or "counts" in k})


if __name__ == "__main__":
    main()
