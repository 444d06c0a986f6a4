"""Timing of focr_decode_pages (BASELINE config 4 geometry) for P pages: wall clock per call at two batch sizes."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from font_ocr_b200 import native, ncc, focr

pkg, font, bank_h = bench.make_bank()
ctx = ncc.Context(0)
fbank = focr.GlyphBank(ctx, font, bench.TEXT_SIZE)
R_W, R_H = bench.R_W, bench.R_H
distinct = [pkg.pages.make_focr_page(font, bench.TEXT_SIZE, R_W, R_H, seed=7000 + i)[0] for i in range(4)]
for P in (int(a) for a in (sys.argv[1:] or ["32", "8"])):
    pages = np.stack([distinct[i % 4] for i in range(P)])
    fmax = (R_H - 39 + 14) // 15
    g = np.zeros((P, fmax, 512), np.uint16)
    n, y, l = np.zeros((P, fmax), np.uint32), np.zeros((P, fmax), np.uint32), np.zeros(P, np.uint32)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        native.check(native.lib().focr_decode_pages(ctx._h, fbank._h, native.ptr(pages), R_W * R_H, R_W, R_H, P, 45, 39, 608, 12, 15,
                                                    fmax, 512, native.ptr(g), native.ptr(n), native.ptr(y), native.ptr(l)))
        ts.append(time.perf_counter() - t0)
    print(f"P={P}: {1e3 * min(ts[1:]):.2f} ms per call, {P / min(ts[1:]):.0f} pages/s, lines {int(l[0])}")
