import os, sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import bench
from font_ocr_b200 import native, ncc
pkg, font, bank_h = bench.make_bank()
tpls = [t.pixels for t in bank_h.templates]
T, P = len(tpls), 100
ctx = ncc.Context(0)
bank = ncc.Bank(ctx, tpls)
pages = bench.make_pages(pkg, bank_h, P, 0, distinct=8)
out = np.zeros((P, T, 1024), native.MATCH_DTYPE); cnt = np.zeros((P, T), np.uint32)
for _ in range(2):
    ncc.scan_pages(ctx, bank, pages, 0.8, 1024, out=out, counts=cnt)
ts = []
for _ in range(4):
    t0 = time.perf_counter(); ncc.scan_pages(ctx, bank, pages, 0.8, 1024, out=out, counts=cnt); ts.append(time.perf_counter() - t0)
print("threads", os.environ.get("FOCR_STAGE_THREADS"), "pageable pages/s", P / np.median(ts), flush=True)
