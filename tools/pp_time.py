import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from font_ocr_b200 import native, ncc
pkg, font, bank_h = bench.make_bank()
tpls = [t.pixels for t in bank_h.templates]; T = len(tpls)
ctx = ncc.Context(0); bank = ncc.Bank(ctx, tpls)
for P in (16, 40, 100):
    pages = torch.from_numpy(bench.make_pages(pkg, bank_h, P, 0, distinct=8)).cuda()
    out = torch.empty(P * T * 1024 * 8, dtype=torch.uint8, device="cuda"); cnt = torch.empty(P * T, dtype=torch.int32, device="cuda")
    ncc.scan_pages_device(ctx, bank, pages.data_ptr(), bench.R_W * bench.R_H, bench.R_W, bench.R_W, bench.R_H, P, 0.8, 1024, out.data_ptr(), cnt.data_ptr())
    ctx.sync()
    for it in range(3):
        t0 = time.perf_counter()
        r = ncc.process_hits_device(ctx, out.data_ptr(), cnt.data_ptr(), T, 1024, P, bank_h.letters(), 0.95, 5, raw=True)
        dt = time.perf_counter() - t0
        print(P, it, "ms/page", 1e3 * dt / P, "lines", len(r[0]), "sel", len(r[2]), flush=True)
    del pages, out, cnt
