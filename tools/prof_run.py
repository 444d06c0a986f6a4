"""Small fixed workload for the ncu captures committed under profiles/ (one GPU):
config 3 (16 pages 2480x3508, 296 templates) twice through focr_ncc_scan_device, then config 4 (focr, 8 pages)
twice through focr_decode_pages.  Prints the CUDA-event stage times of the second scan."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from font_ocr_b200 import focr, native, ncc

P = int(sys.argv[1]) if len(sys.argv) > 1 else 16
THR = float(sys.argv[2]) if len(sys.argv) > 2 else 0.8
pkg, font, bank_h = bench.make_bank()
tpls = [t.pixels for t in bank_h.templates]
T = len(tpls)
ctx = ncc.Context(0)
bank = ncc.Bank(ctx, tpls)
pages = torch.from_numpy(bench.make_pages(pkg, bank_h, P, 0, distinct=8)).cuda()
out = torch.empty(P * T * 1024 * 8, dtype=torch.uint8, device="cuda")
cnt = torch.empty(P * T, dtype=torch.int32, device="cuda")
for i in range(2):
    if i == 1:
        ctx.profile(True); ctx.profile_read()
    ncc.scan_pages_device(ctx, bank, pages.data_ptr(), bench.R_W * bench.R_H, bench.R_W, bench.R_W, bench.R_H, P, THR, 1024,
                          out.data_ptr(), cnt.data_ptr())
pr = ctx.profile_read(); ctx.profile(False)
print("scan", {k: round(v[0] / P, 4) for k, v in pr.items()}, "ms/page", flush=True)
fbank = focr.GlyphBank(ctx, font, 13)
fpages = np.stack([pkg.pages.make_focr_page(font, 13, bench.R_W, bench.R_H, seed=7000 + i)[0] for i in range(2)] * 4)
for i in range(2):
    r = focr.decode_images(ctx, fbank, fpages, 45, 39, 608, 12, 15)
print("focr", len(r), "pages", len(r[0]), "lines")
