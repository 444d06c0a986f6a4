"""Timing experiments on the tcgen05 scan kernel (FOCR_TC_DBG modes; results of modes != 0 are wrong on purpose)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from font_ocr_b200 import native, ncc

modes = [int(m) for m in sys.argv[1].split(",")] if len(sys.argv) > 1 else [0, 1, 2, 4, 61]  # bit mask, see scan_tc.cu (experiments build)
P = int(sys.argv[2]) if len(sys.argv) > 2 else 32
pkg, font, bank_h = bench.make_bank()
tpls = [t.pixels for t in bank_h.templates]
if len(sys.argv) > 3:  # only the templates of one box width
    tpls = [t for t in tpls if t.shape[1] == int(sys.argv[3])]
T = len(tpls)
print("templates", T, "box", tpls[0].shape)
ctx = ncc.Context(0)
bank = ncc.Bank(ctx, tpls)
pages = torch.from_numpy(bench.make_pages(pkg, bank_h, P, 0, distinct=8)).cuda()
out = torch.empty(P * T * 1024 * 8, dtype=torch.uint8, device="cuda")
cnt = torch.empty(P * T, dtype=torch.int32, device="cuda")
for mode in modes:
    os.environ["FOCR_TC_DBG"] = str(mode)
    for i in range(3):
        if i == 2:
            ctx.profile(True); ctx.profile_read()
        ncc.scan_pages_device(ctx, bank, pages.data_ptr(), bench.R_W * bench.R_H, bench.R_W, bench.R_W, bench.R_H, P, 0.8, 1024,
                              out.data_ptr(), cnt.data_ptr())
    pr = ctx.profile_read(); ctx.profile(False)
    tot = sum(v[0] for k, v in pr.items() if k != "exact")
    print("mode", mode, "pages/s %.1f" % (P / tot * 1e3), {k: round(v[0] / P, 4) for k, v in pr.items()}, "ms/page", flush=True)
