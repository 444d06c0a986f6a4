"""BASELINE config 5 (95 printable-ASCII glyphs, -t 24, --x-bits 3 --y-bits 2: 3040 templates of about 27x26) on a few
2480x3508 pages: pages/s of the device-resident scan and the kernel that ran."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import font_ocr_b200 as pkg
from font_ocr_b200 import native, ncc

P = int(sys.argv[1]) if len(sys.argv) > 1 else 4
font = pkg.raster.Font()
alphabet = "".join(chr(c) for c in range(32, 127))
bank_h = pkg.raster.TemplateBank(font, 24, x_bits=3, y_bits=2, alphabet=alphabet)
tpls = [t.pixels for t in bank_h.templates]
sizes = {}
for t in tpls:
    sizes[t.shape[::-1]] = sizes.get(t.shape[::-1], 0) + 1
print("templates", len(tpls), "box sizes", sizes, flush=True)
T = len(tpls)
ctx = ncc.Context(0)
ctx.set_kernel(native.KERNEL_TCGEN05 if len(sys.argv) < 3 else native.KERNEL_SIMT)
bank = ncc.Bank(ctx, tpls)
pages = torch.from_numpy(np.stack([pkg.pages.make_ncc_page(bank_h, 2480, 3508, seed=i, shifts="bank")[0] for i in range(P)])).cuda()
out = torch.empty(P * T * 1024 * 8, dtype=torch.uint8, device="cuda")
cnt = torch.empty(P * T, dtype=torch.int32, device="cuda")
for it in range(2):
    ctx.profile(True); ctx.profile_read()
    t0 = time.perf_counter()
    ncc.scan_pages_device(ctx, bank, pages.data_ptr(), 2480 * 3508, 2480, 2480, 3508, P, 0.8, 1024, out.data_ptr(), cnt.data_ptr())
    ctx.sync()
    dt = time.perf_counter() - t0
    pr = ctx.profile_read()
    ops = sum(2.0 * t.shape[0] * t.shape[1] * (2480 - t.shape[1]) * (3508 - t.shape[0]) for t in tpls) * P
    print(f"pages/s {P / dt:.2f}  dense TOP/s {ops / dt / 1e12:.0f}  hits {int(cnt.sum())}", {k: round(v[0] / P, 2) for k, v in pr.items()}, "ms/page", flush=True)
