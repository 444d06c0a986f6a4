"""Small workload for compute-sanitizer (memcheck): every kernel of the library on small shapes --
the tcgen05 scan with one and two box sizes per launch, packed 8-wide boxes, boxes wider than 16, the SIMT scan,
statistics / finalize / exact pass, device process_hits, both focr decode kernels (dense tiles and row tasks)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import font_ocr_b200 as pkg
from font_ocr_b200 import focr, native, ncc

rng = np.random.default_rng(1)
ctx = ncc.Context(0)
page = rng.integers(0, 256, (2, 150, 300), dtype=np.uint8)
page[rng.random(page.shape) < 0.6] = 255
banks = {
    "two box sizes": [rng.integers(0, 256, (14, 15), dtype=np.uint8) for _ in range(40)] + [rng.integers(0, 256, (14, 14), dtype=np.uint8) for _ in range(20)],
    "packed 8-wide": [rng.integers(0, 256, (7, 8), dtype=np.uint8) for _ in range(33)] + [rng.integers(0, 256, (7, 5), dtype=np.uint8) for _ in range(9)],
    "wide": [rng.integers(0, 256, (25, 27), dtype=np.uint8) for _ in range(70)],
}
for kernel in (native.KERNEL_TCGEN05, native.KERNEL_SIMT):
    ctx.set_kernel(kernel)
    for name, tpls in banks.items():
        bank = ncc.Bank(ctx, tpls)
        m, c = ncc.scan_pages(ctx, bank, page, 0.3)
        print(kernel, name, int(c.sum()), flush=True)
        bank.close()
ctx.set_kernel(native.KERNEL_AUTO)
font = pkg.raster.Font()
bank_h = pkg.raster.TemplateBank(font, 13, x_bits=1)
tpls = [t.pixels for t in bank_h.templates]
pages = np.stack([pkg.pages.make_ncc_page(bank_h, 608, 300, seed=s, shifts="bank")[0] for s in range(2)])
bank = ncc.Bank(ctx, tpls)
T = len(tpls)
dev = torch.from_numpy(pages).cuda()
out = torch.zeros(2 * T * 1024 * 8, dtype=torch.uint8, device="cuda")
cnt = torch.zeros(2 * T, dtype=torch.int32, device="cuda")
ncc.scan_pages_device(ctx, bank, dev.data_ptr(), 608 * 300, 608, 608, 300, 2, 0.8, 1024, out.data_ptr(), cnt.data_ptr())
lines = ncc.process_hits_device(ctx, out.data_ptr(), cnt.data_ptr(), T, 1024, 2, bank_h.letters(), 0.95, 5)
print("process_hits", [len(l) for l in lines], flush=True)
bank.close()
fpage = np.stack([pkg.pages.make_focr_page(font, 13, 700, 39 + 15 * 5 + 7, seed=5, fill=1.0)[0]] * 2)
fbank = focr.GlyphBank(ctx, font, 13)
for legacy in ("", "1"):
    if legacy:
        os.environ["FOCR_DECODE_LEGACY"] = "1"
    r = focr.decode_images(ctx, fbank, fpage, 45, 39, 608, 12, 15)
    print("focr", "legacy" if legacy else "tiles", len(r[0]), flush=True)
fbank.close()
ctx.close()
print("done")
