"""Debug helper (not a test): numerators and match lists of the tcgen05 kernel vs the SIMT kernel."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import font_ocr_b200 as pkg
from font_ocr_b200 import native, ncc

font = pkg.raster.Font()
size = float(sys.argv[1]) if len(sys.argv) > 1 else 13
bank_h = pkg.raster.TemplateBank(font, size)
tpls = [t.pixels for t in bank_h.templates]
page, lines, _ = pkg.pages.make_ncc_page(bank_h, 608, 400, seed=0)
ctx = ncc.Context(0)
bank = ncc.Bank(ctx, tpls)
for t in (0, 5, 73):
    ctx.set_kernel(native.KERNEL_SIMT)
    a0 = ncc.numerators(ctx, bank, t, page)
    ctx.set_kernel(native.KERNEL_TCGEN05)
    a1 = ncc.numerators(ctx, bank, t, page)
    n_h, n_w = tpls[t].shape
    ys, xs = page.shape[0] - n_h + 1, page.shape[1] - n_w + 1
    d = a0[1:ys, 1:xs] != a1[1:ys, 1:xs]
    print(f"template {t}: numerators differ at {int(d.sum())} of {d.size} windows; simt sum {int(a0.sum())} tc sum {int(a1.sum())}")
    if d.any():
        yy, xx = np.nonzero(d)
        print("  first diffs (y,x,simt,tc):", [(int(y) + 1, int(x) + 1, int(a0[y + 1, x + 1]), int(a1[y + 1, x + 1])) for y, x in zip(yy[:8], xx[:8])])
        print("  rows with diffs:", np.unique(yy)[:20] + 1, " cols:", np.unique(xx)[:20] + 1)
ctx.set_kernel(native.KERNEL_SIMT)
m0, c0 = ncc.scan_pages(ctx, bank, page, 0.8)
ctx.set_kernel(native.KERNEL_TCGEN05)
m1, c1 = ncc.scan_pages(ctx, bank, page, 0.8)
print("counts equal:", np.array_equal(c0, c1), int(c0.sum()), int(c1.sum()))
print("lists equal:", m0.tobytes() == m1.tobytes())
if not np.array_equal(c0, c1):
    bad = np.nonzero(c0[0] != c1[0])[0]
    print("templates with different counts:", bad[:20], c0[0][bad[:20]], c1[0][bad[:20]])
if not np.array_equal(c0, c1) or m0.tobytes() != m1.tobytes():
    t = int(np.nonzero(c0[0] != c1[0])[0][0]) if not np.array_equal(c0, c1) else 0
    a, b = m0[0, t, :c0[0, t]], m1[0, t, :c1[0, t]]
    print("template", t, "simt", len(a), "tc", len(b))
    sa = {(int(r["x"]), int(r["y"])) for r in a}
    extra = [(int(r["x"]), int(r["y"]), float(r["similarity"])) for r in b if (int(r["x"]), int(r["y"])) not in sa]
    print("extra in tc (first 12):", extra[:12])
    keys = [(int(r["y"]), int(r["x"])) for r in b]
    print("duplicates in tc:", len(keys) - len(set(keys)))
