timeout 900 python -m pytest tests/test_gpu_ncc.py -x -q -m gpu 2>&1 | tail -5
timeout 300 python tools/prof_run.py 16 0.8 2>&1 | grep "scan\|Error" | tail -2
python - <<'PY'
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import font_ocr_b200 as pkg
from font_ocr_b200 import ncc, native
import os
font = pkg.raster.Font()
bank_h = pkg.raster.TemplateBank(font, 7, x_bits=2)
tpls = [t.pixels for t in bank_h.templates]
print("sizes", bank_h.sizes(), len(tpls))
pages = np.stack([pkg.pages.make_ncc_page(bank_h, 2480, 3508, seed=s, shifts="bank")[0] for s in range(8)])
for nopack in (0, 1):
    if nopack: os.environ["FOCR_TC_NOPACK"] = "1"
    ctx = ncc.Context(0); ctx.set_kernel(native.KERNEL_TCGEN05)
    bank = ncc.Bank(ctx, tpls)
    dev = torch.from_numpy(pages).cuda()
    T = len(tpls)
    out = torch.empty(8 * T * 1024 * 8, dtype=torch.uint8, device="cuda"); cnt = torch.empty(8 * T, dtype=torch.int32, device="cuda")
    for i in range(3):
        if i == 2: ctx.profile(True); ctx.profile_read()
        ncc.scan_pages_device(ctx, bank, dev.data_ptr(), 2480 * 3508, 2480, 2480, 3508, 8, 0.8, 1024, out.data_ptr(), cnt.data_ptr())
    pr = ctx.profile_read()
    print("nopack" if nopack else "packed", {k: round(v[0] / 8, 4) for k, v in pr.items()}, "hits", int(cnt.sum()))
    bank.close(); ctx.close()
PY
