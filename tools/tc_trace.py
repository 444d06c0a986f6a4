"""Role time budget of the tcgen05 scan kernel's CTA 0 (FOCR_TC_TRACE): cycles per output row each role spends
in each of its waits / phases."""
import glob, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["FOCR_TC_TRACE"] = "gpurun_out/tc_trace"
import torch
import bench
from font_ocr_b200 import native, ncc

P = 4
pkg, font, bank_h = bench.make_bank()
tpls = [t.pixels for t in bank_h.templates]
T = len(tpls)
ctx = ncc.Context(0)
bank = ncc.Bank(ctx, tpls)
pages = torch.from_numpy(bench.make_pages(pkg, bank_h, P, 0)).cuda()
out = torch.empty(P * T * 1024 * 8, dtype=torch.uint8, device="cuda")
cnt = torch.empty(P * T, dtype=torch.int32, device="cuda")
for i in range(2):
    ncc.scan_pages_device(ctx, bank, pages.data_ptr(), bench.R_W * bench.R_H, bench.R_W, bench.R_W, bench.R_H, P, float(os.environ.get("THR", "0.8")), 1024,
                          out.data_ptr(), cnt.data_ptr())
ctx.sync()
rows = P * 20 * 3494 / 148.0  # output rows per CTA (approx.)
roles = {0: ("mma", ["a_full", "a2_full", "t_empty", "issue", "commits"]),
         8: ("epi0 (per own job = 2 rows' worth / nsub)", ["t_full", "slow screens (cycles)", "slow screens (count)", "seen->released", "seen->job done", "all screens"]),
         16: ("toeplitz0 (per 4 rows)", ["raw_full", "a_empty"]),
         24: ("a2_0 (per 4 rows)", ["a2_empty"]),
         32: ("tma", ["raw_empty"])}
for fn in sorted(glob.glob("gpurun_out/tc_trace.*")):
    t = np.fromfile(fn, dtype=np.int64)
    print(fn, "approx rows per CTA %.0f" % rows)
    for base, (name, slots) in roles.items():
        tot = t[base]
        parts = ", ".join(f"{s} {t[base + 1 + i] / rows:.0f}" for i, s in enumerate(slots) if s != "-")
        print(f"  {name:24s} total {tot / rows:7.0f} cycles/row   waits: {parts}")
