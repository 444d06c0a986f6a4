"""Event timeline of the tcgen05 scan kernel's CTA 0 (experiments build, FOCR_TC_TRACE + FOCR_TC_TIMELINE=first job):
clock64 stamps of the issuing warp and of three epilogue warps for 256 consecutive jobs."""
import glob, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["FOCR_TC_TRACE"] = "gpurun_out/tc_tl"
os.environ.setdefault("FOCR_TC_TIMELINE", "2000")
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
names = {0: ("issuer 0", ["wait t_empty", "go", "issued", "committed"]),
         1: ("issuer 1", ["wait t_empty", "go", "issued", "committed"]),
         2: ("epi e0 (team 0)", ["seen", "2 chunks in", "released", "done"]),
         3: ("epi e4 (team 1)", ["seen", "2 chunks in", "released", "done"])}
for fn in sorted(glob.glob("gpurun_out/tc_tl.*")):
    t = np.fromfile(fn, dtype=np.int64)[64:].reshape(4, 256, 4)
    base = t[0, 8, 0]
    print(fn)
    for j in range(8, 32):
        row = [f"job {j:3d}"]
        for r in range(4):
            if t[r, j].any():
                row.append(names[r][0] + " " + " ".join(f"{int(v - base):6d}" if v else "     -" for v in t[r, j]))
        print(" | ".join(row))
    for r in (0, 1):
        iss = t[r, 8:250]
        iss = iss[iss[:, 0] != 0]
        print("%s per own job: period %.0f  wait %.0f  go->issued %.0f  commit %.0f" % (
            names[r][0], np.diff(iss[:, 0]).mean(), (iss[:, 1] - iss[:, 0]).mean(), (iss[:, 2] - iss[:, 1]).mean(), (iss[:, 3] - iss[:, 2]).mean()))
    for r in (2, 3):
        e = t[r, 8:250]
        e = e[e[:, 0] != 0]
        print("%s per own job: period %.0f  seen->2 chunks in %.0f  ->released %.0f  ->done %.0f  done->next seen %.0f" % (
            names[r][0], np.diff(e[:, 0]).mean(), (e[:, 1] - e[:, 0]).mean(), (e[:, 2] - e[:, 1]).mean(), (e[:, 3] - e[:, 2]).mean(),
            (e[1:, 0] - e[:-1, 3]).mean()))
    jobs = np.arange(8, 240)
    for r in (2, 3):
        own = [j for j in jobs if t[r, j, 0] != 0]
        lat = [t[r, j, 0] - t[j & 1, j, 3] for j in own]
        print("%s: commit -> seen %.0f (min %.0f)" % (names[r][0], np.mean(lat), np.min(lat)))
        lat = [t[(j + 3) & 1, j + 3, 1] - t[r, j, 2] for j in own]
        print("%s: released -> issuer go (job + 3) %.0f (min %.0f)" % (names[r][0], np.mean(lat), np.min(lat)))
