"""How far ahead of the tensor pipe may one thread issue tcgen05.mma?  (issue time vs completion time for n MMAs)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.microbench import microbench as native
lib = native.lib()
for n in (224, 96):
    for ksteps, iters in ((1, 1), (2, 1), (4, 1), (8, 1), (8, 2), (8, 4), (8, 8), (8, 32)):
        cyc, ms = np.zeros(1), np.zeros(1)
        native.check(lib.focr_bench_umma_i8(0, n, ksteps, iters, 2, native.ptr(cyc), native.ptr(ms)))
        tot = cyc[0] * ksteps * iters
        print(f"N={n} MMAs={ksteps * iters:4d}: issue done after {lib.focr_bench_umma_issue_cycles():8.0f} cycles, complete after {tot:8.0f}", flush=True)
