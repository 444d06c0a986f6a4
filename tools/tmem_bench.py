"""TMEM load/store micro-benchmark driver: cycles per tcgen05.ld/st 32x32b.x32 round vs warps, with and without MMAs."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.microbench import microbench as native

lib = native.lib()
res = []
for mma_n in (0, 224, 96):
    for mode in (0, 1, 2, 3):
        for nw in (4, 8, 12, 16):
            a, b = np.zeros(1), np.zeros(1)
            iters = 4000
            # keep the tensor core busy for about as long as the ld/st warps run
            native.check(lib.focr_bench_tmem(0, nw, mode, iters, mma_n, 3000, native.ptr(a), native.ptr(b)))
            r = {"mma_n": mma_n, "mode": ["ld", "ld+st", "st", "ld+max"][mode], "warps": nw, "cycles_per_round_per_warp": float(a[0]),
                 "cycles_per_unit_per_quarter": float(a[0]) / (nw / 4), "cycles_per_mma": float(b[0])}
            res.append(r)
            print(r, flush=True)
json.dump(res, open("gpurun_out/tmem_bench.json", "w"), indent=1)
