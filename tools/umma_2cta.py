"""cycles per tcgen05.mma kind::i8 as a function of N: cta_group::1 (M = 128) against cta_group::2 (M = 256, issued by the
leader of a CTA pair, each CTA holding N/2 columns of B).  Writes gpurun_out/umma_2cta.json."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "microbench"))
import microbench as mb

mb.build()
lib = mb.lib()
res = []
for n in (64, 96, 128, 144, 160, 192, 224, 256):
    row = {"n": n}
    for name, fn in (("cta1", lib.focr_bench_umma_i8), ("cta2", lib.focr_bench_umma_i8_2cta)):
        cyc, ms = np.zeros(1), np.zeros(1)
        mb.check(fn(0, n, 7, 4000, 1, mb.ptr(cyc), mb.ptr(ms)))
        row[name + "_cycles_per_mma"] = round(float(cyc[0]), 1)
    row["ideal_cycles"] = n / 2
    res.append(row)
    print(row, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/umma_2cta.json", "w"), indent=1)
