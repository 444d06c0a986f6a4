"""tcgen05.mma kind::i8 micro-benchmark driver: cycles per MMA and achieved TOP/s vs N and accumulators."""
import ctypes as C, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.microbench import microbench as native

lib = native.lib()
res = []
for n in (64, 80, 112, 128, 160, 224, 256):
    for nacc in (1, 2):
        if nacc * ((n + 31) // 32 * 32) > 512:
            continue
        cyc, ms = np.zeros(1), np.zeros(1)
        native.check(lib.focr_bench_umma_i8(0, n, 7, 4000, nacc, native.ptr(cyc), native.ptr(ms)))
        ops = 2.0 * 128 * n * 32 * 7 * 4000 * 148
        res.append({"n": n, "nacc": nacc, "cycles_per_mma": float(cyc[0]), "ideal_cycles": n / 2,
                    "tops": ops / (ms[0] * 1e-3) / 1e12})
        print(res[-1])
json.dump(res, open("gpurun_out/umma_i8_bench.json", "w"), indent=1)
