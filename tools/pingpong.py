"""Hand-shake round-trip latency on one SM: signal -> waiting warps answer -> signaller waits (cycles)."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.microbench import microbench as native
lib = native.lib()
res = []
for use_commit in (0, 1):
    for kind, name in ((0, "try_wait+20us hint"), (1, "try_wait"), (2, "test_wait poll")):
        for nwait in (1, 8):
            c = np.zeros(1)
            native.check(lib.focr_bench_pingpong(0, kind, use_commit, nwait, 20000, native.ptr(c)))
            r = {"signal": "tcgen05.commit" if use_commit else "mbarrier.arrive", "wait": name, "waiting_warps": nwait,
                 "cycles_per_round_trip": float(c[0])}
            res.append(r)
            print(r, flush=True)
json.dump(res, open("gpurun_out/pingpong.json", "w"), indent=1)
