#!/usr/bin/env python
"""bench.py -- font-ocr NCC template scan on B200: pages/sec (and template-window evals/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--pages P]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2], the configuration the pages/sec metric is quoted on): a batch of
synthetic 2480x3508 (300 dpi) pages of base64 text, `-t 13 --x-bits 2` -> 4 subpixel offsets x 74
letters = 296 templates, threshold 0.8, 1024 matches per (page, template).  One STEP = one pass of
the hot path over one batch of P pages per GPU (default P = 100; the batch, 870 MB, is far larger
than the 126 MB L2, so no flush is needed between steps).  Pages shard across ranks with no
data-path collective (weak scaling: every rank scans its own P pages); `strong` (N > 1) is ONE
P-page batch sharded over the N GPUs by the library's multi-device entry (focr_multi_ncc_scan).

Printed JSON (one line, rank 0): `value` = pages/s with the pages already resident in HBM, timed
with CUDA events on the library's stream; `e2e` = the same through the public host-buffer C-ABI call
(pinned host pages -> H2D -> kernels -> D2H of the match lists inside the timed region; `pageable`
beside it for a caller whose buffers are not pinned); `roofline` = the correlation kernel against
the integer tensor pipe, peak = the tcgen05.mma kind::i8 rate measured live on this GPU;
`cpu_baseline` = the reference's own AVX2 kernel (oracle/_ref) on the box's host cores, on a bounded
sample.  `config1`, `config2`, `config5`, `focr` (config 4) are the other BASELINE configs, each
with its own roofline / cpu_baseline, measured beside the headline.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

R_W, R_H = 2480, 3508
TEXT_SIZE, X_BITS, Y_BITS = 13, 2, 0
THRESHOLD, N_OUT = 0.8, 1024
WORKLOAD = "config3: 2480x3508 synthetic base64 pages, -t 13 --x-bits 2 (296 templates), thr 0.8, n_out 1024"
SM_COUNT = 148


def make_bank():
    import font_ocr_b200 as pkg

    font = pkg.raster.Font()
    bank = pkg.raster.TemplateBank(font, TEXT_SIZE, x_bits=X_BITS, y_bits=Y_BITS)
    return pkg, font, bank


def make_pages(pkg, bank, n, seed0, distinct=None):
    """n pages; `distinct` (default all) are generated, the rest repeat them cyclically."""
    distinct = n if distinct is None else min(distinct, n)
    base = [pkg.pages.make_ncc_page(bank, R_W, R_H, seed=seed0 + i, shifts="bank")[0] for i in range(distinct)]
    return np.stack([base[i % distinct] for i in range(n)])


def class_table(tpls):
    """[(n_w, n_h, n_templates)] per box size."""
    d = {}
    for t in tpls:
        k = t.shape[::-1]
        d[k] = d.get(k, 0) + 1
    return [(w, h, c) for (w, h), c in sorted(d.items())]


def dense_windows(n_w, n_h, r_w=R_W, r_h=R_H):
    return (r_w - n_w) * (r_h - n_h)  # x in [1, r_w-n_w], y in [1, r_h-n_h]


def effective_window_fraction(page, n_w, n_h):
    """W_eff / W_dense for one page: the windows the reference actually scores (ncc.rs:279-314)."""
    inv = (255 - page).astype(np.int64)
    c = np.zeros((inv.shape[0] + 1, inv.shape[1] + 1), np.int64)
    c[1:, 1:] = inv.cumsum(0).cumsum(1)
    sp = c[n_h:, n_w:] - c[:-n_h, n_w:] - c[n_h:, :-n_w] + c[:-n_h, :-n_w]
    nz = sp[1:, 1:] != 0
    any_ = nz.any(1)
    first = nz.argmax(1)
    last = nz.shape[1] - 1 - nz[:, ::-1].argmax(1)
    w_eff = int(((last - first + 1) * any_).sum())
    return w_eff / float(nz.size)


def algorithmic_ops(tpls, pages, n_pages_total):
    """(ops over W_eff windows, ops over dense windows, mean W_eff/W_dense) for scanning n_pages_total pages like
    `pages` (u8 [k, h, w], the first two are sampled) with `tpls`: 2 * n_w * n_h per (template, window), unpadded box
    (SURVEY 8d)."""
    r_h, r_w = pages.shape[1:]
    classes = class_table(tpls)
    frac = {(w, h): float(np.mean([effective_window_fraction(pages[i], w, h) for i in range(min(2, len(pages)))]))
            for w, h, _ in classes}
    dense = sum(2.0 * w * h * c * dense_windows(w, h, r_w, r_h) for w, h, c in classes) * n_pages_total
    eff = sum(2.0 * w * h * c * dense_windows(w, h, r_w, r_h) * frac[(w, h)] for w, h, c in classes) * n_pages_total
    return eff, dense, float(np.mean(list(frac.values())))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []   # lines: (time received, text)
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.device)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)   # the sampler's last line may still be in flight
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        # only the samples taken between mark_begin() and mark_end(): the timed legs (nvidia-smi is started long before,
        # it needs up to a second to deliver its first line)
        t0 = self.t0 if self.t0 is not None else 0.0
        t1 = (self.t1 if self.t1 is not None else time.time()) + 0.06
        for ts, ln in self.lines:
            if ts < t0 or ts > t1:
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def dist_setup():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


# ------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference_sample(pkg, bank, cores, stride, seed0=9000):
    """The reference's own kernel on the host cores: one worker per core, one page per worker
    (the reference's rayon loop, ncc.rs:839-846), every `stride`-th template; returns the time one
    full page would take per worker (SAT/stat preparation + all T templates), extrapolated linearly."""
    from oracle import oracle as O

    impl = "reference" if O.ref_lib() is not None else "port"
    tpls = [t.pixels for t in bank.templates]
    sample = list(range(0, len(tpls), stride))
    pages = make_pages(pkg, bank, cores, seed0)
    out = [None] * cores

    def work(i):
        t0 = time.perf_counter()
        s = O.Searcher(pages[i], impl)
        sizes = sorted({t.shape for t in tpls})
        t_prep = time.perf_counter() - t0
        t_stats = t_scan = 0.0
        hits = 0
        for sz in sizes:  # the reference re-runs prepare_for_size whenever the box size changes
            ts = time.perf_counter()
            s.prepare_for_size(sz[1], sz[0])
            t_stats += time.perf_counter() - ts
            for k in sample:
                if tpls[k].shape != sz:
                    continue
                ts = time.perf_counter()
                hits += len(s.search_c_u8(tpls[k], THRESHOLD))
                t_scan += time.perf_counter() - ts
        out[i] = (t_prep + t_stats, t_scan, hits)

    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(i,)) for i in range(cores)]
    [t.start() for t in th]
    [t.join() for t in th]
    wall = time.perf_counter() - t0
    scale = len(tpls) / float(len(sample))
    full_page_s = max(p + s * scale for p, s, _ in out)  # slowest worker bounds the batch
    return {"impl": impl, "cores": cores, "wall_s": wall, "full_page_s": full_page_s,
            "pages_per_s": cores / full_page_s, "sample_templates": len(sample), "templates": len(tpls),
            "kernel_ns_per_px_per_template": 1e9 * float(np.mean([s for _, s, _ in out])) / (len(sample) * R_W * R_H)}


def run_reference(args):
    rank, world, local = dist_setup()
    if rank != 0:
        return 0
    pkg, font, bank = make_bank()
    cores = os.cpu_count() or 1
    stride = args.cpu_stride
    times = []
    for i in range(args.warmup + args.steps):
        r = cpu_reference_sample(pkg, bank, cores, stride, seed0=9000 + 100 * i)
        if i >= args.warmup:
            times.append(r)
    full = float(np.mean([r["full_page_s"] for r in times]))
    value = cores / full
    classes = class_table([t.pixels for t in bank.templates])
    evals_page = sum(c * dense_windows(w, h) for w, h, c in classes)
    sample = (f"{cores} pages (one per core) x every {stride}th template ({times[0]['sample_templates']} of "
              f"{times[0]['templates']}), SAT+stats once per page, extrapolated linearly to all templates")
    line = {
        "impl": "reference", "metric": "pages/sec", "value": value, "unit": "pages/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean([r["wall_s"] for r in times])),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "pages_per_gpu_per_step": cores, "l2": "n/a (CPU)"},
        "evals_per_sec": value * evals_page,
        "cpu_baseline": {"value": value, "unit": "pages/s", "cores": cores, "kind": "reference" if times[0]["impl"] == "reference" else "port",
                         "sample": sample, "kernel_ns_per_px_per_template": times[0]["kernel_ns_per_px_per_template"]},
        "e2e": {"value": value, "unit": "pages/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------ helpers of our arm
def same_matches(a_m, a_c, b_m, b_c):
    """Two scan results are byte-identical over their valid entries (entries beyond the count are unspecified)."""
    if not np.array_equal(a_c, b_c):
        return False
    valid = np.arange(a_m.shape[-1])[None, None, :] < a_c[:, :, None]
    av, bv = a_m.view(np.uint64), b_m.view(np.uint64)
    return bool(np.array_equal(av[valid], bv[valid]))


def oracle_spot_check(pages, tpls, m, c, picks):
    """(page, template) match lists against the compiled reference kernel (oracle/_ref; the C port when absent) --
    outside any timed region.  Returns a description; raises on a mismatch."""
    from oracle import oracle as O

    impl = "reference" if O.ref_lib() is not None else "port"
    wide = any(t.shape[1] > 16 for t in tpls)
    if wide:
        impl = "port"
    for p in sorted({p for p, _ in picks}):
        s = O.Searcher(pages[p], impl)
        for t in [t for pp, t in picks if pp == p]:
            exp = s.search_c_u8(tpls[t], THRESHOLD, allow_wide=wide)
            n = int(c[p, t])
            if n != len(exp) or m[p, t, :n].tobytes() != exp.tobytes():
                raise AssertionError(f"page {p} template {t}: GPU match list differs from the oracle ({impl})")
    return {"checked": [list(x) for x in picks], "against": "oracle/_ref (compiled reference kernel)" if impl == "reference" else "oracle C port",
            "identical": True}


def single_page_config(ctx, native, ncc, pkg, font, stream, torch, name, x_bits, y_bits, int8_peak, with_cpu):
    """BASELINE config 1 / 2: ONE 608x800 page at -t 13: latency device-resident and end to end (pinned host buffers),
    match lists checked against the oracle, roofline of the whole call, the reference kernel on one host core beside it
    (the reference parallelises over pages, ncc.rs:839-846: one page = one core)."""
    bank_h = pkg.raster.TemplateBank(font, 13, x_bits=x_bits, y_bits=y_bits)
    tpls = [t.pixels for t in bank_h.templates]
    T = len(tpls)
    page = pkg.pages.make_ncc_page(bank_h, 608, 800, seed=0, shifts="bank")[0]
    bank = ncc.Bank(ctx, tpls)
    pin = torch.from_numpy(page[None]).pin_memory()
    dev = pin.cuda()
    out_dev = torch.empty(T * N_OUT * 8, dtype=torch.uint8, device="cuda")
    cnt_dev = torch.empty(T, dtype=torch.int32, device="cuda")
    out_pin = torch.empty(T * N_OUT * 8, dtype=torch.uint8).pin_memory()
    cnt_pin = torch.empty(T, dtype=torch.int32).pin_memory()
    m = out_pin.numpy().view(native.MATCH_DTYPE).reshape(1, T, N_OUT)
    c = cnt_pin.numpy().view(np.uint32).reshape(1, T)

    def dev_call():
        ncc.scan_pages_device(ctx, bank, dev.data_ptr(), 608 * 800, 608, 608, 800, 1, THRESHOLD, N_OUT, out_dev.data_ptr(),
                              cnt_dev.data_ptr())

    def e2e_call():
        ncc.scan_pages(ctx, bank, pin.numpy(), THRESHOLD, N_OUT, out=m, counts=c)

    res = {}
    for key, fn in (("device", dev_call), ("e2e", e2e_call)):
        ms = []
        for i in range(13):  # the first three calls warm up (allocations)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if key == "device":
                e0.record(stream)
                fn()
                e1.record(stream)
                e1.synchronize()
                ms.append(e0.elapsed_time(e1))
            else:
                t0 = time.perf_counter()
                fn()
                ms.append(1e3 * (time.perf_counter() - t0))
        res[key] = float(np.median(ms[3:]))
    ctx.profile(True)
    ctx.profile_read()
    for _ in range(5):
        dev_call()
    prof = ctx.profile_read()
    ctx.profile(False)
    scan_ms = (prof["scan"][0] - prof["exact"][0]) / 5
    eff, dense, frac = algorithmic_ops(tpls, page[None], 1)
    check = oracle_spot_check(page[None], tpls, m, c, [(0, t) for t in range(0, T, max(T // 6, 1))])
    out = {"workload": f"{name}: one 608x800 page, -t 13 --x-bits {x_bits} --y-bits {y_bits} ({T} templates), thr 0.8",
           "templates": T, "box_sizes": [list(x) for x in class_table(tpls)],
           "latency_ms_device": res["device"], "latency_ms_e2e": res["e2e"],
           "pages_per_s_device": 1e3 / res["device"], "pages_per_s_e2e": 1e3 / res["e2e"],
           "timing": "device: CUDA events around focr_ncc_scan_device (page in HBM); e2e: host wall clock around focr_ncc_scan "
                     "(pinned host page -> H2D -> kernels -> D2H of the match lists); median of 10 calls after 3 warm-up calls",
           "roofline": {"bound": "tensor", "unit": "TFLOP/s", "achieved": eff / (scan_ms / 1e3) / 1e12 if scan_ms > 0 else None,
                        "peak": int8_peak, "frac": (eff / (scan_ms / 1e3) / 1e12 / int8_peak) if (int8_peak and scan_ms > 0) else None,
                        "kernel": "scan_tc_kernel", "kernel_ms_per_page": scan_ms, "whole_call_achieved": eff / (res["device"] / 1e3) / 1e12,
                        "w_eff_over_w_dense": frac, "stage_ms": {k: v[0] / 5 for k, v in prof.items() if k != "decode"}},
           "oracle_check": check, "hits": int(c.sum())}
    if with_cpu:
        from oracle import oracle as O

        impl = "reference" if O.ref_lib() is not None else "port"
        t0 = time.perf_counter()
        O.get_hits(page, tpls, THRESHOLD, impl)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": 1.0 / dt, "unit": "pages/s", "cores": 1, "kind": impl if impl == "reference" else "port",
                               "sample": "the same page, all templates, SAT + statistics + kernel (Searcher::new + search_c_u8 per template); "
                                         "one core: the reference parallelises over pages only", "latency_ms": 1e3 * dt}
    bank.close()
    return out


def focr_config4(ctx, native, pkg, font, torch, peaks, with_cpu):
    """BASELINE config 4: focr's least-squared-distance line decode (main.rs:87-218, -x 45 -y 39 -w 608 --line-height 12
    --line-advance 15) on 2480x3508 pages through the host-buffer C ABI with the cached glyph rasters."""
    from font_ocr_b200 import focr

    fbank = focr.GlyphBank(ctx, font, TEXT_SIZE)
    fdistinct = [pkg.pages.make_focr_page(font, TEXT_SIZE, R_W, R_H, seed=7000 + i)[0] for i in range(4)]
    fP = 32
    fpages_pin = torch.from_numpy(np.stack([fdistinct[i % 4] for i in range(fP)])).pin_memory()
    fpages = fpages_pin.numpy()   # pinned host pages, like the headline's e2e leg
    fmax = (R_H - 39 + 14) // 15
    fg = np.zeros((fP, fmax, 512), np.uint16)
    fn_, fy, fl = np.zeros((fP, fmax), np.uint32), np.zeros((fP, fmax), np.uint32), np.zeros(fP, np.uint32)

    def call(pages):
        native.check(native.lib().focr_decode_pages(ctx._h, fbank._h, native.ptr(pages), R_W * R_H, R_W, R_H, len(pages), 45, 39, 608,
                                                    12, 15, fmax, 512, native.ptr(fg), native.ptr(fn_), native.ptr(fy), native.ptr(fl)))

    ts = []
    for _ in range(6):  # first call = warm-up
        t0 = time.perf_counter()
        call(fpages)
        ts.append(time.perf_counter() - t0)
    dt = float(np.median(ts[1:]))
    pageable = np.array(fpages)   # an ordinary (pageable) copy of the batch
    tp = []
    for _ in range(4):
        t0 = time.perf_counter()
        call(pageable)
        tp.append(time.perf_counter() - t0)
    ctx.profile(True)
    ctx.profile_read()
    for _ in range(3):
        call(fpages)
    prof = ctx.profile_read()
    ctx.profile(False)
    k_ms = prof["decode"][0] / max(prof["decode"][1], 1)
    lines, cells = float(fl.mean()), float(fn_.sum())
    # algorithmic bytes: lines * w * line_height strip bytes + the raster bank, once per call
    bank_bytes, cache = 0, None
    try:
        from oracle import oracle as O

        cache = O.GlyphCache(font, pkg.raster.FOCR_DEFAULT_ALPHABET, TEXT_SIZE)
        bank_bytes = int((cache.rasters["w"].astype(np.int64) * cache.rasters["h"]).sum())
    except Exception:
        pass
    alg_bytes = float(fl.sum()) * 608 * 12 + bank_bytes
    hbm = peaks.get("hbm_gbs")
    achieved = alg_bytes / (k_ms / 1e3) / 1e9 if k_ms > 0 else None
    out = {"workload": "config4: focr line decode, 32 pages 2480x3508, -x 45 -y 39 -w 608 --line-height 12 --line-advance 15, "
                       "67 glyphs x 64 phases cached",
           "pages_per_s": fP / dt, "pages_per_s_pageable": fP / float(np.median(tp[1:])), "lines_per_page": lines, "cells_per_s": cells / dt,
           "timing": "host wall clock around focr_decode_pages (host pages -> H2D of the line band -> kernel -> D2H of the glyph "
                     "indices), median of 5 calls; pinned and pageable host pages",
           "h2d_bytes_per_call": int(fP * 608 * (R_H - 39)), "kernel_ms_per_call": k_ms,
           "roofline": {"bound": "hbm", "unit": "GB/s", "achieved": achieved, "peak": hbm, "frac": (achieved / hbm) if (achieved and hbm) else None,
                        "traffic": None, "kernel": "focr_decode_kernel",
                        "note": "latency-bound, not bandwidth-bound: the pen walk of a line is sequential (the next cell's position "
                                "depends on the chosen glyph, main.rs:176-178), one warp per line; profiles/ holds the ncu evidence",
                        "cells_per_s_kernel": cells / (k_ms / 1e3) if k_ms > 0 else None}}
    if with_cpu and cache is not None:
        from oracle import oracle as O

        cores = os.cpu_count() or 1
        res = [None] * cores

        def work(i):
            t0 = time.perf_counter()
            r = O.decode_image_cached(fdistinct[i % 4][:39 + 15 * 24 + 12], cache, 45, 39, 608, 12, 15)   # the first 24 lines of a page
            res[i] = (time.perf_counter() - t0, len(r))

        th = [threading.Thread(target=work, args=(i,)) for i in range(cores)]
        t0 = time.perf_counter()
        [t.start() for t in th]
        [t.join() for t in th]
        per_line = max(r[0] / max(r[1], 1) for r in res)
        out["cpu_baseline"] = {"value": cores / (per_line * lines), "unit": "pages/s", "cores": cores, "kind": "port",
                               "sample": f"decode_image with CACHED glyph rasters (oracle C restatement: whole-canvas sum_of_squares per glyph per "
                                         f"cell, main.rs:87-110), {cores} threads x the first 24 lines of a page, extrapolated to {lines:.0f} lines per page"}
        t0 = time.perf_counter()
        r = O.decode_image(fdistinct[0][:39 + 15 * 3 + 12], font, pkg.raster.FOCR_DEFAULT_ALPHABET, TEXT_SIZE, 45, 39, 608, 12, 15)
        dt1 = (time.perf_counter() - t0) / max(len(r), 1)
        out["cpu_baseline_per_cell_rasterisation"] = {
            "value": 1.0 / (dt1 * lines), "unit": "pages/s", "cores": 1, "kind": "port",
            "sample": "decode_image with one FreeType rasterisation per glyph per cell like the reference (main.rs:98-106), through the Python "
                      f"restatement (ctypes FreeType: interpreter overhead included), 3 lines on one core, extrapolated to {lines:.0f} lines per page"}
    fbank.close()
    return out


def config5(ctx, ncc, pkg, font, stream, torch, int8_peak):
    """BASELINE config 5: 95 printable-ASCII glyphs at -t 24, --x-bits 3 --y-bits 2 (3040 templates of about 27x26 pixels, wider
    than the reference's AVX2 kernel accepts) on 2480x3508 pages, device-resident."""
    alphabet5 = "".join(chr(ch) for ch in range(32, 127))
    bank5_h = pkg.raster.TemplateBank(font, 24, x_bits=3, y_bits=2, alphabet=alphabet5)
    tpls5 = [t.pixels for t in bank5_h.templates]
    T5, P5 = len(tpls5), 4
    bank5 = ncc.Bank(ctx, tpls5)
    pages5_np = np.stack([pkg.pages.make_ncc_page(bank5_h, R_W, R_H, seed=500 + i, shifts="bank")[0] for i in range(P5)])
    pages5 = torch.from_numpy(pages5_np).cuda()
    out5 = torch.empty(P5 * T5 * N_OUT * 8, dtype=torch.uint8, device="cuda")
    cnt5 = torch.empty(P5 * T5, dtype=torch.int32, device="cuda")

    def call():
        ncc.scan_pages_device(ctx, bank5, pages5.data_ptr(), R_W * R_H, R_W, R_W, R_H, P5, THRESHOLD, N_OUT, out5.data_ptr(), cnt5.data_ptr())

    ms5 = []
    for _ in range(3):  # first = warm-up
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        call()
        e1.record(stream)
        e1.synchronize()
        ms5.append(e0.elapsed_time(e1))
    ctx.profile(True)
    ctx.profile_read()
    call()
    prof = ctx.profile_read()
    ctx.profile(False)
    scan_ms, scan_l = prof["scan"][0] - prof["exact"][0], prof["scan"][1] - prof["exact"][1]
    eff, dense, frac = algorithmic_ops(tpls5, pages5_np, P5)
    t5 = float(np.median(ms5[1:])) / 1e3
    achieved = eff / (scan_ms / 1e3) / 1e12 if scan_ms > 0 else None
    out = {"workload": "config5: 95 glyphs, -t 24 --x-bits 3 --y-bits 2, 2480x3508 pages", "templates": T5,
           "box_sizes": [list(x) for x in class_table(tpls5)], "pages": P5, "pages_per_s": P5 / t5,
           "dense_tops_whole_pipeline": dense / t5 / 1e12, "hits": int(cnt5.sum().item()),
           "roofline": {"bound": "tensor", "unit": "TFLOP/s", "achieved": achieved, "peak": int8_peak,
                        "frac": (achieved / int8_peak) if (achieved and int8_peak) else None, "traffic": None, "kernel": "scan_tc_kernel",
                        "avg_launch_ms": scan_ms / max(scan_l, 1), "launches": scan_l, "w_eff_over_w_dense": frac,
                        "achieved_dense_windows": dense / (scan_ms / 1e3) / 1e12 if scan_ms > 0 else None,
                        "ops": "u8 x u8 -> s32 multiply-adds x 2, unpadded n_w*n_h, W_eff windows; time = correlation kernel launches "
                               "(CUDA events, focr_ctx_profile)",
                        "stage_ms": {k: v[0] for k, v in prof.items() if k != "decode"}},
           "cpu_baseline": None, "cpu_baseline_note": "the reference's AVX2 kernel panics for boxes wider than 16 (ncc.rs:392); the C port is "
                                                      "the parity oracle for this config (tests/test_gpu_ncc.py::test_config5_full_page_sample)"}
    m5 = out5.cpu().numpy().view(np.dtype([("x", np.uint16), ("y", np.uint16), ("similarity", np.float32)])).reshape(P5, T5, N_OUT)
    c5 = cnt5.cpu().numpy().view(np.uint32).reshape(P5, T5)
    out["oracle_check"] = oracle_spot_check(pages5_np, tpls5, m5, c5, [(1, 1234)])
    bank5.close()
    del pages5, out5, cnt5
    return out


def bind_to_gpu_numa_node(local):
    """Multi-rank runs: keep this rank's host threads (and with them its first-touched pinned buffers and the library's
    staging threads) on the NUMA node its GPU hangs off -- N ranks streaming pages through one host memory system otherwise
    cross the socket interconnect at random.  Returns what was done, for the bench line.  BENCH_NO_NUMA=1 skips it."""
    if os.environ.get("BENCH_NO_NUMA"):
        return {"bound": False, "why": "BENCH_NO_NUMA"}
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(local)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:   # nvml prints an 8-digit PCI domain, sysfs a 4-digit one
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return {"bound": False, "why": "the GPU reports no NUMA node"}
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return {"bound": False, "why": "no allowed CPU on the GPU's node"}
        os.sched_setaffinity(0, cpus)
        return {"bound": True, "node": node, "cpus": len(cpus)}
    except Exception as e:   # noqa: BLE001 -- a missing sysfs entry must not fail the bench
        return {"bound": False, "why": f"{type(e).__name__}: {e}"}


# ------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    rank, world, local = dist_setup()
    numa = bind_to_gpu_numa_node(local) if world > 1 else {"bound": False, "why": "single rank"}
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        cpu_group = dist.new_group(backend="gloo")   # host-side waits that must not occupy the GPUs
    torch.cuda.set_device(local)
    sampler = ClockSampler(local)   # started now (nvidia-smi takes a while to deliver); only the timed legs' samples are used
    sampler.start()
    from font_ocr_b200 import native, ncc

    pkg, font, bank_h = make_bank()
    tpls = [t.pixels for t in bank_h.templates]
    T = len(tpls)
    P = args.pages
    ctx = ncc.Context(local)
    if args.kernel != "auto":
        ctx.set_kernel(native.KERNEL_SIMT if args.kernel == "simt" else native.KERNEL_TCGEN05)
    bank = ncc.Bank(ctx, tpls)
    pages_np = make_pages(pkg, bank_h, P, seed0=1000 * rank, distinct=args.distinct)
    pages_pin = torch.from_numpy(pages_np).pin_memory()
    pages_dev = pages_pin.cuda(non_blocking=False)
    out_dev = torch.empty(P * T * N_OUT * 8, dtype=torch.uint8, device="cuda")
    counts_dev = torch.empty(P * T, dtype=torch.int32, device="cuda")
    out_pin = torch.empty(P * T * N_OUT * 8, dtype=torch.uint8).pin_memory()
    counts_pin = torch.empty(P * T, dtype=torch.int32).pin_memory()
    out_np = out_pin.numpy().view(native.MATCH_DTYPE).reshape(P, T, N_OUT)
    counts_np = counts_pin.numpy().view(np.uint32).reshape(P, T)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local))

    def step_device():
        ncc.scan_pages_device(ctx, bank, pages_dev.data_ptr(), R_W * R_H, R_W, R_W, R_H, P, THRESHOLD, N_OUT,
                              out_dev.data_ptr(), counts_dev.data_ptr())

    def step_e2e():
        ncc.scan_pages(ctx, bank, pages_pin.numpy(), THRESHOLD, N_OUT, out=out_np, counts=counts_np)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, profile=False):
        for _ in range(warmup):
            fn()
        ctx.sync()
        if profile:
            ctx.profile(True)
            ctx.profile_read()
        l0 = ctx.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        e1.synchronize()
        barrier()
        ms = e0.elapsed_time(e1)
        prof = ctx.profile_read() if profile else None
        if profile:
            ctx.profile(False)
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), ctx.launch_count - l0, prof

    for _ in range(args.warmup):   # warm up before the clock samples are marked (timed() warms up again: cheap)
        step_device()
    ctx.sync()
    sampler.mark_begin()
    ms_dev, launches, prof = timed(step_device, args.steps, args.warmup, profile=True)
    ms_e2e, _, _ = timed(step_e2e, args.steps, max(args.warmup, 1))
    sampler.mark_end()
    clocks = sampler.stop()
    clocks["window"] = "the device-resident and end-to-end timed legs (warm-up steps of those legs included), nvidia-smi -lms 50"
    # the device-resident and the host-buffer paths must agree byte for byte (this also reads the results back)
    chk_counts = counts_dev.cpu().numpy().view(np.uint32).reshape(P, T)
    chk_m = out_dev.cpu().numpy().view(native.MATCH_DTYPE).reshape(P, T, N_OUT)
    assert same_matches(chk_m, chk_counts, out_np, counts_np), "device-resident and e2e paths disagree"
    del chk_m
    # a caller with PAGEABLE buffers (a Rust Vec<u8>, ncc.rs:575): the library stages through its own pinned buffers
    pageable_pages = np.array(pages_np)
    pg_out = np.zeros((P, T, N_OUT), native.MATCH_DTYPE)
    pg_counts = np.zeros((P, T), np.uint32)

    def step_pageable():
        ncc.scan_pages(ctx, bank, pageable_pages, THRESHOLD, N_OUT, out=pg_out, counts=pg_counts)

    ms_pg, _, _ = timed(step_pageable, min(args.steps, 3), 1)
    assert same_matches(pg_out, pg_counts, out_np, counts_np), "pageable and pinned e2e paths disagree"
    # ... and the same ordinary buffers page-locked IN PLACE through the library (focr_pin_register: what a Rust caller does
    # with its Vec<u8> once): direct DMA again, registration outside the timed region
    pg_out[:] = 0
    pg_counts[:] = 0
    pins = [ctx.pin(a) for a in (pageable_pages, pg_out, pg_counts)]
    ms_reg, _, _ = timed(step_pageable, min(args.steps, 3), 1)
    for pn in pins:
        pn.release()
    assert same_matches(pg_out, pg_counts, out_np, counts_np), "registered and pinned e2e paths disagree"
    del pageable_pages, pg_out

    # ONE batch sharded by page over all N GPUs through the library's multi-device entry (strong scaling); rank 0 drives all
    # devices from one process, the other ranks wait on the host (gloo) so that their GPUs are idle
    strong = None
    if world > 1:
        barrier()
        if rank == 0:
            try:
                mctx = ncc.MultiContext(devices=list(range(world)))
                mbank = ncc.MultiBank(mctx, tpls)
                s_out = np.zeros((P, T, N_OUT), native.MATCH_DTYPE)
                s_out_pin = torch.from_numpy(s_out.view(np.uint8).reshape(-1)).pin_memory()
                s_m = s_out_pin.numpy().view(native.MATCH_DTYPE).reshape(P, T, N_OUT)
                s_counts_pin = torch.zeros(P * T, dtype=torch.int32).pin_memory()
                s_c = s_counts_pin.numpy().view(np.uint32).reshape(P, T)
                ts = []
                for i in range(2 + max(args.steps, 3)):
                    mctx.sync()
                    t0 = time.perf_counter()
                    ncc.scan_pages_multi(mctx, mbank, pages_pin.numpy(), THRESHOLD, N_OUT, out=s_m, counts=s_c)
                    ts.append(time.perf_counter() - t0)
                same = same_matches(s_m, s_c, out_np, counts_np)
                t_med = float(np.median(ts[2:]))
                strong = {"scaling": "strong", "n_gpus": world, "pages": P, "value": P / t_med, "unit": "pages/s",
                          "ms_per_batch": 1e3 * t_med, "speedup_vs_1gpu_e2e": (ms_e2e / args.steps / 1e3) / t_med,
                          "identical_to_1gpu": bool(same),
                          "entry": "focr_multi_ncc_scan: one process, one context + host thread per GPU, contiguous page blocks, "
                                   "results gathered in place by page index, no collective",
                          "timing": "host wall clock around the call (pinned host pages -> per-GPU H2D, kernels, D2H), median of "
                                    f"{len(ts) - 2} after 2 warm-up calls",
                          "limiter": "fixed per-call costs that do not shrink with the block: the 2-4-8-16 page chunk ramp (pipeline fill), "
                                     "the drain of the last chunk's D2H, and N x PCIe traffic through one host memory system"}
                assert same, "multi-GPU result differs from the 1-GPU result"
                mbank.close()
                mctx.close()
            except Exception as ex:
                strong = {"error": str(ex)[:300]}
        dist.barrier(group=cpu_group)

    # SURVEY 8f rank 1 (reported beside the headline, not part of it): process_hits on the device, on the match lists
    # the scan left in HBM, timed with CUDA events on the library's stream including the D2H of the surviving lines;
    # next to it the C++ host mirror of the same function on the first page's hits
    post = None
    if rank == 0:
        try:
            letters = bank_h.letters()
            lp, ls, st_, sm_ = ncc.process_hits_device(ctx, out_dev.data_ptr(), counts_dev.data_ptr(), T, N_OUT, P, letters,
                                                       0.95, 5, raw=True)
            # timed: the raw C call with pinned result buffers sized from the first call (what a host integration does)
            import ctypes as C
            line_cap, sel_cap = len(lp) + 64, len(st_) + 4096
            pin = lambda n, dt: torch.empty(n, dtype=dt).pin_memory()
            b_lp, b_ls, b_st, b_sm = pin(line_cap, torch.int32), pin(line_cap + 1, torch.int32), pin(sel_cap, torch.int32), \
                pin(sel_cap * 2, torch.int32)
            nl_, ns_ = np.zeros(1, np.uint32), np.zeros(1, np.uint32)
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                native.check(native.lib().focr_process_hits_device(
                    ctx._h, native.ptr(out_dev.data_ptr()), native.ptr(counts_dev.data_ptr()), T, N_OUT, P, C.c_float(0.95), 5,
                    line_cap, sel_cap, native.ptr(nl_), native.ptr(ns_), native.ptr(b_lp.data_ptr()), native.ptr(b_ls.data_ptr()),
                    native.ptr(b_st.data_ptr()), native.ptr(b_sm.data_ptr())))
                e1.record(stream)
                e1.synchronize()
            assert int(nl_[0]) == len(lp) and int(ns_[0]) == len(st_)
            assert np.array_equal(b_st.numpy()[:len(st_)].view(np.uint32), st_)
            lines = ncc.process_hits_device(ctx, out_dev.data_ptr(), counts_dev.data_ptr(), T, N_OUT, P, letters, 0.95, 5,
                                            only_page=0)
            t0 = time.perf_counter()
            host_lines = ncc.host_process_hits(ncc.get_hits(out_np[0], counts_np[0], letters), 0.95, 5)
            t_host = time.perf_counter() - t0
            assert ncc.lines_to_text(host_lines) == ncc.lines_to_text(lines[0])
            post = {"kernel": "focr_process_hits_device (ncc.rs:723-786)", "ms_per_page": e0.elapsed_time(e1) / P,
                    "note": "stream time around the C call per page (6 small kernels + CUB radix sort / scans + D2H of the surviving hits into pinned buffers)",
                    "lines_per_page": len(lp) / float(P), "survivors_per_page": len(st_) / float(P),
                    "host_mirror_ms_per_page": 1e3 * t_host, "host_mirror": "C++ process_hits via ctypes incl. marshalling, 1 page"}
        except Exception as ex:  # never let the extra measurement break the bench line
            post = {"error": str(ex)[:200]}

    total_pages = P * world
    value = total_pages * args.steps / (ms_dev / 1e3)
    e2e_value = total_pages * args.steps / (ms_e2e / 1e3)
    classes = class_table(tpls)
    evals_page = sum(c * dense_windows(w, h) for w, h, c in classes)
    peaks = load_peaks()

    line = None
    if rank == 0:
        # the measured integer tensor peak: our own tcgen05.mma kind::i8 micro-benchmark, run live on this GPU
        # (M128 N256 K32 back to back on every SM; tools/microbench, not part of the product library)
        int8_peak, int8_src = None, None
        try:
            from tools.microbench import microbench

            int8_peak = microbench.int8_peak_tops(local, sm_count=SM_COUNT)
            int8_src = "measured live: tcgen05.mma kind::i8 M128 N256 K32 back to back on all SMs (tools/microbench); MEASURED_PEAKS.json has no integer peak"
        except Exception as ex:
            int8_src = f"micro-benchmark unavailable ({str(ex)[:80]})"
        bf16 = peaks.get("bf16_tflops_sustained")
        peak2 = 2.0 * bf16 if bf16 else 2800.0
        if int8_peak is None:   # fallback, flagged: int8 dense = 2 x bf16
            int8_peak = peak2
            int8_src += "; FALLBACK 2 x bf16_tflops_sustained (MEASURED_PEAKS.json)" if bf16 else "; FALLBACK 2 x 1.4 PFLOP/s (B200_PROFILING.md)"

        # roofline of the dominant kernel (the correlation scan): algorithmic integer ops / measured time
        ops_eff, ops_dense, w_frac = algorithmic_ops(tpls, pages_np, P * args.steps)
        # "scan" = correlation kernel + exact pass over its survivors; the roofline is for the correlation kernel alone
        scan_ms, scan_launches = prof["scan"]
        exact_ms, exact_launches = prof.get("exact", (0.0, 0))
        scan_ms -= exact_ms
        scan_launches -= exact_launches
        achieved = ops_eff / (scan_ms / 1e3) / 1e12 if scan_ms > 0 else 0.0
        traffic = None
        try:  # per-launch DRAM bytes of the correlation kernel from the committed ncu capture
            tj = json.load(open(os.path.join(ROOT, "profiles", "r2_scan_tc_traffic.json")))
            traffic = tj["dram_bytes_per_page"] * (P * args.steps) / max(scan_launches, 1)
        except Exception:
            pass
        roofline = {"bound": "tensor", "achieved": achieved, "peak": int8_peak, "unit": "TFLOP/s", "frac": achieved / int8_peak,
                    "traffic": traffic, "kernel": "scan_tc_kernel (" + args.kernel + ")", "peak_source": int8_src,
                    "peak_2x_bf16_sustained": peak2, "frac_of_2x_bf16_sustained": achieved / peak2,
                    "ops": "u8 x u8 -> s32 multiply-adds x 2, unpadded n_w*n_h, W_eff windows (SURVEY 8d)",
                    "achieved_dense_windows": ops_dense / (scan_ms / 1e3) / 1e12 if scan_ms > 0 else 0.0,
                    "whole_step_achieved": ops_eff / (ms_dev / 1e3) / 1e12,
                    "avg_launch_ms": scan_ms / max(scan_launches, 1), "launches": scan_launches,
                    "w_eff_over_w_dense": w_frac,
                    "stage_ms_per_step": {k: v[0] / args.steps for k, v in prof.items() if k != "decode"}}
        oracle_check = None
        try:
            oracle_check = oracle_spot_check(pages_np, tpls, out_np, counts_np, [(0, 5), (P // 2, 150), (P - 1, T - 1)])
        except AssertionError:
            raise
        except Exception as ex:
            oracle_check = {"error": str(ex)[:200]}
        cpu = None
        if world == 1 and not args.no_cpu:
            r = cpu_reference_sample(pkg, bank_h, os.cpu_count() or 1, args.cpu_stride)
            cpu = {"value": r["pages_per_s"], "unit": "pages/s", "cores": r["cores"],
                   "kind": "reference" if r["impl"] == "reference" else "port",
                   "sample": f"{r['cores']} pages (one per core) x every {args.cpu_stride}th template "
                             f"({r['sample_templates']} of {r['templates']}), extrapolated linearly",
                   "kernel_ns_per_px_per_template": r["kernel_ns_per_px_per_template"]}

        def guarded(fn, *a):
            try:
                return fn(*a)
            except Exception as ex:   # never let a side measurement break the bench line
                return {"error": f"{type(ex).__name__}: {str(ex)[:300]}"}

        side = world == 1 or args.side_configs
        cfg1 = cfg2 = cfg5 = focr_line = None
        if side and not args.no_small:
            cfg1 = guarded(single_page_config, ctx, native, ncc, pkg, font, stream, torch, "config1", 0, 0, int8_peak, not args.no_cpu)
            cfg2 = guarded(single_page_config, ctx, native, ncc, pkg, font, stream, torch, "config2", 2, 2, int8_peak, not args.no_cpu)
        if side and not args.no_focr:
            focr_line = guarded(focr_config4, ctx, native, pkg, font, torch, peaks, not args.no_cpu)
        if side and not args.no_config5:
            cfg5 = guarded(config5, ctx, ncc, pkg, font, stream, torch, int8_peak)

        line = {
            "metric": "pages/sec", "value": value, "unit": "pages/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "pages_per_gpu_per_step": P, "distinct_pages": min(args.distinct or P, P),
                       "templates": T, "box_sizes": [[w, h, c] for w, h, c in classes], "kernel": args.kernel,
                       "l2": "inputs (870 MB per 100 pages) larger than L2; no flush", "numa": numa},
            "evals_per_sec": value * evals_page,
            "e2e": {"value": e2e_value, "unit": "pages/s", "h2d_bytes_per_step": int(pages_pin.numel()) * world,
                    "d2h_bytes_per_step": int(out_pin.numel() + counts_pin.numel() * 4) * world,
                    "ms_per_step": ms_e2e / args.steps, "host_buffers": "pinned (cudaHostAlloc): DMA straight from / to the caller's buffers",
                    "pageable": {"value": total_pages * min(args.steps, 3) / (ms_pg / 1e3), "unit": "pages/s",
                                 "host_buffers": "pageable (numpy): staged through the library's own pinned buffers by host threads",
                                 "identical_to_pinned": True},
                    "registered": {"value": total_pages * min(args.steps, 3) / (ms_reg / 1e3), "unit": "pages/s",
                                   "host_buffers": "the same numpy buffers page-locked in place with focr_pin_register (outside the timed region)",
                                   "identical_to_pinned": True},
                    "identical_to_device_path": True},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
            "oracle_check": oracle_check, "strong": strong,
            "process_hits_device": post, "config1": cfg1, "config2": cfg2, "focr": focr_line, "config5": cfg5,
        }
    bank.close()
    ctx.close()
    if world > 1:
        dist.barrier(group=cpu_group)
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pages", type=int, default=100, help="pages per GPU per step")
    ap.add_argument("--distinct", type=int, default=None, help="distinct synthetic pages to generate (default: all)")
    ap.add_argument("--kernel", default="auto", choices=["auto", "simt", "tcgen05"])
    ap.add_argument("--cpu-stride", type=int, default=8, help="CPU baseline scans every k-th template")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-small", action="store_true", help="skip the config-1 / config-2 single-page measurements")
    ap.add_argument("--no-focr", action="store_true", help="skip the config-4 focr line-decode measurement")
    ap.add_argument("--no-config5", action="store_true", help="skip the config-5 (large template bank) measurement")
    ap.add_argument("--side-configs", action="store_true", help="run configs 1, 2, 4, 5 also when N > 1 (default: N = 1 only)")
    args = ap.parse_args()
    if args.impl == "reference":
        # the reference arm's step is a bounded CPU sample: keep the default run within minutes
        if "--steps" not in " ".join(sys.argv):
            args.steps = 3
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
