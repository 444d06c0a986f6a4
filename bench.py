#!/usr/bin/env python
"""bench.py -- font-ocr NCC template scan on B200: pages/sec (and template-window evals/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--pages P]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2], the configuration the pages/sec metric is quoted on): a batch of
synthetic 2480x3508 (300 dpi) pages of base64 text, `-t 13 --x-bits 2` -> 4 subpixel offsets x 74
letters = 296 templates, threshold 0.8, 1024 matches per (page, template).  One STEP = one pass of
the hot path over one batch of P pages per GPU (default P = 100; the batch, 870 MB, is far larger
than the 126 MB L2, so no flush is needed between steps).  Pages shard across ranks with no
data-path collective (weak scaling: every rank scans its own P pages).

Printed JSON (one line, rank 0): `value` = pages/s with the pages already resident in HBM, timed
with CUDA events on the library's stream; `e2e` = the same through the public host-buffer C-ABI call
(pinned host pages -> H2D -> kernels -> D2H of the match lists inside the timed region);
`roofline` = the correlation kernel against the integer tensor pipe; `cpu_baseline` = the
reference's own AVX2 kernel (oracle/_ref) on the box's host cores, on a bounded sample.
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


def class_table(bank):
    """[(n_w, n_h, n_templates)] per box size."""
    d = {}
    for t in bank.templates:
        k = t.pixels.shape[::-1]
        d[k] = d.get(k, 0) + 1
    return [(w, h, c) for (w, h), c in sorted(d.items())]


def dense_windows(n_w, n_h):
    return (R_W - n_w) * (R_H - n_h)  # x in [1, r_w-n_w], y in [1, r_h-n_h]


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


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.device)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
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


def dist_setup(n_gpus):
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    return rank, world, local


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
    rank, world, local = dist_setup(args.gpus)
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
    classes = class_table(bank)
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


# ------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    rank, world, local = dist_setup(args.gpus)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
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

    sampler = ClockSampler(local)
    sampler.start()
    ms_dev, launches, prof = timed(step_device, args.steps, args.warmup, profile=True)
    clocks = sampler.stop()
    ms_e2e, _, _ = timed(step_e2e, args.steps, max(args.warmup, 1))
    # the device-resident and the host-buffer paths must agree (and this reads the results back)
    chk_counts = counts_dev.cpu().numpy().view(np.uint32).reshape(P, T)
    assert np.array_equal(chk_counts, counts_np), "device-resident and e2e paths disagree"

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

    # BASELINE config 4 beside the headline: focr's least-squared-distance line decode (main.rs:87-218,
    # -x 45 -y 39 -w 608 --line-height 12 --line-advance 15) on 2480x3508 pages through the host-buffer C ABI
    # (H2D of the pages and D2H of the glyph indices inside the timed region), with the cached glyph rasters
    focr_line = None
    if rank == 0 and not args.no_focr:
        try:
            from font_ocr_b200 import focr
            fbank = focr.GlyphBank(ctx, font, TEXT_SIZE)
            fdistinct = [pkg.pages.make_focr_page(font, TEXT_SIZE, R_W, R_H, seed=7000 + i)[0] for i in range(4)]
            fpages_pin = torch.from_numpy(np.stack([fdistinct[i % 4] for i in range(32)])).pin_memory()
            fpages = fpages_pin.numpy()   # pinned host pages, like the headline's e2e leg
            fP, fmax = len(fpages), (R_H - 39 + 14) // 15
            fg = np.zeros((fP, fmax, 512), np.uint16)
            fn_, fy, fl = np.zeros((fP, fmax), np.uint32), np.zeros((fP, fmax), np.uint32), np.zeros(fP, np.uint32)
            ts = []
            for _ in range(4):  # first call = warm-up
                t0 = time.perf_counter()
                native.check(native.lib().focr_decode_pages(ctx._h, fbank._h, native.ptr(fpages), R_W * R_H, R_W, R_H, fP, 45, 39, 608,
                                                            12, 15, fmax, 512, native.ptr(fg), native.ptr(fn_), native.ptr(fy),
                                                            native.ptr(fl)))
                ts.append(time.perf_counter() - t0)
            dt = float(np.median(ts[1:]))
            focr_line = {"workload": "config4: focr line decode, 32 pages 2480x3508, -x 45 -y 39 -w 608 --line-height 12 --line-advance 15, "
                                     "67 glyphs x 64 phases cached",
                         "pages_per_s": fP / dt, "lines_per_page": float(fl.mean()), "cells_per_s": float(fn_.sum()) / dt,
                         "timing": "host wall clock around focr_decode_pages (pinned host pages -> H2D of the line band -> kernel -> D2H of the glyph indices), median of 3"}
            fbank.close()
        except Exception as ex:
            focr_line = {"error": str(ex)[:200]}

    # BASELINE config 5 beside the headline: 95 printable-ASCII glyphs at -t 24, --x-bits 3 --y-bits 2 (3040 templates of
    # about 27x26 pixels, wider than the reference's AVX2 kernel accepts) on 2480x3508 pages, device-resident
    config5 = None
    if rank == 0 and not args.no_config5:
        try:
            alphabet5 = "".join(chr(ch) for ch in range(32, 127))
            bank5_h = pkg.raster.TemplateBank(font, 24, x_bits=3, y_bits=2, alphabet=alphabet5)
            tpls5 = [t.pixels for t in bank5_h.templates]
            T5, P5 = len(tpls5), 4
            bank5 = ncc.Bank(ctx, tpls5)
            pages5 = torch.from_numpy(np.stack([pkg.pages.make_ncc_page(bank5_h, R_W, R_H, seed=500 + i, shifts="bank")[0]
                                                for i in range(P5)])).cuda()
            out5 = torch.empty(P5 * T5 * N_OUT * 8, dtype=torch.uint8, device="cuda")
            cnt5 = torch.empty(P5 * T5, dtype=torch.int32, device="cuda")
            ms5 = []
            for _ in range(3):  # first = warm-up
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                ncc.scan_pages_device(ctx, bank5, pages5.data_ptr(), R_W * R_H, R_W, R_W, R_H, P5, THRESHOLD, N_OUT,
                                      out5.data_ptr(), cnt5.data_ptr())
                e1.record(stream)
                e1.synchronize()
                ms5.append(e0.elapsed_time(e1))
            ops5 = sum(2.0 * t.shape[0] * t.shape[1] * (R_W - t.shape[1]) * (R_H - t.shape[0]) for t in tpls5) * P5
            t5 = float(np.median(ms5[1:])) / 1e3
            config5 = {"workload": "config5: 95 glyphs, -t 24 --x-bits 3 --y-bits 2, 2480x3508 pages", "templates": T5,
                       "box_sizes": sorted({t.shape[::-1] for t in tpls5}), "pages": P5, "pages_per_s": P5 / t5,
                       "dense_tops": ops5 / t5 / 1e12, "ops": "2 x n_w x n_h x templates x dense windows (unpadded box), whole pipeline time",
                       "hits": int(cnt5.sum().item())}
            bank5.close()
            del pages5, out5, cnt5
        except Exception as ex:
            config5 = {"error": str(ex)[:200]}

    total_pages = P * world
    value = total_pages * args.steps / (ms_dev / 1e3)
    e2e_value = total_pages * args.steps / (ms_e2e / 1e3)
    classes = class_table(bank_h)
    evals_page = sum(c * dense_windows(w, h) for w, h, c in classes)

    line = None
    if rank == 0:
        # roofline of the dominant kernel (the correlation scan): algorithmic integer ops / measured time
        frac_eff = {(w, h): float(np.mean([effective_window_fraction(pages_np[i], w, h) for i in range(min(2, P))]))
                    for w, h, _ in classes}
        ops_dense = sum(2.0 * w * h * c * dense_windows(w, h) for w, h, c in classes) * P * args.steps
        ops_eff = sum(2.0 * w * h * c * dense_windows(w, h) * frac_eff[(w, h)] for w, h, c in classes) * P * args.steps
        # "scan" = correlation kernel + exact pass over its survivors; the roofline is for the correlation kernel alone
        scan_ms, scan_launches = prof["scan"]
        exact_ms, exact_launches = prof.get("exact", (0.0, 0))
        scan_ms -= exact_ms
        scan_launches -= exact_launches
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        bf16 = peaks.get("bf16_tflops_sustained")
        peak_src = "2 x MEASURED_PEAKS.json bf16_tflops_sustained (int8 dense = 2x bf16; int8 is not in the file)"
        if bf16 is None:
            bf16, peak_src = 1400.0, "2 x fallback sustained bf16 1.4 PFLOP/s (B200_PROFILING.md; MEASURED_PEAKS.json absent)"
        peak = 2.0 * bf16
        achieved = ops_eff / (scan_ms / 1e3) / 1e12 if scan_ms > 0 else 0.0
        # our own tcgen05.mma kind::i8 micro-benchmark, run live (M128 N256 K32 back to back on every SM)
        int8_peak = None
        try:
            import ctypes as C
            cyc, ms_ = np.zeros(1), np.zeros(1)
            native.check(native.lib().focr_bench_umma_i8(ctx._h, 256, 7, 4000, 1, native.ptr(cyc), native.ptr(ms_)))
            int8_peak = 2.0 * 128 * 256 * 32 * 7 * 4000 * 148 / (float(ms_[0]) * 1e-3) / 1e12
        except Exception:
            pass
        traffic = None
        try:  # per-launch DRAM bytes of the correlation kernel from the committed ncu capture
            tj = json.load(open(os.path.join(ROOT, "profiles", "r1_scan_tc_traffic.json")))
            # per-launch DRAM bytes were captured on a 16-page chunk; scale to this run's average chunk
            if "dram_bytes_per_page" in tj:
                traffic = tj["dram_bytes_per_page"] * (P * args.steps * len(classes)) / max(scan_launches, 1)
            else:
                traffic = tj["dram_bytes_per_launch"]
        except Exception:
            pass
        roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "traffic": traffic, "kernel": "ncc scan (" + args.kernel + ")", "peak_int8_measured": int8_peak,
                    "frac_of_int8_measured": (achieved / int8_peak) if int8_peak else None, "ops": "u8 x u8 -> s32 multiply-adds x 2, unpadded n_w*n_h, W_eff windows",
                    "achieved_dense_windows": ops_dense / (scan_ms / 1e3) / 1e12 if scan_ms > 0 else 0.0,
                    "avg_launch_ms": scan_ms / max(scan_launches, 1), "launches": scan_launches,
                    "w_eff_over_w_dense": float(np.mean(list(frac_eff.values()))), "peak_source": peak_src,
                    "stage_ms_per_step": {k: v[0] / args.steps for k, v in prof.items()}}
        cpu = None
        if world == 1 and not args.no_cpu:
            r = cpu_reference_sample(pkg, bank_h, os.cpu_count() or 1, args.cpu_stride)
            cpu = {"value": r["pages_per_s"], "unit": "pages/s", "cores": r["cores"],
                   "kind": "reference" if r["impl"] == "reference" else "port",
                   "sample": f"{r['cores']} pages (one per core) x every {args.cpu_stride}th template "
                             f"({r['sample_templates']} of {r['templates']}), extrapolated linearly",
                   "kernel_ns_per_px_per_template": r["kernel_ns_per_px_per_template"]}
        line = {
            "metric": "pages/sec", "value": value, "unit": "pages/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "pages_per_gpu_per_step": P, "distinct_pages": min(args.distinct or P, P),
                       "templates": T, "box_sizes": [[w, h, c] for w, h, c in classes], "kernel": args.kernel,
                       "l2": "inputs (870 MB per 100 pages) larger than L2; no flush"},
            "evals_per_sec": value * evals_page,
            "e2e": {"value": e2e_value, "unit": "pages/s", "h2d_bytes_per_step": int(pages_pin.numel()) * world,
                    "d2h_bytes_per_step": int(out_pin.numel() + counts_pin.numel() * 4) * world,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
            "process_hits_device": post, "focr": focr_line, "config5": config5,
        }
    bank.close()
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pages", type=int, default=100, help="pages per GPU per step")
    ap.add_argument("--distinct", type=int, default=None, help="distinct synthetic pages to generate (default: all)")
    ap.add_argument("--kernel", default="auto", choices=["auto", "simt", "tcgen05"])
    ap.add_argument("--cpu-stride", type=int, default=8, help="CPU baseline scans every k-th template")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-focr", action="store_true", help="skip the config-4 focr line-decode measurement")
    ap.add_argument("--no-config5", action="store_true", help="skip the config-5 (large template bank) measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
