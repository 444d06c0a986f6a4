/*
 * focr_b200.h -- C ABI of libfocr_b200.so, the B200 (sm_100a) drop-in for font-ocr's native hot path.
 *
 * What it replaces in the reference (all citations are into /root/reference/src):
 *   - ncc.cpp:48-63 / ncc.cpp:253-268  `ncc_8_u8` / `ncc_16_u8`, bound by the Rust host through the
 *     `unsafe extern "C"` block ncc.rs:92-126 and called from `Searcher::search_c_u8` ncc.rs:346-390.
 *     Section 1 below exports THE SAME TWO SYMBOLS WITH THE SAME SIGNATURES, so the unmodified Rust
 *     extern block links against this library.
 *   - the per-(page, offset, letter) call loop of `get_hits` ncc.rs:587-702 around that kernel:
 *     one kernel call per template is launch-bound on a GPU, so Section 2 is the batched form a
 *     maintainer would switch `get_hits` to (INTEGRATION.md shows the Rust binding).
 *   - `score_glyph` + `sum_of_squares` inside `decode_line` main.rs:87-181 (in-process Rust today,
 *     no FFI): Section 3 is the new entry the focr binary would call.
 *
 * Conventions: plain C types only; no exceptions cross the boundary; every function that can
 * fail returns an int status (FOCR_OK == 0) and leaves a message in focr_last_error().
 * All buffers named `host` are caller-owned host memory that the callee borrows for the call
 * (same ownership as the reference, ncc.rs:128-141,239-245).  Device memory, pinned staging and
 * streams live inside the opaque context.  There is no CPU fallback: without a CUDA device every
 * compute entry returns FOCR_ERR_CUDA.
 */
#ifndef FOCR_B200_H
#define FOCR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ncc.cpp:7-10 `struct Match` == ncc.rs:66-72 `MatchC` (#[repr(C)]): 8 bytes. */
typedef struct focr_match {
    uint16_t x, y;
    float similarity;
} focr_match;

enum {
    FOCR_OK = 0,
    FOCR_ERR_CUDA = 1,        /* CUDA runtime/driver error or no device */
    FOCR_ERR_ARG = 2,         /* NULL / zero-sized / inconsistent argument */
    FOCR_ERR_UNSUPPORTED = 3, /* shape outside the supported range (see focr_limits) */
    FOCR_ERR_NOMEM = 4
};

/* ------------------------------------------------------------------------------------------
 * Section 1 -- compat shim: the reference's FFI, symbol for symbol (ncc.cpp:48-63, 253-268).
 *
 * Semantics kept: `reference` is the INVERTED page (255 - gray, ncc.rs:887-892), r_w x r_h,
 * row-major; `needle_u8` is N x n_h bytes (N = 8 resp. 16), rows zero-padded (ncc.rs:925-935);
 * the scan covers y in [1, r_h-n_h], x in [1, r_w-n_w]; a window is a hit iff its f64 similarity
 * is != +inf and > (double)threshold; hits are written in (y, x) raster order; when the n_out-th
 * hit is written the call returns n_out at once (ncc.cpp:225-227); otherwise it returns the count.
 * `acc`, `patch_sum`, `patch_rnorm` and `start_end` are accepted for signature compatibility and
 * ignored (may be NULL): the window statistics are recomputed on the device from the page, which
 * gives identical results because start/end only skip windows whose similarity is NaN
 * (SURVEY.md section 8a K3).  Thread-safe (one internal context per device, serialised by a mutex).
 * On a CUDA failure the shim prints the error to stderr and aborts, like the reference's
 * `.unwrap()` call sites would on an impossible state -- it never returns a made-up count.
 * ------------------------------------------------------------------------------------------ */
size_t ncc_8_u8(uint8_t *reference, size_t r_w, size_t r_h, uint8_t *needle_u8, size_t n_w, size_t n_h,
                uint32_t *acc, size_t acc_len, uint32_t *patch_sum, double *patch_rnorm,
                uint16_t *start_end, float threshold, focr_match *out, size_t n_out);
size_t ncc_16_u8(uint8_t *reference, size_t r_w, size_t r_h, uint8_t *needle_u8, size_t n_w, size_t n_h,
                 uint32_t *acc, size_t acc_len, uint32_t *patch_sum, double *patch_rnorm,
                 uint16_t *start_end, float threshold, focr_match *out, size_t n_out);

/* ------------------------------------------------------------------------------------------
 * Section 2 -- batched NCC scan (replaces the loop ncc.rs:587-702 around search_c_u8).
 * ------------------------------------------------------------------------------------------ */
typedef struct focr_ctx focr_ctx;   /* one per GPU; calls on one context are serialised by the caller */
typedef struct focr_bank focr_bank; /* device-resident template bank == the (glyph, shift) raster cache */

typedef struct focr_limits {
    uint32_t max_template_w, max_template_h; /* ncc.rs:392 panics above 16; lifted here */
    uint32_t max_page_w, max_page_h;         /* u16 coordinates, ncc.cpp:8 */
    uint32_t max_n_out;
} focr_limits;

/* which correlation kernel a context uses */
enum {
    FOCR_KERNEL_AUTO = 0,  /* tcgen05 where the shape is supported, else SIMT */
    FOCR_KERNEL_SIMT = 1,  /* dp4a kernel (scan_simt.cu) */
    FOCR_KERNEL_TCGEN05 = 2 /* tcgen05 integer-MMA kernel (scan_tc.cu); FOCR_ERR_UNSUPPORTED if it cannot run the shape */
};

const char *focr_version(void);
const char *focr_last_error(void); /* thread-local */
void focr_get_limits(focr_limits *out);

/* A context owns one GPU's streams, scratch and pinned staging.  It is NOT re-entrant: calls that take the same context must
 * not overlap (one host thread per context at a time -- the reference's rayon workers map to one context per GPU, see
 * focr_multi below, which runs its contexts from its own host threads).  Different contexts are independent. */
int focr_ctx_create(int device, focr_ctx **out);
void focr_ctx_destroy(focr_ctx *ctx);
int focr_ctx_set_kernel(focr_ctx *ctx, int kernel);
/* the cudaStream_t all of this context's device work is enqueued on (for callers that time with
 * CUDA events or interleave their own work) */
void *focr_ctx_stream(focr_ctx *ctx);
int focr_ctx_sync(focr_ctx *ctx);
/* number of kernel launches this context has issued so far (bench.py's gpu_launches) */
uint64_t focr_ctx_launch_count(const focr_ctx *ctx);
/* Per-stage device timing: when enabled, CUDA events are recorded on the context's stream around
 * each stage of the pipeline.  focr_ctx_profile_read synchronises, adds up the event intervals
 * recorded since the last read into ms_out[FOCR_STAGE_*] / launches_out[FOCR_STAGE_*] and resets.
 * This is how bench.py measures the correlation kernel's average launch duration live. */
/* FOCR_STAGE_EXACT (the exact f64 pass over the tensor-core screen's survivors) is a sub-interval of FOCR_STAGE_SCAN. */
/* FOCR_STAGE_DECODE is the focr line-decode kernel (focr_decode_pages). */
enum { FOCR_STAGE_INVERT = 0, FOCR_STAGE_STATS = 1, FOCR_STAGE_SCAN = 2, FOCR_STAGE_FINALIZE = 3, FOCR_STAGE_EXACT = 4,
       FOCR_STAGE_DECODE = 5, FOCR_N_STAGES = 6 };
int focr_ctx_profile(focr_ctx *ctx, int enable);
int focr_ctx_profile_read(focr_ctx *ctx, double *ms_out, uint64_t *launches_out);

/* Upload a template bank.  Template i is the A8 canvas the reference's `render` (ncc.rs:143-196)
 * returns: n_w[i] x n_h[i] bytes, rows tightly packed (stride n_w[i]), starting at
 * pixels + offsets[i].  Order is the reference's iteration order (offset index, alphabet index,
 * ncc.rs:587,630); every result below is indexed by this template index.  Templates may have
 * different sizes (ncc.rs:600-626: the box can change with the subpixel offset); they are grouped
 * by size internally. */
int focr_bank_create(focr_ctx *ctx, const uint8_t *pixels, const uint64_t *offsets, const uint16_t *n_w,
                     const uint16_t *n_h, uint32_t n_templates, focr_bank **out);
void focr_bank_destroy(focr_bank *bank);
uint32_t focr_bank_size(const focr_bank *bank);

/* Scan n_pages pages of r_w x r_h GRAY pixels (what `image::open(..).into_luma8()` yields,
 * ncc.rs:575; the library applies image_to_u8's 255-p itself) against every template of the bank.
 * pages_host + p*page_stride is page p, rows tightly packed.
 * out_host[(p*T + t)*n_out + k] is the k-th hit of template t on page p in the reference's
 * emission order ((y, x) raster order, truncated at n_out exactly like ncc.cpp:225-227);
 * counts_host[p*T + t] is what search_c_u8 would have got back as n_matches (== n_out when full).
 * The timed end-to-end path: H2D of the pages, all kernels, D2H of the match lists.
 * The host buffers may be pageable (a Rust Vec<u8>, ncc.rs:575) or pinned: pageable ones are staged through
 * pinned buffers the context owns (a few host threads copy while the previous chunk computes); pinned ones
 * (cudaHostAlloc / cudaHostRegister) are used for the DMA directly. */
int focr_ncc_scan(focr_ctx *ctx, const focr_bank *bank, const uint8_t *pages_host, size_t page_stride,
                  uint32_t r_w, uint32_t r_h, uint32_t n_pages, float threshold, uint32_t n_out,
                  focr_match *out_host, uint32_t *counts_host);

/* Page-locking for callers without a CUDA binding of their own (the Rust side of INTEGRATION.md): pin a buffer the
 * caller already owns -- the Vec<u8> of pages (ncc.rs:575), the match / count arrays -- so that focr_ncc_scan and
 * focr_decode_pages DMA from / into it directly instead of staging it (config 3: 2070 instead of ~1750 pages/s), or
 * allocate pinned memory outright.  Registration costs ~0.1 ms per MB once; keep it for the lifetime of the buffer.
 * focr_pin_unregister / focr_pin_free must see the same pointer.  Pinned memory is valid for every context. */
int focr_pin_register(focr_ctx *ctx, void *ptr, size_t bytes);
int focr_pin_unregister(focr_ctx *ctx, void *ptr);
int focr_pin_alloc(focr_ctx *ctx, size_t bytes, void **out);
int focr_pin_free(focr_ctx *ctx, void *ptr);

/* Same scan with the pages already resident in device memory (gray, row pitch `pitch` bytes) and
 * the results left in device memory.  All work is enqueued on focr_ctx_stream() back to back; the
 * call then waits ONCE for the stream and checks the chunks' overflow flags (a candidate or hit list
 * that overflowed grows and the scan is repeated), so the results are complete when it returns.
 * This is what bench.py times as `value` (inputs in HBM when the timed region starts). */
int focr_ncc_scan_device(focr_ctx *ctx, const focr_bank *bank, const uint8_t *pages_dev, size_t page_stride,
                         size_t pitch, uint32_t r_w, uint32_t r_h, uint32_t n_pages, float threshold,
                         uint32_t n_out, focr_match *out_dev, uint32_t *counts_dev);

/* ---- one process, several GPUs (csrc/multi.cpp): the reference's page-parallel driver, `pages.par_iter().map(..)`
 * + `sort_by_key(page index)` (ncc.rs:839-847, main.rs:443-468).  A focr_multi owns one context per device; banks are
 * replicated to every device once; a scan splits the batch into contiguous page blocks (sizes differ by at most one,
 * lower devices first -- focr_multi_page_block reports them), every device runs its block through focr_ncc_scan /
 * focr_decode_pages from its own host thread, and the results land at the block's place in the caller's arrays: the
 * output is indexed by page exactly like the single-device call.  Nothing is exchanged between GPUs (no collective).
 * devices == NULL: devices 0 .. n_devices-1, or every visible device when n_devices == 0.  On failure the message of the
 * first failing device is left in focr_last_error(). */
typedef struct focr_multi focr_multi;
typedef struct focr_multi_bank focr_multi_bank;
int focr_multi_create(const int *devices, uint32_t n_devices, focr_multi **out);
void focr_multi_destroy(focr_multi *m);
uint32_t focr_multi_size(const focr_multi *m);
focr_ctx *focr_multi_ctx(focr_multi *m, uint32_t i);
void focr_multi_page_block(const focr_multi *m, uint32_t n_pages, uint32_t i, uint32_t *first, uint32_t *count);
int focr_multi_bank_create(focr_multi *m, const uint8_t *pixels, const uint64_t *offsets, const uint16_t *n_w,
                           const uint16_t *n_h, uint32_t n_templates, focr_multi_bank **out);
void focr_multi_bank_destroy(focr_multi_bank *bank);
int focr_multi_ncc_scan(focr_multi *m, const focr_multi_bank *bank, const uint8_t *pages_host, size_t page_stride,
                        uint32_t r_w, uint32_t r_h, uint32_t n_pages, float threshold, uint32_t n_out,
                        focr_match *out_host, uint32_t *counts_host);

/* Parity probes (tests/): the raw quantities north_star wants bit-exact.
 * window stats for one page and one box size: s_p[y*r_w+x], s2_p[y*r_w+x] for x<=r_w-n_w, y<=r_h-n_h
 * (other entries are left untouched); patch_rnorm as the reference's f64 plane (ncc.rs:306-312). */
int focr_window_stats(focr_ctx *ctx, const uint8_t *page_gray_host, uint32_t r_w, uint32_t r_h, uint32_t n_w,
                      uint32_t n_h, uint32_t *s_p_host, uint64_t *s2_p_host, double *patch_rnorm_host);
/* raw correlation numerators acc[y*r_w+x] (ncc.cpp:108-166) of template t for one page */
int focr_ncc_numerators(focr_ctx *ctx, const focr_bank *bank, uint32_t t, const uint8_t *page_gray_host,
                        uint32_t r_w, uint32_t r_h, uint32_t *acc_host);

/* ------------------------------------------------------------------------------------------
 * Section 3 -- focr least-squared-distance line decode (main.rs:87-181 on the device).
 *
 * The glyph bank is the (glyph, subpixel shift) raster cache README.md:44 asks for.  font-kit hands
 * FreeType the translation `origin + pos` (main.rs:102) as a 26.6 delta `d = (int)(tx * 64.0f)`
 * (f32 arithmetic, truncation); the bitmap only depends on the phase d & 63 and moves by whole pixels
 * with d >> 6.  So for each glyph g < n_glyphs (alphabet order, main.rs:125-128) and each phase
 * s in 0..63 the bank holds the FreeType bitmap rendered with delta (s, -origin.y*64) and its
 * placement on the line canvas for d >> 6 == 0: left = bitmap_left, top = -bitmap_top.
 * origin_x is main.rs:147's origin.x (an integer: minus the left edge of the alphabet's raster
 * bounds); the kernel evaluates d from f32(origin_x + pos) exactly like the host would.
 * advance_px[g] is `advance(g).x / units_per_em * size * kern_x` evaluated in f32 in that order
 * (main.rs:176-178).
 * ------------------------------------------------------------------------------------------ */
typedef struct focr_glyph_raster {
    uint64_t offset;    /* into `pixels`: rows tightly packed, w*h bytes */
    int16_t left, top;  /* top-left of the bitmap on the line canvas when (d >> 6) == 0 */
    uint16_t w, h;
} focr_glyph_raster;

typedef struct focr_glyph_bank focr_glyph_bank;

int focr_glyph_bank_create(focr_ctx *ctx, const uint8_t *pixels, size_t n_pixel_bytes,
                           const focr_glyph_raster *rasters /* [n_glyphs][64] */, const float *advance_px,
                           uint32_t n_glyphs, int32_t origin_x, focr_glyph_bank **out);
void focr_glyph_bank_destroy(focr_glyph_bank *bank);

/* Decode rectangles (x_start, y_start + i*line_advance, width, line_height), i = 0.. of each page
 * exactly like decode_image main.rs:183-218: crop clamped to the page, stop at the first
 * zero-height crop, skip all-white strips, stop at the first empty decode.
 * glyphs_host[(p*max_lines + l)*max_cells + c] = alphabet index chosen for cell c of the l-th
 * DECODED line of page p; n_cells_host[p*max_lines + l] its length; line_y_host[...] its y;
 * n_lines_host[p] the number of decoded lines.  Lines beyond max_lines / cells beyond max_cells
 * make the call fail with FOCR_ERR_ARG rather than truncate silently. */
int focr_decode_pages(focr_ctx *ctx, const focr_glyph_bank *bank, const uint8_t *pages_host, size_t page_stride,
                      uint32_t r_w, uint32_t r_h, uint32_t n_pages, uint32_t x_start, uint32_t y_start,
                      uint32_t width, uint32_t line_height, uint32_t line_advance, uint32_t max_lines,
                      uint32_t max_cells, uint16_t *glyphs_host, uint32_t *n_cells_host,
                      uint32_t *line_y_host, uint32_t *n_lines_host);

/* focr on several GPUs: see focr_multi above (main.rs:443-468). */
typedef struct focr_multi_glyph_bank focr_multi_glyph_bank;
int focr_multi_glyph_bank_create(focr_multi *m, const uint8_t *pixels, size_t n_pixel_bytes,
                                 const focr_glyph_raster *rasters, const float *advance_px, uint32_t n_glyphs,
                                 int32_t origin_x, focr_multi_glyph_bank **out);
void focr_multi_glyph_bank_destroy(focr_multi_glyph_bank *bank);
int focr_multi_decode_pages(focr_multi *m, const focr_multi_glyph_bank *bank, const uint8_t *pages_host, size_t page_stride,
                            uint32_t r_w, uint32_t r_h, uint32_t n_pages, uint32_t x_start, uint32_t y_start,
                            uint32_t width, uint32_t line_height, uint32_t line_advance, uint32_t max_lines,
                            uint32_t max_cells, uint16_t *glyphs_host, uint32_t *n_cells_host,
                            uint32_t *line_y_host, uint32_t *n_lines_host);

/* Device-side process_hits (ncc.rs:723-786 with partition_by, ncc.rs:1036-1052) for a batch of pages whose match
 * lists are still resident in HBM -- the out_dev / counts_dev buffers focr_ncc_scan_device filled (template order =
 * the reference's get_hits order, ncc.rs:675-681).  Anchor filter, the two stable sorts (as one radix sort by
 * (page, y, x, get_hits position)), first-element-anchored overlap groups and the last-maximum pick all run on the GPU;
 * only the surviving hits come back:
 *   line_page_host[l]    page of output line l (lines are ordered by page, then y; every distinct y is a line)
 *   line_start_host[l]   index of the line's first survivor in sel_* (n_lines + 1 entries)
 *   sel_tpl_host[k], sel_host[k]   bank template index and {x, y, similarity} of survivor k, x ascending in a line
 * n_lines_out / n_sel_out always receive the sizes; if line_cap or sel_cap is too small the call returns
 * FOCR_ERR_NOMEM and the caller retries with larger arrays.  A page without any anchor line simply has no lines here
 * (the reference panics in partition_by, ncc.rs:1040; the host mirrors keep that behaviour).  At most 1024 pages per call,
 * T * n_out <= 2^22. */
int focr_process_hits_device(focr_ctx *ctx, const focr_match *matches_dev, const uint32_t *counts_dev, uint32_t T,
                             uint32_t n_out, uint32_t n_pages, float anchor_threshold, int32_t overlap, uint32_t line_cap,
                             uint32_t sel_cap, uint32_t *n_lines_out, uint32_t *n_sel_out, uint32_t *line_page_host,
                             uint32_t *line_start_host, uint32_t *sel_tpl_host, focr_match *sel_host);

/* main.rs:510-516 `sum_of_squares` for n_pairs pairs of equal-length strips (parity probe; the
 * decode kernel uses the algebraically identical Sum(ref^2) - 2*dot + Sum(g^2) form). */
int focr_sum_of_squares(focr_ctx *ctx, const uint8_t *xs_host, const uint8_t *ys_host, size_t len,
                        uint32_t n_pairs, int64_t *out_host);

/* ------------------------------------------------------------------------------------------
 * Section 4 -- C hooks into the C++ host mirror (font-ocr_b200/host/focr_host.hpp), used by tests.
 * The mirror itself (Searcher, get_hits, process_hits, partition_by, decode_image_vec) is a C++
 * API with the reference's names; these two entries expose it to ctypes.
 * ------------------------------------------------------------------------------------------ */
/* process_hits (ncc.rs:723-786) on n hits given in get_hits order.  out_index receives, line after
 * line, the indices (into the input arrays) of the hits that survive; line l is
 * out_index[line_offsets[l] .. line_offsets[l+1]).  out_index and line_offsets need n and n+1
 * entries.  Where the reference panics (no anchor line at all -> partition_by on an empty slice,
 * ncc.rs:1040) the call returns FOCR_ERR_ARG with "panic: ..." in focr_last_error(). */
int focr_host_process_hits(const int32_t *xs, const int32_t *ys, const float *sims, const uint32_t *letters,
                           uint32_t n, float anchor_threshold, int32_t overlap, uint32_t *out_index,
                           uint32_t *line_offsets, uint32_t *n_lines);
/* EXTENSION, opt-in (the reference "does not currently detect spaces", README.md:46): the text of one output line of
 * process_hits (hits in x order) with round(excess / space_px) spaces inserted wherever the pen travel between two hits
 * exceeds the left glyph's advance.  adv_letters / adv_px give the advance in pixels per letter; out receives code points
 * (n_out is always set; FOCR_ERR_NOMEM if out_cap is too small).  space_px <= 0 inserts nothing. */
int focr_host_line_text_with_spaces(const int32_t *xs, const uint32_t *letters, uint32_t n, const uint32_t *adv_letters,
                                    const float *adv_px, uint32_t n_adv, float space_px, uint32_t *out, uint32_t out_cap,
                                    uint32_t *n_out);
/* Searcher::new + Searcher::search_c_u8 (ncc.rs:231-261, 332-404) for one gray page and one tight
 * n_w x n_h needle; out needs 1024 entries.  Widths above 16 return FOCR_ERR_UNSUPPORTED
 * ("panic: not handled", ncc.rs:392). */
int focr_host_search_c_u8(const uint8_t *gray, uint32_t r_w, uint32_t r_h, const uint8_t *needle, uint32_t n_w,
                          uint32_t n_h, float threshold, focr_match *out, uint32_t *n_out);

/* ------------------------------------------------------------------------------------------
 * Section 5 -- C++ FreeType driver (font-ocr_b200/host/focr_raster.cpp): the template and glyph-raster producers of the
 * path's input side, i.e. the (glyph, subpixel shift) raster cache (README.md:44).  What a
 * maintainer would call instead of re-rendering every template for every page (ncc.rs:561,631) and every glyph for every
 * cell (main.rs:98-106).  freetype_so: path of a libfreetype shared object (opened with dlopen; there are no FreeType
 * headers in this image).  Where the reference .unwrap()s a missing glyph the calls return FOCR_ERR_ARG with "panic: ...".
 * PARITY UNPINNED for the rasters (font-kit / pathfinder semantics restated from memory); pinned byte for byte to the Python
 * producer every test uses (tests/test_raster_native.py).
 * ------------------------------------------------------------------------------------------ */
typedef struct focr_host_font focr_host_font;
typedef struct focr_host_tbank focr_host_tbank;
typedef struct focr_host_gbank focr_host_gbank;
int focr_host_font_open(const char *freetype_so, const char *font_path, focr_host_font **out);
void focr_host_font_close(focr_host_font *font);
/* the reference's --hinting (HintingOptions::Full, ncc.rs:547-551, main.rs:394-398): later banks are rasterised hinted */
void focr_host_font_set_hinting(focr_host_font *font, int on);
/* per-letter metrics in pixels at `size`: the left bearing --raw prints (ncc.rs:683-698) and the f32 pen advance (main.rs:176-178) */
int focr_host_font_glyph_metrics(focr_host_font *font, uint32_t letter, float size, float *bearing_x_px, float *advance_px);
/* every template get_hits renders for one page (ncc.rs:587-640) in its iteration order (offset index, alphabet index);
 * box_mode 0 = BoxSize::Alphabet (ncc.rs:600-626), 1 = Font (ncc.rs:589-599), 2 = Char (ncc.rs:627) */
int focr_host_tbank_render(focr_host_font *font, float size, const uint32_t *alphabet, uint32_t n_alphabet, uint32_t x_bits,
                           uint32_t y_bits, int box_mode, int pad_x, int pad_y, focr_host_tbank **out);
uint32_t focr_host_tbank_count(const focr_host_tbank *bank);
uint64_t focr_host_tbank_pixel_bytes(const focr_host_tbank *bank);
/* the arrays focr_bank_create takes, plus the letters and the corrected y offsets --raw prints; any pointer may be NULL */
int focr_host_tbank_get(const focr_host_tbank *bank, uint8_t *pixels, uint64_t *offsets, uint16_t *n_w, uint16_t *n_h,
                        uint32_t *letters, float *corrected_y);
void focr_host_tbank_free(focr_host_tbank *bank);
/* focr's raster cache: 64 horizontal 26.6 phases per alphabet glyph + f32 advances -- the arrays focr_glyph_bank_create takes
 * (rasters [n_alphabet][64], advance_px [n_alphabet], origin_xy = main.rs:147's origin) */
int focr_host_gbank_render(focr_host_font *font, float size, const uint32_t *alphabet, uint32_t n_alphabet, float kern_x,
                           focr_host_gbank **out);
uint64_t focr_host_gbank_pixel_bytes(const focr_host_gbank *bank);
int focr_host_gbank_get(const focr_host_gbank *bank, uint8_t *pixels, focr_glyph_raster *rasters, float *advance_px,
                        int32_t *origin_xy);
void focr_host_gbank_free(focr_host_gbank *bank);

#ifdef __cplusplus
}
#endif
#endif /* FOCR_B200_H */
