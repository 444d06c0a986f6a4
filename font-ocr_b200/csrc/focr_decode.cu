// focr_decode.cu -- focr's least-squared-distance line decode on the device (main.rs:87-218).
//
// Reference, per cell of a line: for every alphabet glyph, clear a line-sized canvas, rasterise the
// glyph at the pen position with FreeType, and take sum_of_squares(ref strip, canvas) over the WHOLE
// canvas (main.rs:87-110, 510-516); keep the first minimum (main.rs:159-172); advance the pen by that
// glyph's advance in f32 (main.rs:176-178).
//
// Here: the rasters come from the (glyph, 26.6 sub-pixel phase) bank the host uploads once -- the cache
// README.md:44 asks for -- and the score uses the algebraically identical form
//     SSD = Sum(ref^2) - 2*Sum_box(ref*g) + Sum_box(g^2)
// where only the glyph's clipped bitmap box contributes and Sum(ref^2) is the same for every glyph, so
// the argmin (with the reference's first-minimum tie-break) is decided by the exact integer
// Sum_box(g*(g - 2*ref)).  One warp walks one line: lanes score different glyphs of a cell in parallel,
// the pen walk itself is sequential (the advance depends on the chosen glyph).
// The device copy of the bank pads every bitmap row to a multiple of 4 bytes (zeros) at a 4-byte-aligned offset, so that a
// row that is not clipped horizontally is scored four pixels at a time: Sum(g*g) and Sum(g*ref) with __dp4a on the bank
// word and the byte-shifted strip word.  Horizontally clipped cells (a line's first / last glyph) take the per-pixel loop.
#include <algorithm>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

namespace focr {

struct GlyphBankDev {
    const uint8_t *pixels;
    const focr_glyph_raster *rasters;  // [n_glyphs][64]
    const float *advance_px;           // [n_glyphs]
    uint32_t n_glyphs;
    // row tasks: for every 26.6 phase the bitmap rows of all glyphs, in glyph order, as
    // {byte offset of the row in `pixels`, glyph | w << 16, (u16)left | (u16)(top + row) << 16, 0}
    const uint4 *tasks;                // [task_off[64]]
    const uint32_t *task_off;          // [65]
    // dense tiles (the fast path): for every phase and glyph the bitmap drawn on a tile_rows x tile_wb byte tile whose
    // top-left corner sits at canvas position (dint + tile_left, 0): [64][n_glyphs][tile_rows][tile_wb], zero filled.
    // Rows above the canvas (never visible, Canvas::blit_from clips them) are dropped.  nullptr: bitmaps too large.
    const uint8_t *tiles;
    int tile_rows, tile_wb, tile_left;
};

constexpr uint32_t FD_MAX_SCORES = 256;   // glyphs per bank the row-task path keeps scores for (per warp, shared memory)

constexpr int FD_WARPS = 4;  // lines per block

struct DecodeArgs {
    const uint8_t *pages;  // gray, tight rows
    size_t page_stride;
    uint32_t r_w, r_h, n_pages;
    uint32_t x_start, y_start, width, line_height, line_advance;
    uint32_t mem_y0, mem_line_rows;   // where line 0 starts in a device page and how many rows lie between consecutive lines there
                                      // (the device copy may hold only the rectangles' rows: line_height of every line_advance)
    uint32_t max_lines, max_cells;
    GlyphBankDev bank;
    float origin_x;        // main.rs:147 origin.x (an integer value)
    uint16_t *glyphs;      // [page][max_lines][max_cells]
    uint32_t *n_cells;     // [page][max_lines]; 0xFFFFFFFF = strip skipped (all white) or beyond the page
    unsigned int *error;   // set to 1 when a line needs more than max_cells
};

__global__ void __launch_bounds__(FD_WARPS * 32) focr_decode_kernel(DecodeArgs a)
{
    extern __shared__ __align__(16) uint8_t smem[];   // FD_WARPS strips + 16 bytes (the word loads may run past the last strip)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t line = blockIdx.x * FD_WARPS + warp;
    const uint32_t page = blockIdx.y;
    if (line >= a.max_lines) return;
    uint32_t *n_cells_out = a.n_cells + (size_t)page * a.max_lines + line;
    // crop_imm clamps the rectangle to the image (main.rs:201-203)
    const uint32_t y0 = a.y_start + line * a.line_advance;
    const uint32_t xs = min(a.x_start, a.r_w), ys = min(y0, a.r_h);
    const uint32_t w = min(a.width, a.r_w - xs), h = min(a.line_height, a.r_h - ys);
    if (h == 0 || w == 0) {  // height 0 ends the page (main.rs:205-207); width 0 is an all-white strip
        if (lane == 0) *n_cells_out = 0xFFFFFFFFu;
        return;
    }
    // stage the inverted strip (main.rs:150) in shared memory
    uint8_t *ref = smem + (size_t)warp * a.width * a.line_height;
    // per-warp glyph scores of the row-task path, behind the strips (16-byte aligned); nullptr: per-glyph path
    int *scores = a.bank.tasks ? (int *)(smem + (((size_t)FD_WARPS * a.width * a.line_height + 16 + 15) & ~(size_t)15)) +
                                     (size_t)warp * FD_MAX_SCORES
                               : nullptr;
    const uint8_t *src = a.pages + (size_t)page * a.page_stride + ((size_t)a.mem_y0 + (size_t)line * a.mem_line_rows) * a.r_w + xs;
    uint32_t any_ink = 0;
    for (uint32_t i = lane; i < w * h; i += 32) {
        const uint32_t r = i / w, c = i - r * w;
        const uint8_t v = 255 - src[(size_t)r * a.r_w + c];
        ref[i] = v;
        any_ink |= v;
    }
    any_ink = __reduce_or_sync(0xffffffffu, any_ink);
    __syncwarp();
    if (any_ink == 0) {  // all pixels == 255: skipped (main.rs:208-211)
        if (lane == 0) *n_cells_out = 0xFFFFFFFFu;
        return;
    }
    uint16_t *out = a.glyphs + ((size_t)page * a.max_lines + line) * a.max_cells;
    float pos = 0.f;  // Vector2F pen, x component (main.rs:123)
    uint32_t n = 0;
    const float wf = (float)w;
    while (pos < wf) {  // main.rs:158
        // font-kit: translation -> 26.6 by `(x * 64.0) as i32` (truncation); translation = origin + pos
        const float tx = __fadd_rn(a.origin_x, pos);
        const int d26 = (int)__fmul_rn(tx, 64.0f);
        const int dint = d26 >> 6, frac = d26 & 63;
        int best_score = 0x7fffffff;
        uint32_t best_g = 0xFFFFFFFFu;
        if (scores) {
            // Row tasks: the lanes share the bitmap ROWS of all glyphs of this phase evenly (a lane per row, ~23 rows each
            // for a 67-glyph alphabet) instead of whole glyphs (2-3 glyphs of very different sizes per lane: half of the
            // lanes idle); a row's partial score goes to the glyph's slot with a shared-memory atomic.
            for (uint32_t g = lane; g < a.bank.n_glyphs; g += 32) scores[g] = 0;
            __syncwarp();
            const uint32_t t0 = a.bank.task_off[frac], t1 = a.bank.task_off[frac + 1];
            for (uint32_t t = t0 + lane; t < t1; t += 32) {
                const uint4 tk = __ldg(a.bank.tasks + t);
                const uint32_t g = tk.y & 0xFFFFu;
                const int gw_ = (int)(tk.y >> 16);
                const int dx = (int)(int16_t)(tk.z & 0xFFFFu) + dint, ry = (int)(int16_t)(tk.z >> 16);
                if (ry < 0 || ry >= (int)h) continue;   // Canvas::blit_from clips the bitmap to the canvas
                const int bx0 = max(0, -dx), bx1 = min(gw_, (int)w - dx);
                const uint8_t *brow8 = a.bank.pixels + tk.x;
                int part = 0;
                if (bx0 == 0 && bx1 == gw_) {
                    const uint32_t *brow = (const uint32_t *)brow8;
                    const uintptr_t ra = (uintptr_t)(ref + ry * (int)w + dx);   // the strip bytes under this bitmap row
                    const uint32_t *rw = (const uint32_t *)(ra & ~(uintptr_t)3);
                    const int sh = (int)(ra & 3) * 8, nwords = (gw_ + 3) >> 2;
                    uint32_t lo = rw[0], g2 = 0, gr_dot = 0;
                    for (int q = 0; q < nwords; q++) {
                        const uint32_t hi = rw[q + 1];
                        const uint32_t gw = __ldg(brow + q);
                        g2 = __dp4a(gw, gw, g2);
                        gr_dot = __dp4a(gw, __funnelshift_r(lo, hi, sh), gr_dot);   // padding bytes of gw are 0
                        lo = hi;
                    }
                    part = (int)g2 - 2 * (int)gr_dot;
                } else {
                    const uint8_t *rrow = ref + ry * (int)w + dx;
                    for (int bx = bx0; bx < bx1; bx++) {
                        const int gv = brow8[bx], rv = rrow[bx];
                        part += gv * (gv - 2 * rv);
                    }
                }
                if (part) atomicAdd(scores + g, part);
            }
            __syncwarp();
            for (uint32_t g = lane; g < a.bank.n_glyphs; g += 32) {
                const int score = scores[g];
                if (score < best_score) {  // within a lane glyph indices ascend: strict < keeps the first minimum
                    best_score = score;
                    best_g = g;
                }
            }
            __syncwarp();
        } else
        for (uint32_t g = lane; g < a.bank.n_glyphs; g += 32) {
            const focr_glyph_raster gr = a.bank.rasters[(size_t)g * 64 + frac];
            const uint8_t *bm = a.bank.pixels + gr.offset;
            const int dx = gr.left + dint, dy = gr.top;
            // Canvas::blit_from clips the bitmap to the canvas
            const int bx0 = max(0, -dx), bx1 = min((int)gr.w, (int)w - dx);
            const int by0 = max(0, -dy), by1 = min((int)gr.h, (int)h - dy);
            int score = 0;
            const int pitch = (gr.w + 3) & ~3;       // device bank: rows padded with zeros to whole words
            if (bx0 == 0 && bx1 == (int)gr.w) {
                const int nwords = pitch >> 2;
                for (int by = by0; by < by1; by++) {
                    const uint32_t *brow = (const uint32_t *)(bm + by * pitch);
                    const uintptr_t ra = (uintptr_t)(ref + (dy + by) * (int)w + dx);   // the strip bytes under this bitmap row
                    const uint32_t *rw = (const uint32_t *)(ra & ~(uintptr_t)3);
                    const int sh = (int)(ra & 3) * 8;
                    uint32_t lo = rw[0];
                    uint32_t g2 = 0, gr_dot = 0;
                    for (int q = 0; q < nwords; q++) {
                        const uint32_t hi = rw[q + 1];
                        const uint32_t gw = __ldg(brow + q);
                        g2 = __dp4a(gw, gw, g2);
                        gr_dot = __dp4a(gw, __funnelshift_r(lo, hi, sh), gr_dot);   // padding bytes of gw are 0
                        lo = hi;
                    }
                    score += (int)g2 - 2 * (int)gr_dot;
                }
            } else {
                for (int by = by0; by < by1; by++) {
                    const uint8_t *brow = bm + by * pitch;
                    const uint8_t *rrow = ref + (dy + by) * w + dx;
                    for (int bx = bx0; bx < bx1; bx++) {
                        const int gv = brow[bx], rv = rrow[bx];
                        score += gv * (gv - 2 * rv);
                    }
                }
            }
            if (score < best_score) {  // within a lane glyph indices ascend: strict < keeps the first minimum
                best_score = score;
                best_g = g;
            }
        }
        // warp argmin with the reference's tie-break: min_by_key returns the FIRST minimum (main.rs:159-172)
#pragma unroll
        for (int d = 16; d; d >>= 1) {
            const int os = __shfl_xor_sync(0xffffffffu, best_score, d);
            const uint32_t og = __shfl_xor_sync(0xffffffffu, best_g, d);
            if (os < best_score || (os == best_score && og < best_g)) {
                best_score = os;
                best_g = og;
            }
        }
        if (n >= a.max_cells) {
            if (lane == 0) atomicExch(a.error, 1u);
            break;
        }
        if (lane == 0) out[n] = (uint16_t)best_g;
        n++;
        pos = __fadd_rn(pos, a.bank.advance_px[best_g]);  // main.rs:176-178 (advance precomputed in f32)
    }
    if (lane == 0) *n_cells_out = n;
}

// ---------------------------------------------------------------------------------------------
// The fast path: every cell is a small integer GEMV.  For the cell's 26.6 phase the bank holds one dense
// tile_rows x tile_wb byte tile per glyph (zeros around the bitmap), all on the same canvas grid; the strip bytes under
// that grid -- the WINDOW, tile_wb columns from canvas column dint + tile_left -- are aligned once per cell into shared
// memory and then every lane scores its glyphs against it:  Sum g*g - 2 * Sum g*ref  with two __dp4a per 4 pixels, one
// 16-byte bank load and one 16-byte (broadcast) window load per 16 pixels.  Same integers as the reference's
// sum_of_squares over the whole canvas minus the glyph-independent Sum ref^2: the strip is zero padded on
// both sides so that pixels outside the canvas add nothing to Sum g*ref, and in the cells whose window crosses the
// canvas edge Sum g*g only counts the columns inside (Canvas::blit_from clips the bitmap, main.rs:98-106).
constexpr int FD_PAD = 64;   // zero columns on either side of a staged strip (>= tile_wb + |origin + tile_left|, checked on the host)

template <int WORDS>   // 16-byte groups per tile row: tile_wb = 16 * WORDS
__global__ void __launch_bounds__(FD_WARPS * 32) focr_decode_tile_kernel(DecodeArgs a)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t line = blockIdx.x * FD_WARPS + warp;
    const uint32_t page = blockIdx.y;
    if (line >= a.max_lines) return;
    uint32_t *n_cells_out = a.n_cells + (size_t)page * a.max_lines + line;
    // crop_imm clamps the rectangle to the image (main.rs:201-203)
    const uint32_t y0 = a.y_start + line * a.line_advance;
    const uint32_t xs = min(a.x_start, a.r_w), ys = min(y0, a.r_h);
    const uint32_t w = min(a.width, a.r_w - xs), h = min(a.line_height, a.r_h - ys);
    if (h == 0 || w == 0) {  // height 0 ends the page (main.rs:205-207); width 0 is an all-white strip
        if (lane == 0) *n_cells_out = 0xFFFFFFFFu;
        return;
    }
    // per warp: the inverted strip with FD_PAD zero columns on both sides (row pitch sp, a multiple of 4), then the window
    const int sp = (int)((a.width + 2 * FD_PAD + 3) & ~3u);
    const int R = a.bank.tile_rows, WB = 16 * WORDS;
    const size_t strip_bytes = ((size_t)sp * a.line_height + 16 + 15) & ~(size_t)15;   // + 16: the word loads may run past the last row
    uint8_t *strip = smem + (size_t)warp * (strip_bytes + (size_t)R * WB);
    uint4 *win = (uint4 *)(strip + strip_bytes);
    {
        uint32_t *z = (uint32_t *)strip;
        for (int i = lane; i < (int)(strip_bytes / 4); i += 32) z[i] = 0u;
    }
    __syncwarp();
    const uint8_t *src = a.pages + (size_t)page * a.page_stride + ((size_t)a.mem_y0 + (size_t)line * a.mem_line_rows) * a.r_w + xs;
    uint32_t any_ink = 0;
    for (uint32_t i = lane; i < w * h; i += 32) {
        const uint32_t r = i / w, c = i - r * w;
        const uint8_t v = 255 - src[(size_t)r * a.r_w + c];   // main.rs:150
        strip[r * sp + FD_PAD + c] = v;
        any_ink |= v;
    }
    any_ink = __reduce_or_sync(0xffffffffu, any_ink);
    __syncwarp();
    if (any_ink == 0) {  // all pixels == 255: skipped (main.rs:208-211)
        if (lane == 0) *n_cells_out = 0xFFFFFFFFu;
        return;
    }
    uint16_t *out = a.glyphs + ((size_t)page * a.max_lines + line) * a.max_cells;
    const int rows = min((int)h, R);            // tile rows on the canvas (a clamped last strip has fewer)
    const int n_glyphs = (int)a.bank.n_glyphs;
    const size_t glyph_stride = (size_t)R * WORDS;                 // uint4 per glyph tile
    float pos = 0.f;  // Vector2F pen, x component (main.rs:123)
    uint32_t n = 0;
    const float wf = (float)w;
    while (pos < wf) {  // main.rs:158
        // font-kit: translation -> 26.6 by `(x * 64.0) as i32` (truncation); translation = origin + pos
        const float tx = __fadd_rn(a.origin_x, pos);
        const int d26 = (int)__fmul_rn(tx, 64.0f);
        const int dint = d26 >> 6, frac = d26 & 63;
        const int x0 = dint + a.bank.tile_left;                    // canvas column of the tile's first column
        // the window: rows x WB strip bytes from column x0, aligned into shared memory (one 4-byte word per lane and step)
        {
            const int nwords = rows * WORDS * 4;
            for (int i = lane; i < nwords; i += 32) {
                const int r = i / (WORDS * 4), q = i - r * (WORDS * 4);
                const uintptr_t p = (uintptr_t)(strip + r * sp + FD_PAD + x0 + 4 * q);
                const uint32_t *wp = (const uint32_t *)(p & ~(uintptr_t)3);
                ((uint32_t *)win)[i] = __funnelshift_r(wp[0], wp[1], (int)(p & 3) * 8);
            }
        }
        __syncwarp();
        const bool edge = x0 < 0 || x0 + WB > (int)w;             // warp-uniform: some tile columns lie outside the canvas
        uint32_t cmask[WORDS * 4];
#pragma unroll
        for (int q = 0; q < WORDS * 4; q++) {
            uint32_t m = 0;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int cx = x0 + 4 * q + b;
                if (cx >= 0 && cx < (int)w) m |= 0xFFu << (8 * b);
            }
            cmask[q] = m;
        }
        const uint4 *tile0 = (const uint4 *)a.bank.tiles + (size_t)frac * n_glyphs * glyph_stride;
        int best_score = 0x7fffffff;
        uint32_t best_g = 0xFFFFFFFFu;
        for (int g = lane; g < n_glyphs; g += 32) {
            const uint4 *tile = tile0 + (size_t)g * glyph_stride;
            uint32_t g2 = 0, gr = 0;
            if (!edge) {
#pragma unroll 4
                for (int r = 0; r < rows; r++) {
#pragma unroll
                    for (int k = 0; k < WORDS; k++) {
                        const uint4 t = __ldg(tile + r * WORDS + k);
                        const uint4 v = win[r * WORDS + k];
                        g2 = __dp4a(t.x, t.x, g2), gr = __dp4a(t.x, v.x, gr);
                        g2 = __dp4a(t.y, t.y, g2), gr = __dp4a(t.y, v.y, gr);
                        g2 = __dp4a(t.z, t.z, g2), gr = __dp4a(t.z, v.z, gr);
                        g2 = __dp4a(t.w, t.w, g2), gr = __dp4a(t.w, v.w, gr);
                    }
                }
            } else {
                for (int r = 0; r < rows; r++) {
#pragma unroll
                    for (int k = 0; k < WORDS; k++) {
                        const uint4 t = __ldg(tile + r * WORDS + k);
                        const uint4 v = win[r * WORDS + k];
                        g2 = __dp4a(t.x & cmask[4 * k], t.x, g2), gr = __dp4a(t.x, v.x, gr);
                        g2 = __dp4a(t.y & cmask[4 * k + 1], t.y, g2), gr = __dp4a(t.y, v.y, gr);
                        g2 = __dp4a(t.z & cmask[4 * k + 2], t.z, g2), gr = __dp4a(t.z, v.z, gr);
                        g2 = __dp4a(t.w & cmask[4 * k + 3], t.w, g2), gr = __dp4a(t.w, v.w, gr);
                    }
                }
            }
            const int score = (int)g2 - 2 * (int)gr;
            if (score < best_score) {  // within a lane glyph indices ascend: strict < keeps the first minimum
                best_score = score;
                best_g = (uint32_t)g;
            }
        }
        // warp argmin with the reference's tie-break: min_by_key returns the FIRST minimum (main.rs:159-172)
#pragma unroll
        for (int d = 16; d; d >>= 1) {
            const int os = __shfl_xor_sync(0xffffffffu, best_score, d);
            const uint32_t og = __shfl_xor_sync(0xffffffffu, best_g, d);
            if (os < best_score || (os == best_score && og < best_g)) {
                best_score = os;
                best_g = og;
            }
        }
        if (n >= a.max_cells) {
            if (lane == 0) atomicExch(a.error, 1u);
            break;
        }
        if (lane == 0) out[n] = (uint16_t)best_g;
        n++;
        pos = __fadd_rn(pos, a.bank.advance_px[best_g]);  // main.rs:176-178 (advance precomputed in f32)
        __syncwarp();   // the window is rewritten by the next cell
    }
    if (lane == 0) *n_cells_out = n;
}

__global__ void __launch_bounds__(256) sum_of_squares_kernel(const uint8_t *xs, const uint8_t *ys, size_t len,
                                                             long long *out)
{
    // main.rs:510-516 for pair blockIdx.x
    const uint8_t *x = xs + (size_t)blockIdx.x * len, *y = ys + (size_t)blockIdx.x * len;
    long long s = 0;
    for (size_t i = threadIdx.x; i < len; i += blockDim.x) {
        const int d = (int)x[i] - (int)y[i];
        s += (long long)(d * d);
    }
    __shared__ long long red[256];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int k = 128; k; k >>= 1) {
        if ((int)threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = red[0];
}

}  // namespace focr

using namespace focr;

// api.cu owns the context; only what this file needs is re-declared here
extern "C" void *focr_ctx_stream(focr_ctx *ctx);
int focr_internal_device(const focr_ctx *ctx);
void focr_internal_count_launch(focr_ctx *ctx, int n);
int focr_internal_fail(int code, const std::string &msg);
void *focr_internal_stage_begin(focr_ctx *ctx, int stage);
void focr_internal_stage_end(void *timer);
void focr_internal_streams(focr_ctx *ctx, cudaStream_t out[3]);   // compute, H2D, D2H
bool focr_internal_host_pinned(const void *p);
void focr_internal_parallel_for(int n, const std::function<void(int)> &fn);   // api.cu: the staging thread pool
int focr_internal_stage_threads();
void focr_internal_parallel_copy(uint8_t *dst, size_t dst_stride, const uint8_t *src, size_t src_stride, size_t row_bytes, size_t rows);

struct focr_glyph_bank {
    focr_ctx *ctx;
    uint8_t *pixels;
    focr_glyph_raster *rasters;
    float *advance;
    uint32_t n_glyphs;
    float origin_x;
    uint4 *tasks;          // row tasks per phase (GlyphBankDev), nullptr when a bitmap is too large for the packing
    uint32_t *task_off;
    uint8_t *tiles;        // dense per-(phase, glyph) tiles (GlyphBankDev), nullptr when a bitmap does not fit a 32 x 32 tile
    int tile_rows, tile_wb, tile_left;
};

#define FCU(call)                                                                                          \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess)                                                                             \
            return focr_internal_fail(FOCR_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

extern "C" int focr_glyph_bank_create(focr_ctx *ctx, const uint8_t *pixels, size_t n_pixel_bytes,
                                      const focr_glyph_raster *rasters, const float *advance_px, uint32_t n_glyphs,
                                      int32_t origin_x, focr_glyph_bank **out)
{
    if (!ctx || !pixels || !rasters || !advance_px || !out || n_glyphs == 0 || n_glyphs > 65535)
        return focr_internal_fail(FOCR_ERR_ARG, "focr_glyph_bank_create: bad argument");
    for (size_t i = 0; i < (size_t)n_glyphs * 64; i++)
        if (rasters[i].offset + (size_t)rasters[i].w * rasters[i].h > n_pixel_bytes)
            return focr_internal_fail(FOCR_ERR_ARG, "focr_glyph_bank_create: raster outside the pixel buffer");
    // the kernels accumulate Sum g*(g - 2*ref) in 32-bit integers (the reference's sum_of_squares is i64, main.rs:510-516):
    // |term| <= 255*255 per pixel, so a bitmap must stay below 2^31 / 65025 = 33025 pixels (a ~180 px glyph)
    for (size_t i = 0; i < (size_t)n_glyphs * 64; i++)
        if ((uint64_t)rasters[i].w * rasters[i].h * 65025ull >= (1ull << 31))
            return focr_internal_fail(FOCR_ERR_UNSUPPORTED, "focr_glyph_bank_create: glyph bitmap of " + std::to_string(rasters[i].w) +
                                                                "x" + std::to_string(rasters[i].h) + " pixels would overflow the 32-bit score");
    FCU(cudaSetDevice(focr_internal_device(ctx)));
    focr_glyph_bank *b = new focr_glyph_bank();
    b->ctx = ctx;
    b->n_glyphs = n_glyphs;
    b->origin_x = (float)origin_x;
    // device copy: every bitmap row padded with zeros to a multiple of 4 bytes, every bitmap at a 4-byte-aligned offset
    std::vector<focr_glyph_raster> rs(rasters, rasters + (size_t)n_glyphs * 64);
    size_t total = 0;
    for (auto &r : rs) total += (size_t)((r.w + 3) & ~3) * r.h;
    std::vector<uint8_t> padded(total ? total : 4, 0);
    size_t off = 0;
    for (size_t i = 0; i < rs.size(); i++) {
        const size_t pitch = (size_t)((rs[i].w + 3) & ~3);
        for (uint32_t y = 0; y < rs[i].h; y++)
            memcpy(padded.data() + off + y * pitch, pixels + rasters[i].offset + (size_t)y * rs[i].w, rs[i].w);
        rs[i].offset = off;
        off += pitch * rs[i].h;
    }
    FCU(cudaMalloc((void **)&b->pixels, padded.size()));
    FCU(cudaMalloc((void **)&b->rasters, (size_t)n_glyphs * 64 * sizeof(focr_glyph_raster)));
    FCU(cudaMalloc((void **)&b->advance, n_glyphs * sizeof(float)));
    // row tasks per phase, glyph order (GlyphBankDev::tasks)
    b->tasks = nullptr, b->task_off = nullptr;
    {
        std::vector<uint4> tasks;
        std::vector<uint32_t> toff(65, 0);
        bool ok = padded.size() < (1ull << 32);
        for (uint32_t ph = 0; ph < 64 && ok; ph++) {
            toff[ph] = (uint32_t)tasks.size();
            for (uint32_t g = 0; g < n_glyphs && ok; g++) {
                const focr_glyph_raster &r = rs[(size_t)g * 64 + ph];
                const uint32_t pitch = (r.w + 3u) & ~3u;
                for (uint32_t y = 0; y < r.h; y++) {
                    const int ry = (int)r.top + (int)y;
                    if (ry < -32768 || ry > 32767) { ok = false; break; }
                    tasks.push_back(make_uint4((uint32_t)(r.offset + (size_t)y * pitch), g | ((uint32_t)r.w << 16),
                                               (uint32_t)(uint16_t)r.left | ((uint32_t)(uint16_t)(int16_t)ry << 16), 0u));
                }
            }
        }
        toff[64] = (uint32_t)tasks.size();
        if (ok) {
            FCU(cudaMalloc((void **)&b->tasks, std::max<size_t>(tasks.size(), 1) * sizeof(uint4)));
            FCU(cudaMalloc((void **)&b->task_off, 65 * 4));
            FCU(cudaMemcpy(b->tasks, tasks.data(), tasks.size() * sizeof(uint4), cudaMemcpyHostToDevice));
            FCU(cudaMemcpy(b->task_off, toff.data(), 65 * 4, cudaMemcpyHostToDevice));
        }
    }
    // dense tiles (focr_decode_tile_kernel): all bitmaps of the bank on one canvas grid of tile_rows x tile_wb bytes whose
    // first column is canvas column dint + tile_left; rows above the canvas are dropped (always clipped)
    b->tiles = nullptr;
    {
        int left = 0x7fffffff, right = -0x7fffffff, bottom = 0;
        for (auto &r : rs)
            if (r.w && r.h) {
                left = std::min(left, (int)r.left);
                right = std::max(right, (int)r.left + (int)r.w);
                bottom = std::max(bottom, (int)r.top + (int)r.h);
            }
        if (left <= right && bottom >= 1 && bottom <= 32 && right - left <= 32) {
            b->tile_left = left;
            b->tile_rows = bottom;
            b->tile_wb = (right - left + 15) & ~15;
            const size_t tile = (size_t)b->tile_rows * b->tile_wb;
            std::vector<uint8_t> tiles((size_t)64 * n_glyphs * tile, 0);
            for (uint32_t ph = 0; ph < 64; ph++)
                for (uint32_t g = 0; g < n_glyphs; g++) {
                    const focr_glyph_raster &r = rasters[(size_t)g * 64 + ph];
                    uint8_t *t = tiles.data() + ((size_t)ph * n_glyphs + g) * tile;
                    for (int y = 0; y < (int)r.h; y++) {
                        const int cy = (int)r.top + y;
                        if (cy < 0) continue;
                        memcpy(t + (size_t)cy * b->tile_wb + ((int)r.left - left), pixels + r.offset + (size_t)y * r.w, r.w);
                    }
                }
            FCU(cudaMalloc((void **)&b->tiles, tiles.size()));
            FCU(cudaMemcpy(b->tiles, tiles.data(), tiles.size(), cudaMemcpyHostToDevice));
        }
    }
    FCU(cudaMemcpy(b->pixels, padded.data(), padded.size(), cudaMemcpyHostToDevice));
    FCU(cudaMemcpy(b->rasters, rs.data(), rs.size() * sizeof(focr_glyph_raster), cudaMemcpyHostToDevice));
    FCU(cudaMemcpy(b->advance, advance_px, n_glyphs * sizeof(float), cudaMemcpyHostToDevice));
    *out = b;
    return FOCR_OK;
}

extern "C" void focr_glyph_bank_destroy(focr_glyph_bank *b)
{
    if (!b) return;
    cudaSetDevice(focr_internal_device(b->ctx));
    cudaFree(b->pixels);
    cudaFree(b->rasters);
    cudaFree(b->advance);
    cudaFree(b->tasks);
    cudaFree(b->task_off);
    cudaFree(b->tiles);
    delete b;
}

// Grow-only scratch per device, two slots (device buffers + pinned staging): a cudaMalloc / cudaFree pair per buffer and
// call costs more than the decode kernel, and cudaFree synchronises the device.
namespace {
constexpr uint32_t FD_CHUNK = 16;   // pages per chunk: 16 x 232 lines / 4 = 928 blocks, one full wave of the decode kernel
struct DecodeSlot {
    void *dev[4] = {nullptr, nullptr, nullptr, nullptr};   // band, glyphs, cells, error flag
    size_t dev_cap[4] = {0, 0, 0, 0};
    void *host[3] = {nullptr, nullptr, nullptr};           // glyphs, cells + error flag, band staging (pageable callers)
    size_t host_cap[3] = {0, 0, 0};
    cudaEvent_t ev_h2d = nullptr, ev_k = nullptr, ev_d2h = nullptr;
    void *dev_buf(int i, size_t want)
    {
        if (want > dev_cap[i]) {
            if (dev[i]) cudaFree(dev[i]);
            dev[i] = nullptr, dev_cap[i] = 0;
            if (cudaMalloc(&dev[i], want + want / 4) != cudaSuccess) return nullptr;
            dev_cap[i] = want + want / 4;
        }
        return dev[i];
    }
    void *host_buf(int i, size_t want)
    {
        if (want > host_cap[i]) {
            if (host[i]) cudaFreeHost(host[i]);
            host[i] = nullptr, host_cap[i] = 0;
            if (cudaHostAlloc(&host[i], want + want / 4, cudaHostAllocDefault) != cudaSuccess) return nullptr;
            host_cap[i] = want + want / 4;
        }
        return host[i];
    }
    bool events()
    {
        if (ev_h2d) return true;
        return cudaEventCreateWithFlags(&ev_h2d, cudaEventDisableTiming) == cudaSuccess &&
               cudaEventCreateWithFlags(&ev_k, cudaEventDisableTiming) == cudaSuccess &&
               cudaEventCreateWithFlags(&ev_d2h, cudaEventDisableTiming) == cudaSuccess;
    }
};
struct DecodeScratch {
    std::mutex mu;
    DecodeSlot slot[2];
};
DecodeScratch g_decode_scratch[64];
}  // namespace

extern "C" int focr_decode_pages(focr_ctx *ctx, const focr_glyph_bank *bank, const uint8_t *pages_host,
                                 size_t page_stride, uint32_t r_w, uint32_t r_h, uint32_t n_pages, uint32_t x_start,
                                 uint32_t y_start, uint32_t width, uint32_t line_height, uint32_t line_advance,
                                 uint32_t max_lines, uint32_t max_cells, uint16_t *glyphs_host, uint32_t *n_cells_host,
                                 uint32_t *line_y_host, uint32_t *n_lines_host)
{
    if (!ctx || !bank || !pages_host || !glyphs_host || !n_cells_host || !line_y_host || !n_lines_host ||
        n_pages == 0 || r_w == 0 || r_h == 0 || line_advance == 0 || max_lines == 0 || max_cells == 0)
        return focr_internal_fail(FOCR_ERR_ARG, "focr_decode_pages: bad argument");
    if (page_stride < (size_t)r_w * r_h) return focr_internal_fail(FOCR_ERR_ARG, "page_stride smaller than a page");
    const size_t strip = (size_t)width * line_height;
    if (strip == 0) return focr_internal_fail(FOCR_ERR_UNSUPPORTED, "empty line rectangle");
    // which kernel: dense tiles when the bank has them and the padded strips fit, else row tasks, else whole glyphs per lane
    const int ox = (int)bank->origin_x;
    const bool tiles = bank->tiles && ox + bank->tile_left >= -FD_PAD && ox + bank->tile_left + bank->tile_wb + 4 <= FD_PAD &&
                       !getenv("FOCR_DECODE_LEGACY");
    const bool row_tasks = bank->tasks && bank->n_glyphs <= FD_MAX_SCORES;
    size_t smem_bytes;
    if (tiles) {
        const size_t sp = (width + 2 * FD_PAD + 3) & ~(size_t)3;
        smem_bytes = FD_WARPS * (((sp * line_height + 16 + 15) & ~(size_t)15) + (size_t)bank->tile_rows * bank->tile_wb);
    } else {
        smem_bytes = ((strip * FD_WARPS + 16 + 15) & ~(size_t)15) + (row_tasks ? FD_WARPS * FD_MAX_SCORES * 4 : 0);
    }
    if (smem_bytes > 200 * 1024 + 16)
        return focr_internal_fail(FOCR_ERR_UNSUPPORTED, "line rectangle too large for shared memory");
    FCU(cudaSetDevice(focr_internal_device(ctx)));
    cudaStream_t streams[3];
    focr_internal_streams(ctx, streams);
    cudaStream_t st = streams[0], st_h2d = streams[1], st_d2h = streams[2];
    // candidate rectangles per page: i = 0.. until the crop height is 0 (main.rs:199-207)
    const uint32_t cand = y_start >= r_h ? 0 : (r_h - y_start + line_advance - 1) / line_advance;
    if (cand > max_lines)
        return focr_internal_fail(FOCR_ERR_ARG, "focr_decode_pages: the page has " + std::to_string(cand) +
                                                    " candidate lines, max_lines is " + std::to_string(max_lines));
    // Only the band the rectangles can touch goes to the device: columns [x_start, x_start+width) of the rows from y_start on
    // (crop_imm's clamping, main.rs:201-203, applied once here); the kernel sees it as pages of `bw` columns with x_start = 0.
    // For BASELINE config 4 that is 608 of 2480 columns: 4x less H2D traffic, which is what bounds this path.
    const uint32_t bx = std::min(x_start, r_w), by = std::min(y_start, r_h);
    const uint32_t bw = std::min(width, r_w - bx);
    if (cand == 0 || bw == 0) {   // no rectangle has a pixel: every strip is empty / "all white" (main.rs:205-211)
        for (uint32_t p = 0; p < n_pages; p++) n_lines_host[p] = 0;
        return FOCR_OK;
    }
    // Device layout.  compact (line_height <= line_advance, the usual case): only the rectangles' rows, line after line --
    // for config 4 that is 12 of every 15 rows, 20 % less H2D again; else the whole band (rows above `by` are never read).
    const bool compact = line_height <= line_advance && !getenv("FOCR_DECODE_BAND");
    const uint32_t n_full = r_h - by >= line_height ? std::min(cand, (r_h - by - line_height) / line_advance + 1) : 0;   // lines not clamped by the page's end
    const uint32_t h_last = cand > n_full ? r_h - (by + n_full * line_advance) : 0;                                      // rows of the clamped last line
    const size_t band_page = compact ? (size_t)bw * line_height * cand : (size_t)bw * r_h;
    const size_t lines_chunk = (size_t)FD_CHUNK * max_lines;
    const bool staged = !focr_internal_host_pinned(pages_host);
    DecodeScratch &sc = g_decode_scratch[focr_internal_device(ctx) & 63];
    std::lock_guard<std::mutex> lock(sc.mu);   // calls on one device share the scratch
    const uint32_t CH = std::min(FD_CHUNK, n_pages);
    for (auto &sl : sc.slot) {
        if (!sl.dev_buf(0, band_page * CH) || !sl.dev_buf(1, (size_t)CH * max_lines * max_cells * 2) ||
            !sl.dev_buf(2, (size_t)CH * max_lines * 4) || !sl.dev_buf(3, 4) ||
            !sl.host_buf(0, (size_t)CH * max_lines * max_cells * 2) || !sl.host_buf(1, (size_t)CH * max_lines * 4 + 4) ||
            (staged && !sl.host_buf(2, band_page * CH)) || !sl.events())
            return focr_internal_fail(FOCR_ERR_NOMEM, "focr_decode_pages: scratch allocation failed");
        if (n_pages <= FD_CHUNK) break;   // a single chunk only needs one slot
    }
    (void)lines_chunk;
    bool cells_error = false;
    // a finished chunk: compact like decode_image -- skip all-white strips, stop at the first empty text (main.rs:205-216)
    auto finish = [&](uint32_t p0, uint32_t nB, DecodeSlot &sl) {
        const uint16_t *g = (const uint16_t *)sl.host[0];
        const uint32_t *c = (const uint32_t *)sl.host[1];
        if (c[(size_t)nB * max_lines]) cells_error = true;
        for (uint32_t q = 0; q < nB; q++) {
            const uint32_t p = p0 + q;
            uint32_t n = 0;
            for (uint32_t i = 0; i < cand; i++) {
                const uint32_t cells = c[(size_t)q * max_lines + i];
                if (cells == 0xFFFFFFFFu) continue;  // skipped strip
                if (cells == 0) break;               // empty text ends the page
                const size_t dst = ((size_t)p * max_lines + n) * max_cells, srcp = ((size_t)q * max_lines + i) * max_cells;
                memcpy(glyphs_host + dst, g + srcp, (size_t)std::min(cells, max_cells) * 2);
                n_cells_host[(size_t)p * max_lines + n] = cells;
                line_y_host[(size_t)p * max_lines + n] = y_start + i * line_advance;
                n++;
            }
            n_lines_host[p] = n;
        }
    };
    const uint32_t n_chunks = (n_pages + CH - 1) / CH;
    for (uint32_t ci = 0; ci < n_chunks; ci++) {
        const uint32_t p0 = ci * CH, nB = std::min(CH, n_pages - p0);
        DecodeSlot &sl = sc.slot[ci & 1];
        if (ci >= 2) {   // the slot's previous chunk: wait for its results, hand them to the caller
            FCU(cudaEventSynchronize(sl.ev_d2h));
            finish((ci - 2) * CH, CH, sl);
        }
        uint8_t *d_band = (uint8_t *)sl.dev[0];
        const uint8_t *src0 = pages_host + (size_t)p0 * page_stride;
        if (compact) {
            uint8_t *stg = staged ? (uint8_t *)sl.host[2] : nullptr;
            if (staged) {   // pageable caller: a few host threads (one page at a time each) gather the rectangles' rows into pinned staging
                auto gather = [&](uint32_t q0, uint32_t step) {
                    for (uint32_t q = q0; q < nB; q += step) {
                        const uint8_t *sp = src0 + q * page_stride + (size_t)by * r_w + bx;
                        uint8_t *dp = stg + q * band_page;
                        for (uint32_t i = 0; i < cand; i++) {
                            const uint32_t rows = i < n_full ? line_height : h_last;
                            for (uint32_t r = 0; r < rows; r++)
                                memcpy(dp + ((size_t)i * line_height + r) * bw, sp + ((size_t)i * line_advance + r) * r_w, bw);
                        }
                    }
                };
                const uint32_t nt = std::min<uint32_t>(nB, (uint32_t)focr_internal_stage_threads());
                focr_internal_parallel_for((int)nt, [&](int t) { gather((uint32_t)t, nt); });
            }
            for (uint32_t q = 0; q < nB && !staged; q++) {
                const uint8_t *sp = src0 + q * page_stride + (size_t)by * r_w + bx;   // first row of line 0 of this page
                if (n_full) {   // one strided 3-D copy: n_full slices of line_height rows, line_advance rows apart in the page
                    cudaMemcpy3DParms cp;
                    memset(&cp, 0, sizeof(cp));
                    cp.srcPtr = make_cudaPitchedPtr((void *)sp, r_w, r_w, line_advance);
                    cp.dstPtr = make_cudaPitchedPtr(d_band + q * band_page, bw, bw, line_height);
                    cp.extent = make_cudaExtent(bw, line_height, n_full);
                    cp.kind = cudaMemcpyHostToDevice;
                    FCU(cudaMemcpy3DAsync(&cp, st_h2d));
                }
                if (h_last)
                    FCU(cudaMemcpy2DAsync(d_band + q * band_page + (size_t)n_full * line_height * bw, bw, sp + (size_t)n_full * line_advance * r_w,
                                          r_w, bw, h_last, cudaMemcpyHostToDevice, st_h2d));
            }
            if (staged) FCU(cudaMemcpyAsync(d_band, stg, band_page * nB, cudaMemcpyHostToDevice, st_h2d));
        } else if (staged) {   // pageable caller: host threads gather the band into pinned staging, then one contiguous copy
            uint8_t *stg = (uint8_t *)sl.host[2];
            for (uint32_t q = 0; q < nB; q++)
                focr_internal_parallel_copy(stg + q * band_page + (size_t)by * bw, bw, src0 + q * page_stride + (size_t)by * r_w + bx, r_w,
                                            bw, r_h - by);
            FCU(cudaMemcpyAsync(d_band + (size_t)by * bw, stg + (size_t)by * bw, band_page * nB - (size_t)by * bw, cudaMemcpyHostToDevice,
                                st_h2d));
        } else if (page_stride == (size_t)r_w * r_h) {   // contiguous pinned pages: ONE strided copy for the whole chunk
            FCU(cudaMemcpy2DAsync(d_band + (size_t)by * bw, bw, src0 + (size_t)by * r_w + bx, r_w, bw, (size_t)nB * r_h - by,
                                  cudaMemcpyHostToDevice, st_h2d));
        } else {
            for (uint32_t q = 0; q < nB; q++)
                FCU(cudaMemcpy2DAsync(d_band + q * band_page + (size_t)by * bw, bw, src0 + q * page_stride + (size_t)by * r_w + bx, r_w,
                                      bw, r_h - by, cudaMemcpyHostToDevice, st_h2d));
        }
        FCU(cudaEventRecord(sl.ev_h2d, st_h2d));
        FCU(cudaStreamWaitEvent(st, sl.ev_h2d, 0));
        FCU(cudaMemsetAsync(sl.dev[2], 0xFF, (size_t)nB * max_lines * 4, st));
        FCU(cudaMemsetAsync(sl.dev[3], 0, 4, st));
        DecodeArgs a;
        a.pages = d_band;
        a.page_stride = band_page;
        a.r_w = bw;
        a.r_h = r_h;
        a.n_pages = nB;
        a.x_start = 0;
        a.y_start = by;
        a.mem_y0 = compact ? 0 : by;
        a.mem_line_rows = compact ? line_height : line_advance;
        a.width = width;
        a.line_height = line_height;
        a.line_advance = line_advance;
        a.max_lines = max_lines;
        a.max_cells = max_cells;
        a.bank.pixels = bank->pixels;
        a.bank.rasters = bank->rasters;
        a.bank.advance_px = bank->advance;
        a.bank.n_glyphs = bank->n_glyphs;
        a.bank.tasks = row_tasks ? bank->tasks : nullptr;
        a.bank.task_off = bank->task_off;
        a.bank.tiles = bank->tiles;
        a.bank.tile_rows = bank->tile_rows;
        a.bank.tile_wb = bank->tile_wb;
        a.bank.tile_left = bank->tile_left;
        a.origin_x = bank->origin_x;
        a.glyphs = (uint16_t *)sl.dev[1];
        a.n_cells = (uint32_t *)sl.dev[2];
        a.error = (unsigned int *)sl.dev[3];
        dim3 grid((max_lines + FD_WARPS - 1) / FD_WARPS, nB);
        void *tm = focr_internal_stage_begin(ctx, FOCR_STAGE_DECODE);
        if (tiles && bank->tile_wb == 16) {
            FCU(cudaFuncSetAttribute(focr_decode_tile_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024 + 16));
            focr_decode_tile_kernel<1><<<grid, FD_WARPS * 32, smem_bytes, st>>>(a);
        } else if (tiles) {
            FCU(cudaFuncSetAttribute(focr_decode_tile_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024 + 16));
            focr_decode_tile_kernel<2><<<grid, FD_WARPS * 32, smem_bytes, st>>>(a);
        } else {
            FCU(cudaFuncSetAttribute(focr_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024 + 16));
            focr_decode_kernel<<<grid, FD_WARPS * 32, smem_bytes, st>>>(a);
        }
        focr_internal_stage_end(tm);
        FCU(cudaGetLastError());
        focr_internal_count_launch(ctx, 1);
        FCU(cudaEventRecord(sl.ev_k, st));
        FCU(cudaStreamWaitEvent(st_d2h, sl.ev_k, 0));
        uint32_t *c = (uint32_t *)sl.host[1];
        FCU(cudaMemcpyAsync(sl.host[0], sl.dev[1], (size_t)nB * max_lines * max_cells * 2, cudaMemcpyDeviceToHost, st_d2h));
        FCU(cudaMemcpyAsync(c, sl.dev[2], (size_t)nB * max_lines * 4, cudaMemcpyDeviceToHost, st_d2h));
        FCU(cudaMemcpyAsync(c + (size_t)nB * max_lines, sl.dev[3], 4, cudaMemcpyDeviceToHost, st_d2h));
        FCU(cudaEventRecord(sl.ev_d2h, st_d2h));
    }
    for (uint32_t ci = (n_chunks >= 2 ? n_chunks - 2 : 0); ci < n_chunks; ci++) {
        DecodeSlot &sl = sc.slot[ci & 1];
        FCU(cudaEventSynchronize(sl.ev_d2h));
        finish(ci * CH, std::min(CH, n_pages - ci * CH), sl);
    }
    if (cells_error) return focr_internal_fail(FOCR_ERR_ARG, "focr_decode_pages: a line needs more than max_cells cells");
    return FOCR_OK;
}

extern "C" int focr_sum_of_squares(focr_ctx *ctx, const uint8_t *xs_host, const uint8_t *ys_host, size_t len,
                                   uint32_t n_pairs, int64_t *out_host)
{
    if (!ctx || !xs_host || !ys_host || !out_host || n_pairs == 0)
        return focr_internal_fail(FOCR_ERR_ARG, "focr_sum_of_squares: bad argument");
    FCU(cudaSetDevice(focr_internal_device(ctx)));
    cudaStream_t st = (cudaStream_t)focr_ctx_stream(ctx);
    uint8_t *dx = nullptr, *dy = nullptr;
    long long *dout = nullptr;
    const size_t bytes = len * n_pairs;
    FCU(cudaMalloc((void **)&dx, bytes ? bytes : 1));
    FCU(cudaMalloc((void **)&dy, bytes ? bytes : 1));
    FCU(cudaMalloc((void **)&dout, n_pairs * 8));
    FCU(cudaMemcpyAsync(dx, xs_host, bytes, cudaMemcpyHostToDevice, st));
    FCU(cudaMemcpyAsync(dy, ys_host, bytes, cudaMemcpyHostToDevice, st));
    sum_of_squares_kernel<<<n_pairs, 256, 0, st>>>(dx, dy, len, dout);
    FCU(cudaGetLastError());
    focr_internal_count_launch(ctx, 1);
    FCU(cudaMemcpyAsync(out_host, dout, n_pairs * 8, cudaMemcpyDeviceToHost, st));
    FCU(cudaStreamSynchronize(st));
    cudaFree(dx);
    cudaFree(dy);
    cudaFree(dout);
    return FOCR_OK;
}
