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
#include <mutex>
#include <string>
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
};

constexpr uint32_t FD_MAX_SCORES = 256;   // glyphs per bank the row-task path keeps scores for (per warp, shared memory)

constexpr int FD_WARPS = 4;  // lines per block

struct DecodeArgs {
    const uint8_t *pages;  // gray, tight rows
    size_t page_stride;
    uint32_t r_w, r_h, n_pages;
    uint32_t x_start, y_start, width, line_height, line_advance;
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
    const uint8_t *src = a.pages + (size_t)page * a.page_stride + (size_t)ys * a.r_w + xs;
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

struct focr_glyph_bank {
    focr_ctx *ctx;
    uint8_t *pixels;
    focr_glyph_raster *rasters;
    float *advance;
    uint32_t n_glyphs;
    float origin_x;
    uint4 *tasks;          // row tasks per phase (GlyphBankDev), nullptr when a bitmap is too large for the packing
    uint32_t *task_off;
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
    // the kernel accumulates Sum g*(g - 2*ref) in 32-bit integers (the reference's sum_of_squares is i64, main.rs:510-516):
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
    delete b;
}

// Grow-only scratch per device (device buffers + pinned staging for the results): a cudaMalloc / cudaFree pair per buffer
// and call cost more than the decode kernel (12 ms of a 31 ms call for 32 pages), and cudaFree synchronises the device.
namespace {
struct DecodeScratch {
    std::mutex mu;
    void *dev[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t dev_cap[4] = {0, 0, 0, 0};
    void *host[2] = {nullptr, nullptr};
    size_t host_cap[2] = {0, 0};
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
    const bool row_tasks = bank->tasks && bank->n_glyphs <= FD_MAX_SCORES;
    const size_t smem_bytes = ((strip * FD_WARPS + 16 + 15) & ~(size_t)15) + (row_tasks ? FD_WARPS * FD_MAX_SCORES * 4 : 0);
    if (smem_bytes > 200 * 1024 + 16 || strip == 0)
        return focr_internal_fail(FOCR_ERR_UNSUPPORTED, "line rectangle too large for shared memory");
    FCU(cudaSetDevice(focr_internal_device(ctx)));
    cudaStream_t st = (cudaStream_t)focr_ctx_stream(ctx);
    // candidate rectangles per page: i = 0.. until the crop height is 0 (main.rs:199-207)
    const uint32_t cand = y_start >= r_h ? 0 : (r_h - y_start + line_advance - 1) / line_advance;
    if (cand > max_lines)
        return focr_internal_fail(FOCR_ERR_ARG, "focr_decode_pages: the page has " + std::to_string(cand) +
                                                    " candidate lines, max_lines is " + std::to_string(max_lines));
    uint8_t *d_pages = nullptr;
    uint16_t *d_glyphs = nullptr;
    uint32_t *d_cells = nullptr;
    unsigned int *d_err = nullptr;
    const size_t n_lines_tot = (size_t)n_pages * max_lines;
    // Only the band the rectangles can touch goes to the device: columns [x_start, x_start+width) and rows from y_start
    // (crop_imm's clamping, main.rs:201-203, applied once here); the kernel sees it as a page of its own with x_start =
    // y_start = 0.  For BASELINE config 4 that is 608 of 2480 columns: 4x less H2D traffic, which is what bounds this path.
    const uint32_t bx = std::min(x_start, r_w), by = std::min(y_start, r_h);
    const uint32_t bw = std::min(width, r_w - bx), bh = r_h - by;
    const size_t band_bytes = (size_t)bw * bh;
    DecodeScratch &sc = g_decode_scratch[focr_internal_device(ctx) & 63];
    std::lock_guard<std::mutex> lock(sc.mu);   // calls on one device share the scratch
    d_pages = (uint8_t *)sc.dev_buf(0, std::max<size_t>(band_bytes * n_pages, 1));
    d_glyphs = (uint16_t *)sc.dev_buf(1, n_lines_tot * max_cells * 2);
    d_cells = (uint32_t *)sc.dev_buf(2, n_lines_tot * 4);
    d_err = (unsigned int *)sc.dev_buf(3, 4);
    uint16_t *g = (uint16_t *)sc.host_buf(0, n_lines_tot * max_cells * 2);
    uint32_t *c = (uint32_t *)sc.host_buf(1, n_lines_tot * 4 + 4);
    if (!d_pages || !d_glyphs || !d_cells || !d_err || !g || !c)
        return focr_internal_fail(FOCR_ERR_NOMEM, "focr_decode_pages: scratch allocation failed");
    FCU(cudaMemsetAsync(d_cells, 0xFF, n_lines_tot * 4, st));
    FCU(cudaMemsetAsync(d_err, 0, 4, st));
    if (band_bytes)
        for (uint32_t p = 0; p < n_pages; p++)
            FCU(cudaMemcpy2DAsync(d_pages + p * band_bytes, bw, pages_host + p * page_stride + (size_t)by * r_w + bx, r_w, bw,
                                  bh, cudaMemcpyHostToDevice, st));
    DecodeArgs a;
    a.pages = d_pages;
    a.page_stride = band_bytes;
    a.r_w = bw;
    a.r_h = bh;
    a.n_pages = n_pages;
    a.x_start = 0;
    a.y_start = 0;
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
    a.origin_x = bank->origin_x;
    a.glyphs = d_glyphs;
    a.n_cells = d_cells;
    a.error = d_err;
    FCU(cudaFuncSetAttribute(focr_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024 + 16));
    dim3 grid((max_lines + FD_WARPS - 1) / FD_WARPS, n_pages);
    focr_decode_kernel<<<grid, FD_WARPS * 32, smem_bytes, st>>>(a);
    FCU(cudaGetLastError());
    focr_internal_count_launch(ctx, 1);
    FCU(cudaMemcpyAsync(g, d_glyphs, n_lines_tot * max_cells * 2, cudaMemcpyDeviceToHost, st));
    FCU(cudaMemcpyAsync(c, d_cells, n_lines_tot * 4, cudaMemcpyDeviceToHost, st));
    FCU(cudaMemcpyAsync(c + n_lines_tot, d_err, 4, cudaMemcpyDeviceToHost, st));
    FCU(cudaStreamSynchronize(st));
    const unsigned int err = c[n_lines_tot];
    if (err) return focr_internal_fail(FOCR_ERR_ARG, "focr_decode_pages: a line needs more than max_cells cells");
    // compact like decode_image: skip all-white strips, stop at the first empty text (main.rs:205-216)
    for (uint32_t p = 0; p < n_pages; p++) {
        uint32_t n = 0;
        for (uint32_t i = 0; i < cand; i++) {
            const uint32_t cells = c[(size_t)p * max_lines + i];
            if (cells == 0xFFFFFFFFu) continue;  // skipped strip
            if (cells == 0) break;               // empty text ends the page
            const size_t dst = ((size_t)p * max_lines + n) * max_cells, srcp = ((size_t)p * max_lines + i) * max_cells;
            for (uint32_t k = 0; k < cells; k++) glyphs_host[dst + k] = g[srcp + k];
            n_cells_host[(size_t)p * max_lines + n] = cells;
            line_y_host[(size_t)p * max_lines + n] = y_start + i * line_advance;
            n++;
        }
        n_lines_host[p] = n;
    }
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
