// stats.cu -- page staging (invert + pitch) and per-window statistics.  HBM-bound kernels.
//
// Replaces, for a whole batch of pages at once:
//   image_to_u8                       ncc.rs:887-892   (255 - p)
//   ncc_sum_table / ncc_sumsqr_table  ncc.rs:938-974   (the SATs are never materialised: a tile-local
//                                                       separable box sum gives the same exact integers)
//   Searcher::prepare_for_size        ncc.rs:263-318   (s_p, patch_rnorm; start/end is only an optimisation
//                                                       of the CPU scan and is not needed)
#include <algorithm>
#include "common.cuh"
#include "kernels.cuh"

namespace focr {

// ---------------------------------------------------------------------------------------------
// invert + re-pitch: gray (tight rows) -> inverted page with a 128-B-multiple pitch, zero padded.
// 16 output bytes per thread; source rows are only byte aligned in general, so the source is read
// through the 4-byte-aligned words that cover the 16 bytes (funnel shifted), which keeps loads
// coalesced and 4 B wide whatever r_w is.
__global__ void __launch_bounds__(256) stage_invert_kernel(const uint8_t *__restrict__ src, size_t src_page_stride,
                                                           size_t src_pitch, uint8_t *__restrict__ dst,
                                                           size_t dst_page_stride, int dst_pitch, int r_w,
                                                           int r_h, int rows_total, int invert)
{
    const int page = blockIdx.z;
    const int y = blockIdx.y;
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (x >= dst_pitch || y >= rows_total) return;
    uint4 o = make_uint4(0, 0, 0, 0);
    if (y < r_h && x < r_w) {
        const uint8_t *row = src + page * src_page_stride + (size_t)y * src_pitch;
        const uintptr_t a = (uintptr_t)(row + x);
        const uint32_t *w = (const uint32_t *)(a & ~(uintptr_t)3);
        const int sh = (int)(a & 3) * 8;
        const int nbytes = min(16, r_w - x);
        // words needed: ceil((a&3 + nbytes)/4); never read a word that starts beyond the row's last byte
        const int nwords = ((int)(a & 3) + nbytes + 3) >> 2;
        uint32_t v[5];
#pragma unroll
        for (int i = 0; i < 5; i++) v[i] = (i < nwords) ? __ldg(w + i) : 0u;
        uint32_t r[4];
#pragma unroll
        for (int i = 0; i < 4; i++) r[i] = __funnelshift_r(v[i], v[i + 1], sh);
        const uint32_t xm = invert ? 0xFFFFFFFFu : 0u;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int valid = nbytes - 4 * i;  // bytes of this word that are inside the row
            const uint32_t mask = valid >= 4 ? 0xFFFFFFFFu : (valid <= 0 ? 0u : ((1u << (8 * valid)) - 1u));
            r[i] = (r[i] ^ xm) & mask;
        }
        o = make_uint4(r[0], r[1], r[2], r[3]);
    }
    *(uint4 *)(dst + page * dst_page_stride + (size_t)y * dst_pitch + x) = o;
}

// ---------------------------------------------------------------------------------------------
// window statistics: exact Sum(p) and Sum(p^2) over every n_w x n_h window.
//   phase 0  coalesced 16-B loads of the (TH+n_h-1) x (TW+n_w-1) input tile into shared memory
//   phase 1  vertical sliding sums per input column            (thread per column)
//   phase 2  exclusive row prefix sums, warp-shuffle scan      (warp per row)
//   phase 3  window sum = E[x+n_w]-E[x]; f64 normaliser; coalesced plane stores
constexpr int ST_TW = 256, ST_TH = 16, ST_THREADS = 256;  // measured: TH 32 -> 0.092, 16 -> 0.066, 8 -> 0.079 ms/page-pair; TW 224 -> 0.115

__device__ __forceinline__ uint32_t warp_excl_scan(uint32_t v, uint32_t &total)
{
    const int lane = threadIdx.x & 31;
    uint32_t s = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t o = __shfl_up_sync(0xffffffffu, s, d);
        if (lane >= d) s += o;
    }
    total = __shfl_sync(0xffffffffu, s, 31);
    return s - v;
}

__global__ void __launch_bounds__(ST_THREADS) window_stats_kernel(StatsArgs a)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const int n_w = a.n_w, n_h = a.n_h;
    const int x0 = blockIdx.x * ST_TW, y0 = blockIdx.y * ST_TH, page = blockIdx.z;
    const int in_w = ST_TW + max(n_w, a.n_w2) - 1, in_h = ST_TH + n_h - 1;
    const int twp = (in_w + 15) & ~15;      // shared row pitch in bytes
    const int vp = twp + 1;                 // prefix row pitch in words (odd-ish stride, +1 for E[in_w])
    uint8_t *pix = smem;
    uint32_t *vs = (uint32_t *)(smem + (((size_t)in_h * twp + 15) & ~(size_t)15));
    uint32_t *vq = vs + ST_TH * vp;

    const uint8_t *pg = a.inv + (size_t)page * a.inv_page_stride;
    // phase 0
    const int vec_per_row = twp >> 4;
    for (int i = threadIdx.x; i < in_h * vec_per_row; i += ST_THREADS) {
        const int r = i / vec_per_row, c = (i - r * vec_per_row) << 4;
        const int gy = y0 + r, gx = x0 + c;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (gy < a.r_h && gx + 16 <= a.pitch) v = __ldg((const uint4 *)(pg + (size_t)gy * a.pitch + gx));
        *(uint4 *)(pix + r * twp + c) = v;
    }
    __syncthreads();
    // phase 1
    for (int c = threadIdx.x; c < in_w; c += ST_THREADS) {
        uint32_t s = 0, q = 0;
        for (int r = 0; r < n_h; r++) {
            uint32_t p = pix[r * twp + c];
            s += p;
            q += p * p;
        }
        for (int y = 0; y < ST_TH; y++) {
            vs[y * vp + c] = s;
            vq[y * vp + c] = q;
            if (y + 1 < ST_TH) {
                uint32_t pa = pix[(y + n_h) * twp + c], pr = pix[y * twp + c];
                s += pa - pr;
                q += pa * pa - pr * pr;
            }
        }
    }
    __syncthreads();
    // phase 2
    {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        const int ch = (in_w + 31) >> 5;  // entries per lane (<= 9 for TW=256, n_w<=32)
        for (int y = warp; y < ST_TH; y += ST_THREADS / 32) {
            uint32_t *rs = vs + y * vp, *rq = vq + y * vp;
            uint32_t ls[10], lq[10];
            uint32_t ts = 0, tq = 0;
#pragma unroll
            for (int k = 0; k < 10; k++) {
                const int c = lane * ch + k;
                const bool ok = (k < ch) && (c < in_w);
                ls[k] = ok ? rs[c] : 0u;
                lq[k] = ok ? rq[c] : 0u;
                ts += ls[k];
                tq += lq[k];
            }
            uint32_t tot_s, tot_q;
            uint32_t bs = warp_excl_scan(ts, tot_s), bq = warp_excl_scan(tq, tot_q);
#pragma unroll
            for (int k = 0; k < 10; k++) {
                const int c = lane * ch + k;
                if ((k < ch) && (c < in_w)) {
                    rs[c] = bs;
                    rq[c] = bq;
                    bs += ls[k];
                    bq += lq[k];
                }
            }
            if (lane == 31) {
                rs[in_w] = tot_s;
                rq[in_w] = tot_q;
            }
        }
    }
    __syncthreads();
    // phase 3
    {
        const int x = threadIdx.x, gx = x0 + x;
        const size_t plane = (size_t)page * a.plane_page_stride;
        const int ny = min(ST_TH, a.r_h - n_h + 1 - y0);   // rows of this tile that hold windows
        auto emit = [&](const int w, const float inv_n_f, uint32_t *sp_out, float *pf_out, const bool extras) {
            if (gx > a.r_w - w) return;
            const double n_d = (double)(w * n_h);
            const uint32_t n_u = (uint32_t)(w * n_h);
            size_t o = plane + (size_t)y0 * a.spitch + gx;
            const uint32_t *e_s = vs + x, *e_q = vq + x;
            for (int y = 0; y < ny; y++, o += a.spitch, e_s += vp, e_q += vp) {
                const uint32_t sp = e_s[w] - e_s[0];
                const uint32_t s2 = e_q[w] - e_q[0];
                if (!a.pack) sp_out[o] = sp;
                if (extras && a.s2p) a.s2p[o] = s2;   // only the SIMT scan and the parity probe read it
                if (a.pack) {   // boxes of at most 256 pixels: n*s2 and sp^2 fit 32 bits, s_p 16, norm_p 11 -> ONE word per window
                    const uint32_t v32 = n_u * s2 - sp * sp;        // n*norm2_p, an exact non-negative integer; 0 <=> constant window
                    float nrm;
                    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(nrm) : "f"((float)v32 * inv_n_f));
                    const uint32_t pfix = v32 == 0u ? 0xFFFFu : min(__float2uint_rn(nrm * 32.f), 0xFFFEu);   // norm_p rounded to 1/32
                    sp_out[o] = sp | (pfix << 16);
                } else {
                    // n*norm2_p = n*s2 - sp^2 is an exact non-negative integer (< 2^45)
                    const unsigned long long vint = (unsigned long long)n_u * s2 - (unsigned long long)sp * sp;
                    // prefilter operand of the tcgen05 epilogue: norm_p = sqrt(vint/n); +inf marks a
                    // constant window (rnorm_p = inf in the reference -> never a hit)
                    // (sqrt.approx: <= 2 ulp, i.e. < 2 units of the screen's 256-unit margin; the exact pass never reads pf)
                    float nrm;
                    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(nrm) : "f"((float)vint * inv_n_f));
                    pf_out[o] = vint == 0ull ? __int_as_float(0x7f800000) : nrm;
                }
                if (extras && a.rn) a.rn[o] = patch_rnorm(sp, s2, n_d);
            }
        };
        emit(n_w, a.inv_n_f, a.sp, a.pf, true);
        if (a.n_w2) emit(a.n_w2, a.inv_n_f2, a.sp2, a.pf2, false);
    }
}

size_t window_stats_smem(int n_w, int n_h)
{
    const int in_w = ST_TW + n_w - 1, in_h = ST_TH + n_h - 1;
    const int twp = (in_w + 15) & ~15;
    return (((size_t)in_h * twp + 15) & ~(size_t)15) + 2 * (size_t)ST_TH * (twp + 1) * 4;
}

cudaError_t launch_stage_invert(const uint8_t *src, size_t src_page_stride, size_t src_pitch, uint8_t *dst,
                                size_t dst_page_stride, int dst_pitch, int r_w, int r_h, int n_pages, int invert,
                                cudaStream_t st)
{
    const int rows_total = r_h + PAGE_PAD_ROWS;
    dim3 grid((dst_pitch / 16 + 255) / 256, rows_total, n_pages);
    stage_invert_kernel<<<grid, 256, 0, st>>>(src, src_page_stride, src_pitch, dst, dst_page_stride, dst_pitch,
                                              r_w, r_h, rows_total, invert);
    return cudaGetLastError();
}

cudaError_t launch_window_stats(const StatsArgs &a, int n_pages, cudaStream_t st)
{
    const size_t smem = window_stats_smem(std::max(a.n_w, a.n_w2), a.n_h);
    {   // the opt-in is per DEVICE (a process may hold contexts on several GPUs): set it on every launch, it is cheap
        cudaError_t e = cudaFuncSetAttribute(window_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)window_stats_smem(MAX_TPL_W, MAX_TPL_H));
        if (e != cudaSuccess) return e;
    }
    const int xs = a.r_w - std::min(a.n_w, a.n_w2 ? a.n_w2 : a.n_w) + 1, ys = a.r_h - a.n_h + 1;
    dim3 grid((xs + ST_TW - 1) / ST_TW, (ys + ST_TH - 1) / ST_TH, n_pages);
    window_stats_kernel<<<grid, ST_THREADS, smem, st>>>(a);
    return cudaGetLastError();
}

}  // namespace focr
