// scan_simt.cu -- the NCC template scan on the CUDA cores (dp4a), exact end to end.
//
// Replaces ncc_8_u8 / ncc_16_u8 (ncc.cpp:48-251, 253-396) for a whole batch of (page, template)
// pairs per launch.  This is the correctness anchor and the path for shapes the tcgen05 kernel
// (scan_tc.cu) does not cover; every window goes through the reference's f64 epilogue.
//
//   tile       128 consecutive x  x  RY rows  x  up to TCH templates of one box size
//   numerator  exact u32: u8 x u8 products via __dp4a on byte-shifted window words held in registers,
//              template rows broadcast from shared memory (ncc.cpp:108-166: same integers, any order)
//   epilogue   ncc_exact() == ncc.cpp:212-220; hits appended unordered, ordered later (finalize.cu)
#include "common.cuh"
#include "kernels.cuh"

namespace focr {

constexpr int SS_THREADS = 128, SS_RY = 8, SS_TCH_MAX = 32;

template <int NPW>  // padded template row width in 32-bit words: 4 (<=16 px) or 8 (<=32 px)
__global__ void __launch_bounds__(SS_THREADS) scan_simt_kernel(ScanArgs a, int tch, int n_chunks)
{
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr int NP = NPW * 4;
    constexpr int PW = 128 + NP;  // tile pitch in bytes (multiple of 16)
    const int n_w = a.cls.n_w, n_h = a.cls.n_h;
    const int page = blockIdx.z / n_chunks, chunk = blockIdx.z - page * n_chunks;
    const int x0 = blockIdx.x * 128, y0 = blockIdx.y * SS_RY;
    const int rows = SS_RY + n_h - 1;
    uint8_t *tile = smem;
    uint8_t *tpl_s = smem + rows * PW;

    const uint8_t *pg = a.inv + (size_t)page * a.inv_page_stride;
    for (int i = threadIdx.x; i < rows * (PW / 16); i += SS_THREADS) {
        const int r = i / (PW / 16), c = (i - r * (PW / 16)) * 16;
        const int gy = y0 + r, gx = x0 + c;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (gy < a.r_h + PAGE_PAD_ROWS && gx + 16 <= a.pitch) v = __ldg((const uint4 *)(pg + (size_t)gy * a.pitch + gx));
        *(uint4 *)(tile + r * PW + c) = v;
    }
    const int t_begin = chunk * tch;
    const int t_cnt = min(tch, (int)a.cls.n_tpl - t_begin);
    const int t_pad = (t_cnt + 3) & ~3;
    {
        const uint4 *src = (const uint4 *)(a.cls.rows + (size_t)t_begin * n_h * NP);
        const int nvec = t_cnt * n_h * NP / 16, nvec_pad = t_pad * n_h * NP / 16;
        for (int i = threadIdx.x; i < nvec_pad; i += SS_THREADS)
            ((uint4 *)tpl_s)[i] = i < nvec ? __ldg(src + i) : make_uint4(0, 0, 0, 0);
    }
    __syncthreads();

    const int tx = threadIdx.x, gx = x0 + tx;
    const int w0 = tx >> 2, sh = (tx & 3) * 8;
    const bool x_ok = gx >= 1 && gx <= a.r_w - n_w;  // ncc.rs:281: the search starts at x = 1
    const size_t plane = (size_t)page * a.plane_page_stride;

    for (int ry = 0; ry < SS_RY; ry++) {
        const int y = y0 + ry;
        const bool ok = x_ok && y >= 1 && y <= a.r_h - n_h;  // ncc.cpp:98: y starts at 1
        uint32_t s_p = 0;
        double rn_p = 0.0;
        if (ok) {
            s_p = a.sp[plane + (size_t)y * a.spitch + gx];
            rn_p = a.rn[plane + (size_t)y * a.spitch + gx];
        }
        for (int t0 = 0; t0 < t_pad; t0 += 4) {
            uint32_t acc[4] = {0, 0, 0, 0};
            for (int ny = 0; ny < n_h; ny++) {
                const uint32_t *trow = (const uint32_t *)(tile + (ry + ny) * PW) + w0;
                uint32_t wv[NPW + 1], v[NPW];
#pragma unroll
                for (int i = 0; i <= NPW; i++) wv[i] = trow[i];
#pragma unroll
                for (int i = 0; i < NPW; i++) v[i] = __funnelshift_r(wv[i], wv[i + 1], sh);
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint4 *tr = (const uint4 *)(tpl_s + ((size_t)(t0 + k) * n_h + ny) * NP);
#pragma unroll
                    for (int q = 0; q < NPW / 4; q++) {
                        const uint4 tw = tr[q];
                        acc[k] = __dp4a(v[4 * q + 0], tw.x, acc[k]);
                        acc[k] = __dp4a(v[4 * q + 1], tw.y, acc[k]);
                        acc[k] = __dp4a(v[4 * q + 2], tw.z, acc[k]);
                        acc[k] = __dp4a(v[4 * q + 3], tw.w, acc[k]);
                    }
                }
            }
            if (!ok) continue;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (t0 + k >= t_cnt) break;
                const uint32_t t = a.cls.tpl_index[t_begin + t0 + k];
                if (a.acc_out) a.acc_out[(size_t)y * a.r_w + gx] = acc[k];
                const TplInfo ti = a.tpl[t];
                float sim;
                if (ncc_exact(acc[k], s_p, rn_p, ti.s_n, ti.n_recip, ti.rnorm_n, a.thr_d, &sim)) {
                    const unsigned slot = atomicAdd(a.sink.hit_count, 1u);
                    if (slot < a.sink.hit_cap) {
                        Hit h;
                        h.t = t;
                        h.yx = ((uint32_t)y << 16) | (uint32_t)gx;
                        h.sim = sim;
                        h.page = page;
                        a.sink.hits[slot] = h;
                    }
                    atomicAdd(a.sink.rowcount + ((size_t)page * a.sink.T + t) * a.sink.r_h + y, 1u);
                }
            }
        }
    }
}

cudaError_t launch_scan_simt(const ScanArgs &a, int n_pages, cudaStream_t st, int *n_launches)
{
    const int n_w = a.cls.n_w, n_h = a.cls.n_h, np = a.cls.np;
    const int npw = np / 4;
    if (npw != 4 && npw != 8) return cudaErrorInvalidValue;
    const int rows = SS_RY + n_h - 1;
    const size_t tile_bytes = (size_t)rows * (128 + np);
    int tch = min(SS_TCH_MAX, (int)((a.cls.n_tpl + 3) & ~3u));
    while (tch > 4 && tile_bytes + (size_t)tch * n_h * np > 96 * 1024) tch -= 4;
    const size_t smem = tile_bytes + (size_t)tch * n_h * np;
    const int n_chunks = (a.cls.n_tpl + tch - 1) / tch;
    const int xs = a.r_w - n_w + 1, ys = a.r_h - n_h + 1;
    dim3 grid((xs + 127) / 128, (ys + SS_RY - 1) / SS_RY, n_pages * n_chunks);
    if (grid.z > 65535) return cudaErrorInvalidValue;
    cudaError_t e;
    if (npw == 4) {
        e = cudaFuncSetAttribute(scan_simt_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        if (e != cudaSuccess) return e;
        scan_simt_kernel<4><<<grid, SS_THREADS, smem, st>>>(a, tch, n_chunks);
    } else {
        e = cudaFuncSetAttribute(scan_simt_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        if (e != cudaSuccess) return e;
        scan_simt_kernel<8><<<grid, SS_THREADS, smem, st>>>(a, tch, n_chunks);
    }
    if (n_launches) (*n_launches)++;
    return cudaGetLastError();
}

}  // namespace focr
