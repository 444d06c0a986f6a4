// finalize.cu -- turn the unordered hit list of a scan into the reference's output contract:
// per (page, template) the hits in (y, x) raster order, truncated at n_out exactly where
// ncc_8_u8 / ncc_16_u8 return early (ncc.cpp:225-227, 242-244, 371-373, 388-390).
//
//   1. row_cut   : per (page, t) scan the per-row hit counts; y_cut = the row in which the
//                  cumulative count reaches n_out (all rows if it never does).
//   2. select    : keep only hits with y <= y_cut: those above the cut row (at most n_out-1) in one list, those
//                  OF the cut row (at most one per x) in a second one.
//   3. sort_emit : per (page, t) bitonic sorts of the 64-bit keys (y:16 | x:16 | sim bits:32) in shared memory --
//                  the rows above the cut (a power of two <= n_out keys, not n_out + a row's worth rounded up) and
//                  the cut row (a handful) -- then the first min(count, n_out) as Match{u16 x, u16 y, f32}.
#include "common.cuh"
#include "kernels.cuh"

namespace focr {

__global__ void __launch_bounds__(256) row_cut_kernel(FinalizeArgs a)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int PT = a.n_pages * a.T;
    if (warp >= PT) return;
    const unsigned int *rc = a.rowcount + (size_t)warp * a.r_h;
    const int ch = (a.r_h + 31) / 32;
    const int b = lane * ch, e = min(b + ch, (int)a.r_h);
    uint32_t s = 0;
    for (int y = b; y < e; y++) s += rc[y];
    uint32_t incl = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    uint32_t run = incl - s;  // hits in rows before this lane's chunk
    uint32_t cut = 0xFFFFFFFFu;
    if (run < a.n_out && incl >= a.n_out) {  // the n_out-th hit lies in this lane's chunk
        for (int y = b; y < e; y++) {
            run += rc[y];
            if (run >= a.n_out) {
                cut = y;
                break;
            }
        }
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) cut = min(cut, __shfl_xor_sync(0xffffffffu, cut, d));
    if (lane == 0) a.y_cut[warp] = cut;
}

__global__ void __launch_bounds__(256) select_kernel(FinalizeArgs a)
{
    const unsigned n = min(*a.hit_count, a.hit_cap);
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const Hit h = a.hits[i];
        const uint32_t pt = h.page * a.T + h.t;
        const uint32_t y = h.yx >> 16, yc = a.y_cut[pt];
        if (y > yc) continue;
        const unsigned long long key = ((unsigned long long)h.yx << 32) | (unsigned long long)__float_as_uint(h.sim);
        if (y < yc) {   // rows above the cut row: fewer than n_out hits by construction (all hits when there is no cut row)
            const unsigned slot = atomicAdd(a.sel_count + pt, 1u);
            if (slot < a.n_out)
                a.sel[(size_t)pt * a.sel_cap + slot] = key;
            else
                atomicExch(a.overflow, 1u);
        } else {        // the cut row itself: at most one hit per x
            const unsigned slot = atomicAdd(a.sel_count + (size_t)a.n_pages * a.T + pt, 1u);
            if (a.n_out + slot < a.sel_cap)
                a.sel[(size_t)pt * a.sel_cap + a.n_out + slot] = key;
            else
                atomicExch(a.overflow, 1u);
        }
    }
}

// in-place ascending bitonic sort of keys[0 .. m), m a power of two (all threads of the block)
__device__ __forceinline__ void bitonic_sort(unsigned long long *keys, uint32_t m)
{
    for (uint32_t k = 2; k <= m; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) {
                const uint32_t l = i ^ j;
                if (l > i) {
                    const unsigned long long x = keys[i], y = keys[l];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) {
                        keys[i] = y;
                        keys[l] = x;
                    }
                }
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(256) sort_emit_kernel(FinalizeArgs a)
{
    extern __shared__ __align__(16) unsigned long long keys[];
    const uint32_t pt = blockIdx.x;
    const uint32_t c1 = min(a.sel_count[pt], a.n_out);                                   // above the cut row
    const uint32_t c2 = min(a.sel_count[(size_t)a.n_pages * a.T + pt], a.sel_cap - a.n_out);   // the cut row
    uint32_t m1 = 1, m2 = 1;
    while (m1 < c1) m1 <<= 1;
    while (m2 < c2) m2 <<= 1;
    const unsigned long long *src = a.sel + (size_t)pt * a.sel_cap;
    unsigned long long *keys2 = keys + m1;
    for (uint32_t i = threadIdx.x; i < m1; i += blockDim.x) keys[i] = i < c1 ? src[i] : ~0ull;
    for (uint32_t i = threadIdx.x; i < m2; i += blockDim.x) keys2[i] = i < c2 ? src[a.n_out + i] : ~0ull;
    __syncthreads();
    bitonic_sort(keys, m1);
    if (c2 > 1) bitonic_sort(keys2, m2);
    // every key of the cut row is larger than every key above it: the concatenation is the (y, x) raster order
    const uint32_t n = min(c1 + c2, a.n_out);
    focr_match *out = a.out + (size_t)pt * a.n_out;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const unsigned long long k = i < c1 ? keys[i] : keys2[i - c1];
        focr_match mt;
        mt.y = (uint16_t)(k >> 48);
        mt.x = (uint16_t)(k >> 32);
        mt.similarity = __uint_as_float((uint32_t)k);
        out[i] = mt;
    }
    if (threadIdx.x == 0) a.counts[pt] = n;
}

size_t finalize_sel_cap(uint32_t r_w, uint32_t n_out)
{
    // worst case kept by `select`: n_out-1 hits before the cut row + every x of the cut row
    size_t need = (size_t)n_out + r_w;
    size_t cap = 1024;
    while (cap < need) cap <<= 1;
    return cap <= 16384 ? cap : 0;  // 16384 keys = 128 KB of shared memory for the sort
}

cudaError_t launch_finalize(const FinalizeArgs &a, cudaStream_t st, int *n_launches)
{
    const int PT = a.n_pages * a.T;
    row_cut_kernel<<<(PT * 32 + 255) / 256, 256, 0, st>>>(a);
    select_kernel<<<148 * 8, 256, 0, st>>>(a);
    cudaError_t e = cudaFuncSetAttribute(sort_emit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (4096 + 16384) * 8);
    if (e != cudaSuccess) return e;
    // shared memory: the two sorted ranges, each rounded up to a power of two (<= pow2(n_out) + pow2(sel_cap - n_out) keys)
    size_t m1 = 1, m2 = 1;
    while (m1 < a.n_out) m1 <<= 1;
    while (m2 < a.sel_cap - a.n_out) m2 <<= 1;
    sort_emit_kernel<<<PT, 256, (m1 + m2) * 8, st>>>(a);
    if (n_launches) *n_launches += 3;
    return cudaGetLastError();
}

}  // namespace focr
