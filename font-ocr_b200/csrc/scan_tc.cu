// scan_tc.cu -- the NCC correlation as a tcgen05 integer-MMA GEMM (sm_100a), with an exact epilogue.
//
// Replaces the hot loops of ncc_8_u8 / ncc_16_u8 (ncc.cpp:98-248, 302-393) for ALL templates of one
// box size at once.  For an output row y of a 128-window strip,
//
//     D[m, t] = sum_{ny, j} page[y+ny][x0+m+j] * tpl[t][ny][j]          (exact, u8 x u8 -> s32)
//
// is a GEMM with M = 128 windows, N = NB templates (<= 256) and K = 16*n_h (32*n_h for boxes wider
// than 16): A is the Toeplitz (im2col) expansion of the page rows, B the template bank.
//
// * A cannot be described by a UMMA shared-memory descriptor directly (rows of a core matrix are 16 B
//   apart, windows are 1 B apart), so every page row is expanded ONCE into a 128 x 16 B block in shared
//   memory and then reused by all n_h vertical taps and all NB templates: the k-th MMA of a row simply
//   points its descriptor at ring slots (y+2k, y+2k+1) through the leading-dimension byte offset.
// * Raw page rows arrive by TMA bulk copies (cp.async.bulk, mbarrier complete_tx).
// * B (templates, K-major, no swizzle) is loaded once per launch by one bulk copy.
// * Accumulators live in TMEM (double/quad buffered); one elected thread issues tcgen05.mma kind::i8.
// * Epilogue (8 warps): tcgen05.ld -> fp32 prefilter with a proven safety margin -> the reference's
//   exact f64 normalisation (ncc.cpp:212-220) only for the survivors -> warp-aggregated atomic append.
//   The integer numerators are exact, the decision and the f32 score are bit-identical to the CPU.
//
// Warp roles (768 threads, one CTA per SM, persistent over (page, x-strip, y-segment) items):
//   warp 0      TMA producer of raw page rows          warp 1   MMA issuer
//   warp 2      TMEM allocator                         warp 3   idle
//   warps 4-7   Toeplitz expansion (thread = window)   warps 8-23  epilogue (4 per TMEM lane quarter)
#include <cooperative_groups.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "scan_tc.cuh"

namespace cg = cooperative_groups;

#ifndef FOCR_SCALAR_FFMA
#define FOCR_SCALAR_FFMA 1  // 1: four scalar FFMA per column pair; 0: two packed FFMA2 (measured slower: FFMA2 only issues on the fma-heavy pipe)
#endif

namespace focr {

constexpr int TC_THREADS = 768;      // 4 service + 4 expansion + 16 epilogue warps
constexpr int TC_EPI_GROUPS = 4;     // epilogue warps per TMEM lane quarter (= per SM sub-partition)
constexpr int TC_G = 4;               // page rows per pipeline group: one mbarrier handshake per 4 rows
constexpr int TC_RAW_GROUPS = 4;      // raw page-row ring (TMA destination): 4 groups x 4 rows x 160 B
constexpr int TC_RAW_SLOTS = TC_RAW_GROUPS * TC_G;
constexpr int TC_RAW_BYTES = 160;
constexpr int TC_LOOK_GROUPS = 2;     // expanded row groups the producer side may run ahead of the MMA
constexpr int TC_RING_MAX = 12;       // max ring groups
constexpr int TC_MAX_BUF = 8;         // TMEM accumulator buffers
constexpr int TC_YSEG = 128;          // output rows per work item
constexpr int TC_ST_DEPTH = 8;        // rows of window statistics each epilogue warp streams ahead (cp.async ring)
constexpr size_t TC_SMEM_BUDGET = 200 * 1024;

struct TcParams {
    const uint8_t *inv;
    size_t inv_page_stride;
    int pitch, r_w, r_h;
    int n_w, n_h, np;
    int n_hp;          // page rows an output row needs (n_h rounded up to 2 when np == 16)
    int ksteps;        // tcgen05.mma per output row
    int nb;            // templates per launch (multiple of 16; the B tile in shared memory)
    int nsub;          // the N dimension is issued as nsub MMAs of n_mma columns each (finer TMEM buffering)
    int n_mma;         // N of one tcgen05.mma (multiple of 16, <= 128 when nsub > 1)
    int nunits;        // 32-column epilogue units per accumulator buffer
    int nbs;           // TMEM column stride between accumulator buffers (n_mma rounded up to 32)
    int nbuf;          // accumulator buffers
    int ring;          // ring slots (rows) = ring_groups * 4
    int ring_groups;
    int row_pitch;     // bytes per expanded row slot
    int n_entries;     // 16-byte entries per expanded row (128, or 144 for np == 32)
    const uint8_t *btile;    // [2*ksteps][nb][16]
    uint32_t btile_bytes;
    uint32_t col_base;       // this launch's first column within the class (N-block * nb)
    const uint32_t *sp;
    const float *pf;
    int spitch;
    size_t plane_page_stride;
    Hit *cands;                // candidate list (prefilter survivors): {column, y<<16|x, acc bits, page}
    uint32_t cand_cap;
    unsigned int *cand_count;
    int n_pages, n_xstrips, n_ysegs;
    int dbg_mode;       // timing experiments only (env FOCR_TC_DBG): 1 = epilogue skips the TMEM reads, 2 = reads but no filter
    uint32_t *dbg_acc;  // parity probe: raw numerators of column dbg_col, [y*r_w+x] (NULL in production)
    int dbg_col;
    // prefilter constants per column, negated: {-a', -b'}
    float2 cst[256];
};

// ---------------------------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
// Wait for a phase of an mbarrier.  SUSPEND_NS > 0 passes a suspend-time hint to try_wait: the hardware
// parks the thread (no issue slots taken from the epilogue warps) and wakes it AS SOON AS the phase
// completes.  (__nanosleep back-off was measured to add ~1 us to every handoff: its granularity is far
// coarser than the requested 20-200 ns.)
template <int SUSPEND_NS = 0>
__device__ __forceinline__ void mbar_wait_addr(uint32_t addr, uint32_t parity)
{
    uint32_t done;
    do {
        if (SUSPEND_NS > 0) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(addr), "r"(parity), "r"((uint32_t)SUSPEND_NS)
                : "memory");
        } else {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(addr), "r"(parity)
                : "memory");
        }
    } while (!done);
}
template <int SUSPEND_NS = 0>
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    mbar_wait_addr<SUSPEND_NS>(smem_u32(bar), parity);
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tc_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_commit_addr(uint32_t bar_addr)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_addr) : "memory");
}
__device__ __forceinline__ void tc_mma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// UMMA shared-memory descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor):
//   core matrix = 8 rows x 16 B, rows 16 B apart; 8-row groups SBO apart; the two 16-B K chunks LBO apart
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) |
           (1ull << 46);  // version = 1 (Blackwell), base_offset = 0, lbo_mode = 0, layout = SWIZZLE_NONE
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
// tcgen05.wait::ld that is ordered against the USES of v through register dependencies instead of a
// memory clobber, so that independent shared-memory loads (the prefilter constants) can be hoisted above it
__device__ __forceinline__ void tc_wait_ld16(uint32_t (&v)[16])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                   "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]));
}
__device__ __forceinline__ void tc_wait_ld32(uint32_t (&v)[32])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                   "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
                   "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
                   "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31]));
}
__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gmem_src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ uint32_t ldg_now_u32(const uint32_t *p)
{
    uint32_t v;  // volatile: issue the load HERE (the compiler would otherwise sink a prefetch to its use)
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float ldg_now_f32(const float *p)
{
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
// packed fp32x2 FMA (sm_100 FFMA2): two columns per instruction
__device__ __forceinline__ unsigned long long pack2(float lo, float hi)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c)
{
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint32_t tc_ld1(uint32_t taddr)
{
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
    return v;
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------- work items
struct Item {
    int page, x0, ys0, ys1;
};
__device__ __forceinline__ bool get_item(const TcParams &p, int idx, Item &it)
{
    const int per_page = p.n_xstrips * p.n_ysegs;
    if (idx >= p.n_pages * per_page) return false;
    it.page = idx / per_page;
    const int r = idx - it.page * per_page;
    const int ys = r / p.n_xstrips, xs = r - ys * p.n_xstrips;
    it.x0 = xs * 128;
    it.ys0 = 1 + ys * TC_YSEG;  // ncc.cpp:98: the scan starts at y = 1
    it.ys1 = min(it.ys0 + TC_YSEG, p.r_h - p.n_h + 1);
    return true;
}

// A prefilter survivor: the raw numerator goes to the candidate list; cand_exact_kernel applies the
// reference's exact f64 arithmetic afterwards, so the MMA kernel carries no double-precision code.
__device__ __forceinline__ void push_candidate(const TcParams &p, uint32_t acc, uint32_t col, int page, int gx, int y)
{
    // warp-aggregated append: one atomic per group of lanes that have a candidate in this column
    auto g = cg::coalesced_threads();
    unsigned base = 0;
    if (g.thread_rank() == 0) base = atomicAdd(p.cand_count, (unsigned)g.size());
    base = g.shfl(base, 0);
    const unsigned slot = base + g.thread_rank();
    if (slot < p.cand_cap) {
        Hit h;
        h.t = p.col_base + col;  // column within the class; cand_exact_kernel maps it to the bank index
        h.yx = ((uint32_t)y << 16) | (uint32_t)gx;
        h.sim = __uint_as_float(acc);
        h.page = page;
        p.cands[slot] = h;
    }
}

// one 32-column unit of the prefilter: a 0 bit in the result marks a candidate column, where the bit
// is the SIGN of   d_j = acc_j - b'_j*S - a'_j*P      (a', b' shrunk by 2^-12: DESIGN.md "prefilter margin").
// cs[i] = {-b'_2i, -b'_2i+1, -a'_2i, -a'_2i+1}: one LDS.128 (broadcast) and two FFMA2 per column pair.
// Bits: even column j -> bit 31 - j/2, odd column j -> bit 15 - (j-1)/2.
__device__ __forceinline__ uint32_t prefilter_unit(const float4 *__restrict__ cs, const uint32_t (&v)[32],
                                                   unsigned long long SS, unsigned long long PP)
{
    uint32_t m0 = 0, m1 = 0;
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
        const float4 c = cs[j >> 1];
        unsigned long long d = pack2(__int2float_rn((int)v[j]), __int2float_rn((int)v[j + 1]));
        d = ffma2(pack2(c.x, c.y), SS, d);
        d = ffma2(pack2(c.z, c.w), PP, d);
        uint32_t d0, d1;
        asm("mov.b64 {%0,%1}, %2;" : "=r"(d0), "=r"(d1) : "l"(d));
        m0 = __funnelshift_l(d0, m0, 1);
        m1 = __funnelshift_l(d1, m1, 1);
    }
    return __byte_perm(m1, m0, 0x5410);  // (m0 << 16) | (m1 & 0xFFFF) in one PRMT
}

// Fast screen of a 32-column unit: max_j d_j with one 3-input max (FMNMX3) per column pair instead of
// a per-column sign mask.  Almost every unit has no candidate at all (max < 0), so the per-column mask
// (prefilter_unit) is only computed for the few units where some lane's maximum is >= 0.
__device__ __forceinline__ float prefilter_unit_max(const float4 *__restrict__ cs, const uint32_t (&v)[32],
                                                    unsigned long long SS, unsigned long long PP)
{
    float m = __int_as_float(0xff800000);  // -inf
#if FOCR_SCALAR_FFMA
    float S1, S1b, P1, P1b;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(S1), "=f"(S1b) : "l"(SS));
    asm("mov.b64 {%0,%1}, %2;" : "=f"(P1), "=f"(P1b) : "l"(PP));
#endif
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
        const float4 c = cs[j >> 1];
#if FOCR_SCALAR_FFMA
        float d0 = __fmaf_rn(c.x, S1, __int2float_rn((int)v[j])), d1 = __fmaf_rn(c.y, S1, __int2float_rn((int)v[j + 1]));
        d0 = __fmaf_rn(c.z, P1, d0);
        d1 = __fmaf_rn(c.w, P1, d1);
#else
        unsigned long long d = pack2(__int2float_rn((int)v[j]), __int2float_rn((int)v[j + 1]));
        d = ffma2(pack2(c.x, c.y), SS, d);
        d = ffma2(pack2(c.z, c.w), PP, d);
        float d0, d1;
        asm("mov.b64 {%0,%1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
#endif
        asm("max.f32 %0, %1, %2, %3;" : "=f"(m) : "f"(m), "f"(d0), "f"(d1));
    }
    return m;
}

// the rare path (kept out of line, re-reads the unit from TMEM so nothing has to be passed in memory)
__device__ __noinline__ void candidates_of_unit(const TcParams &p, const float4 *cst_unit, uint32_t taddr, int col,
                                                unsigned long long SS, unsigned long long PP, bool valid, int page,
                                                int gx, int y)
{
    uint32_t v[32];
    tc_ld32(taddr, v);
    tc_wait_ld32(v);
    const uint32_t sign = prefilter_unit(cst_unit, v, SS, PP);
    const uint32_t cand = valid ? ~sign : 0u;
    uint32_t any = __reduce_or_sync(0xffffffffu, cand);
    while (any) {  // warp-uniform loop over the (few) columns in which some lane has a candidate
        const int b = 31 - __clz(any);
        any &= ~(1u << b);
        const int j = b >= 16 ? 2 * (31 - b) : 2 * (15 - b) + 1;
        const uint32_t a = tc_ld1(taddr + j);  // re-read that column from TMEM (uniform address)
        tc_wait_ld();
        if ((cand >> b) & 1u) push_candidate(p, a, col + j, page, gx, y);
    }
}

__device__ __forceinline__ void handle_unit(const TcParams &p, const float4 *cst_unit, uint32_t taddr, int col,
                                            const uint32_t (&v)[32], unsigned long long SS, unsigned long long PP,
                                            bool valid, int page, int gx, int y)
{
    const float m = prefilter_unit_max(cst_unit, v, SS, PP);
    if (__any_sync(0xffffffffu, valid && m >= 0.f))
        candidates_of_unit(p, cst_unit, taddr, col, SS, PP, valid, page, gx, y);
}

__global__ void __launch_bounds__(TC_THREADS, 1) scan_tc_kernel(const __grid_constant__ TcParams p)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    // ---- shared memory carve-up
    uint8_t *btile = smem;
    uint8_t *ring = btile + ((p.btile_bytes + 127) & ~127u);
    uint8_t *raw = ring + (size_t)(p.ring + 1) * p.row_pitch;  // +1: mirror of slot 0 for (ring-1, 0) pairs
    uint64_t *bars = (uint64_t *)(raw + TC_RAW_SLOTS * TC_RAW_BYTES);
    uint64_t *bar_btile = bars;                       // 1
    uint64_t *raw_full = bars + 1;                    // TC_RAW_GROUPS
    uint64_t *raw_empty = raw_full + TC_RAW_GROUPS;   // TC_RAW_GROUPS
    uint64_t *a_full = raw_empty + TC_RAW_GROUPS;     // TC_RING_MAX
    uint64_t *a_empty = a_full + TC_RING_MAX;         // TC_RING_MAX
    uint64_t *t_full = a_empty + TC_RING_MAX;         // TC_MAX_BUF
    uint64_t *t_empty = t_full + TC_MAX_BUF;          // TC_MAX_BUF
    uint32_t *tmem_ptr = (uint32_t *)(t_empty + TC_MAX_BUF);
    // 128 x {-b'_2i, -b'_2i+1, -a'_2i, -a'_2i+1}; offset arithmetic on `smem` keeps it a shared-space pointer (LDS.128)
    float4 *cst_s = (float4 *)(smem + (((size_t)((uint8_t *)(tmem_ptr + 4) - smem) + 15) & ~(size_t)15));
    uint32_t *st_ring = (uint32_t *)(cst_s + 128);  // [16 epilogue warps][TC_ST_DEPTH][2][32]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        mbar_init(bar_btile, 1);
        for (int i = 0; i < TC_RAW_GROUPS; i++) {
            mbar_init(raw_full + i, 1);
            mbar_init(raw_empty + i, 4);   // one arrival per expansion warp
        }
        for (int i = 0; i < TC_RING_MAX; i++) {
            mbar_init(a_full + i, 4);
            mbar_init(a_empty + i, 2);     // one tcgen05.commit per MMA-issuing warp
        }
        for (int i = 0; i < TC_MAX_BUF; i++) {
            mbar_init(t_full + i, 1);
            mbar_init(t_empty + i, 4 * TC_EPI_GROUPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x >= 128 && threadIdx.x < 256) {
        const int i = threadIdx.x - 128;
        cst_s[i] = make_float4(p.cst[2 * i].y, p.cst[2 * i + 1].y, p.cst[2 * i].x, p.cst[2 * i + 1].x);
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)),
                     "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    // rows this CTA will stream through the pipeline (all its items); the expansion warps need nothing else
    uint32_t total_rows = 0;
    if (warp < 8) {
        Item it;
        for (int idx = blockIdx.x; get_item(p, idx, it); idx += gridDim.x) total_rows += (it.ys1 - it.ys0) + p.n_hp - 1;
    }

    if (warp == 0) {
        // ================================================================== TMA producer (warp-uniform, one elected lane issues)
        if (elect_one()) {
            mbar_arrive_expect_tx(bar_btile, p.btile_bytes);
            tma_bulk_g2s(btile, p.btile, p.btile_bytes, bar_btile);
        }
        const uint32_t row_bytes = 128 + p.np;
        uint32_t rg = 0, rgpar = 1;  // raw group and the parity of its PREVIOUS use
        uint32_t in_group = 0, slot = 0;
        bool first_round = true;
        Item it;
        for (int idx = blockIdx.x; get_item(p, idx, it); idx += gridDim.x) {
            const uint8_t *src = p.inv + (size_t)it.page * p.inv_page_stride + it.x0 + (size_t)it.ys0 * p.pitch;
            const int n_rows = (it.ys1 - it.ys0) + p.n_hp - 1;
            for (int r = 0; r < n_rows; r++, src += p.pitch) {
                if (in_group == 0 && !first_round) mbar_wait<20000>(raw_empty + rg, rgpar);
                if (elect_one()) {
                    mbar_expect_tx(raw_full + rg, row_bytes);
                    tma_bulk_g2s(raw + slot * TC_RAW_BYTES, src, row_bytes, raw_full + rg);
                    if (in_group == TC_G - 1) mbar_arrive(raw_full + rg);
                }
                __syncwarp();
                slot++;
                if (++in_group == TC_G) {
                    in_group = 0;
                    if (++rg == TC_RAW_GROUPS) rg = 0, slot = 0, rgpar ^= 1, first_round = false;
                }
            }
        }
        if (in_group != 0 && elect_one()) mbar_arrive(raw_full + rg);  // the last, partial group
    } else if (warp == 1 || warp == 2) {
        // ================================================================== MMA issuers (two warps)
        // One warp cannot issue fast enough: its serial instruction stream (descriptor arithmetic, R2UR,
        // UTCIMMA, commits; ~7 cycles per dependent instruction) costs more per output row than the
        // tensor core needs for the row's MMAs.  Output rows therefore alternate between two issuing
        // warps (row parity); they touch different accumulators, so their relative order is free.
        //   * a page-row group goes back to the expansion warps when BOTH issuers have moved past it:
        //     each commits to a_empty[g] (count 2) once its next output no longer reads group g --
        //     tcgen05.commit only tracks the MMAs of the committing thread.
        const uint32_t mw = warp - 1;
        const uint32_t idesc = (2u << 4)                         // D format: S32
                               | (0u << 7) | (0u << 10)          // A, B: unsigned 8-bit
                               | (0u << 15) | (0u << 16)         // A, B: K-major
                               | ((uint32_t)(p.n_mma >> 3) << 17)  // N
                               | ((128u >> 4) << 24);            // M = 128
        const uint32_t ring_n = p.ring, ring_g = p.ring_groups, nbuf = p.nbuf, ksteps = p.ksteps, n_hp = p.n_hp;
        const uint32_t b_lbo16 = ((uint32_t)p.nb * 16u) >> 4;
        const uint32_t a_lbo16 = (p.np == 16 ? (uint32_t)p.row_pitch : 256u) >> 4;
        const uint32_t pitch16 = (uint32_t)p.row_pitch >> 4;
        const uint32_t a_step = p.np == 16 ? 2u : 1u;            // ring slots consumed per K step
        const uint32_t desc_hi = (128u >> 4) | (1u << 14);       // SBO = 128 B, version = 1
        const uint32_t a_lo0 = ((smem_u32(ring) & 0x3FFFFu) >> 4) | (a_lbo16 << 16);
        const uint32_t b_lo0 = ((smem_u32(btile) & 0x3FFFFu) >> 4) | (b_lbo16 << 16);
        const uint32_t a_inc = a_step * pitch16, a_wrap = ring_n * pitch16, a_end = a_lo0 + a_wrap;
        mbar_wait(bar_btile, 0);
        uint32_t g_first = 0;                    // global page-row index of the current output row's first row
        uint32_t s_first = 0;                    // its ring slot (g_first mod ring_n)
        uint32_t rel_g = 0, rel_rows = TC_G;     // next group to hand back; rel_rows = 4*(groups released + 1)
        uint32_t new_g = 0, new_par = 0, rows_ready = 0;  // a_full bookkeeping (per warp)
        uint32_t job = 0, buf = 0, bpar = 0;     // accumulator sequence (all jobs, both warps count them)
        bool first_round = true;
        Item it;
        for (int idx = blockIdx.x; get_item(p, idx, it); idx += gridDim.x) {
            const int n_out_rows = it.ys1 - it.ys0;
            for (int j = 0; j < n_out_rows; j++, job++) {
                if ((job & 1u) == mw) {
                    // hand back every group that lies entirely below this output's first page row
                    while (rel_rows <= g_first) {
                        if (elect_one()) tc_commit(a_empty + rel_g);
                        __syncwarp();
                        rel_rows += TC_G;
                        if (++rel_g == ring_g) rel_g = 0;
                    }
                    while (rows_ready < g_first + n_hp) {   // page rows this output needs
                        mbar_wait<20000>(a_full + new_g, new_par);
                        rows_ready += TC_G;
                        if (++new_g == ring_g) new_g = 0, new_par ^= 1;
                    }
                    if (!first_round) mbar_wait<20000>(t_empty + buf, bpar ^ 1);
                    tc_fence_after();
                    {
                        const bool leader = elect_one();
                        const uint32_t d0 = tmem_base + buf * p.nbs;
                        uint32_t al0 = a_lo0 + s_first * pitch16, bl0 = b_lo0;
                        for (uint32_t k = 0; k < ksteps; k++) {
                            if (leader) {
                                tc_mma_i8(d0, ((uint64_t)desc_hi << 32) | al0, ((uint64_t)desc_hi << 32) | bl0, idesc, k);
                                if (k + 1 == ksteps) tc_commit(t_full + buf);   // accumulator ready for the epilogue
                            }
                            al0 += a_inc;
                            if (al0 >= a_end) al0 -= a_wrap;
                            bl0 += 2 * b_lbo16;
                        }
                    }
                    __syncwarp();
                }
                if (++buf == nbuf) buf = 0, bpar ^= 1, first_round = false;
                g_first++;
                if (++s_first == ring_n) s_first = 0;
            }
            // the item's last n_hp-1 page rows are not the first row of any output
            g_first += n_hp - 1;
            s_first += n_hp - 1;
            while (s_first >= ring_n) s_first -= ring_n;
        }
    } else if (warp >= 4 && warp < 8) {
        // ================================================================== Toeplitz expansion
        // one warp per page row, four rows (one group) in flight per handshake
        const int w = warp - 4;
        const bool mirror = p.np == 16;
        const uint32_t ring_g = p.ring_groups;
        uint32_t rg = 0, rgpar = 0, ag = 0, agpar = 1;
        bool first_round = true;
        for (uint32_t g0 = 0; g0 < total_rows; g0 += TC_G) {
            mbar_wait<20000>(raw_full + rg, rgpar);
            if (!first_round) mbar_wait<20000>(a_empty + ag, agpar);
            if (g0 + w < total_rows) {
                const uint32_t *rw = (const uint32_t *)(raw + (rg * TC_G + w) * TC_RAW_BYTES);
                const uint32_t s = ag * TC_G + w;
                uint8_t *dst = ring + (size_t)s * p.row_pitch;
                for (int ee = lane; ee < p.n_entries; ee += 32) {
                    const uint32_t *wp = rw + (ee >> 2);
                    const int sh = (ee & 3) * 8;
                    const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2], w3 = wp[3], w4 = wp[4];
                    uint4 o;
                    o.x = __funnelshift_r(w0, w1, sh);
                    o.y = __funnelshift_r(w1, w2, sh);
                    o.z = __funnelshift_r(w2, w3, sh);
                    o.w = __funnelshift_r(w3, w4, sh);
                    *(uint4 *)(dst + ee * 16) = o;
                    if (mirror && s == 0) *(uint4 *)(ring + (size_t)p.ring * p.row_pitch + ee * 16) = o;
                }
            }
            fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's async proxy
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(a_full + ag);
                mbar_arrive(raw_empty + rg);
            }
            if (++rg == TC_RAW_GROUPS) rg = 0, rgpar ^= 1;
            if (++ag == ring_g) ag = 0, agpar ^= 1, first_round = false;
        }
    } else if (warp >= 8) {
        // ================================================================== epilogue (16 warps)
        // 4 warps per TMEM lane quarter (= per SM sub-partition); the 32-column units of the accumulator
        // buffers are dealt round-robin to the 4 warps ACROSS buffers so they stay balanced.
        const int e = warp - 8;
        const int q = e & 3;                      // TMEM lane quarter this warp may access (warp % 4)
        const int grp = e >> 2;
        const int m = q * 32 + lane;              // window within the strip
        const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
        const uint32_t nbuf = p.nbuf;
        const int nunits = p.nunits, nsub = p.nsub, urot = p.nunits % TC_EPI_GROUPS;
        int ufirst = grp;  // first unit of the current buffer that belongs to this warp
        uint32_t buf = 0, bpar = 0;
        Item it;
        for (int idx = blockIdx.x; get_item(p, idx, it); idx += gridDim.x) {
            const int gx = it.x0 + m;
            const bool x_ok = gx >= 1 && gx <= p.r_w - p.n_w;  // ncc.rs:281: x starts at 1
            const size_t plane = (size_t)it.page * p.plane_page_stride + gx;
            // Window statistics (s_p, norm_p) of this lane's window: streamed TC_ST_DEPTH rows ahead through a
            // private shared-memory ring with cp.async.  (Register prefetching does not work here: rotating
            // the registers makes the compiler wait for the in-flight load one row after it was issued,
            // and a row can take less than an L2 round trip.)  Every lane reads back only what it copied
            // itself, so cp.async.wait_group is all the synchronisation needed.
            const uint32_t *sp_col = p.sp + plane;
            const float *pf_col = p.pf + plane;
            uint32_t *st_mine = st_ring + (size_t)e * TC_ST_DEPTH * 64 + lane;
            int st_slot = 0;
#pragma unroll 1
            for (int d = 0; d < TC_ST_DEPTH; d++) {
                const int yy = it.ys0 + d;
                if (x_ok && yy < it.ys1) {
                    cp_async4(st_mine + d * 64, sp_col + (size_t)yy * p.spitch);
                    cp_async4(st_mine + d * 64 + 32, pf_col + (size_t)yy * p.spitch);
                }
                cp_async_commit();
            }
            for (int y = it.ys0; y < it.ys1; y++) {
                cp_async_wait<TC_ST_DEPTH - 1>();
                const uint32_t s_p = st_mine[st_slot * 64];
                const float P = __uint_as_float(st_mine[st_slot * 64 + 32]);
                {
                    const int yy = y + TC_ST_DEPTH;
                    if (x_ok && yy < it.ys1) {
                        cp_async4(st_mine + st_slot * 64, sp_col + (size_t)yy * p.spitch);
                        cp_async4(st_mine + st_slot * 64 + 32, pf_col + (size_t)yy * p.spitch);
                    }
                    cp_async_commit();
                    if (++st_slot == TC_ST_DEPTH) st_slot = 0;
                }
                const bool valid = x_ok && P < __int_as_float(0x7f800000);  // +inf marks a constant window
                const float S = (float)s_p;
                const float Pv = valid ? P : 0.f;
                const unsigned long long SS = pack2(S, S), PP = pack2(Pv, Pv);
                for (int sub = 0; sub < nsub; sub++) {
                    mbar_wait<20000>(t_full + buf, bpar);
                    tc_fence_after();
                    const uint32_t tb = tlane + buf * p.nbs;
                    const int cbase = sub * p.n_mma;
                    if (p.dbg_mode != 1) {
                        for (int u = ufirst; u < nunits; u += TC_EPI_GROUPS) {
                            uint32_t v[32];
                            if (p.dbg_mode != 4) {
                                tc_ld32(tb + u * 32, v);
                                tc_wait_ld32(v);
                            } else {
#pragma unroll
                                for (int i = 0; i < 32; i++) v[i] = (uint32_t)(u + i) * 3u;  // timing experiment: math without TMEM reads
                            }
                            if (p.dbg_mode == 3) {  // timing experiment: TMEM reads without math
                                if (v[5] == 0x7fffffffu && v[17] == 0x12345u) push_candidate(p, v[0], 0, 0, 0, 0);
                                continue;
                            }
                            handle_unit(p, cst_s + (((sub * nunits + u) * 32) >> 1), tb + u * 32, cbase + u * 32, v, SS, PP, valid,
                                        it.page, gx, y);
                        }
                    }
                    if (p.dbg_acc && p.dbg_col >= cbase && p.dbg_col < cbase + p.n_mma && grp == 0) {
                        const uint32_t a = tc_ld1(tb + p.dbg_col - cbase);
                        tc_wait_ld();
                        if (x_ok) p.dbg_acc[(size_t)y * p.r_w + gx] = a;
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(t_empty + buf);
                    if (++buf == nbuf) buf = 0, bpar ^= 1;
                    ufirst -= urot;
                    if (ufirst < 0) ufirst += TC_EPI_GROUPS;
                }
            }
        }
    }

    // ---- teardown
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------- exact pass
// Every prefilter survivor gets the reference's own f64 arithmetic (ncc.cpp:212-220 via ncc_exact,
// ncc.rs:309-311 via patch_rnorm): the decision `sim > threshold` and the f32 score are bit-identical
// to the CPU.  Runs right after the MMA launches of a box size, while its statistic planes are live.
struct CandArgs {
    const Hit *cands;
    uint32_t cand_cap;
    const unsigned int *cand_count;
    unsigned int *cand_max;  // high-water mark of the candidate count (overflow detection on the host)
    const uint32_t *tpl_of;  // [n_blocks*nb] bank index per class column (0xFFFFFFFF = padding)
    const TplInfo *tpl;
    const uint32_t *sp, *s2p;
    int spitch;
    size_t plane_page_stride;
    double n_d, thr_d;
    HitSink sink;
};

__global__ void __launch_bounds__(256) cand_exact_kernel(CandArgs a)
{
    const unsigned total = *a.cand_count;
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicMax(a.cand_max, total);
    const unsigned n = min(total, a.cand_cap);
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const Hit c = a.cands[i];
        const uint32_t t = a.tpl_of[c.t];
        if (t == 0xFFFFFFFFu) continue;
        const uint32_t y = c.yx >> 16, x = c.yx & 0xFFFFu;
        const size_t o = (size_t)c.page * a.plane_page_stride + (size_t)y * a.spitch + x;
        const uint32_t s_p = a.sp[o], s2_p = a.s2p[o];
        const TplInfo ti = a.tpl[t];
        const double rn_p = patch_rnorm(s_p, s2_p, a.n_d);
        float sim;
        if (ncc_exact(__float_as_uint(c.sim), s_p, rn_p, ti.s_n, ti.n_recip, ti.rnorm_n, a.thr_d, &sim)) {
            auto g = cg::coalesced_threads();
            unsigned base = 0;
            if (g.thread_rank() == 0) base = atomicAdd(a.sink.hit_count, (unsigned)g.size());
            base = g.shfl(base, 0);
            const unsigned slot = base + g.thread_rank();
            if (slot < a.sink.hit_cap) {
                Hit h;
                h.t = t;
                h.yx = c.yx;
                h.sim = sim;
                h.page = c.page;
                a.sink.hits[slot] = h;
            }
            atomicAdd(a.sink.rowcount + ((size_t)c.page * a.sink.T + t) * a.sink.r_h + y, 1u);
        }
    }
}

// ---------------------------------------------------------------------------------------------- host side
static size_t tc_smem_bytes(uint32_t btile_bytes, int ring, int row_pitch)
{
    return ((btile_bytes + 127) & ~127u) + (size_t)(ring + 1) * row_pitch + TC_RAW_SLOTS * TC_RAW_BYTES +
           (1 + 2 * TC_RAW_GROUPS + 2 * TC_RING_MAX + 2 * TC_MAX_BUF) * 8 + 64 + 128 * 16 + 16 * TC_ST_DEPTH * 64 * 4;
}

int tc_class_build(TcClass &tc, const uint8_t *rows_host, uint32_t n_w, uint32_t n_h, uint32_t np, uint32_t n_tpl,
                   const uint32_t *bank_index, const TplInfo *info)
{
    tc.supported = false;
    tc.n_w = n_w;
    tc.n_h = n_h;
    tc.np = np;
    tc.n_tpl = n_tpl;
    const uint32_t n_hp = np == 16 ? (n_h + 1) & ~1u : n_h;
    tc.kchunks = n_h * (np / 16);
    tc.ksteps = (tc.kchunks + 1) / 2;
    const int ring_groups = ((int)n_hp + 1 + TC_G - 1 + TC_G - 1) / TC_G + TC_LOOK_GROUPS;
    const int ring = ring_groups * TC_G;
    const int row_pitch = np == 16 ? 2048 : 2304;
    if (ring_groups > TC_RING_MAX) return 0;  // unsupported shape -> SIMT kernel
    // largest NB (multiple of 16, <= 256) whose B tile fits next to the ring
    int nb_max = getenv("FOCR_TC_NBMAX") ? atoi(getenv("FOCR_TC_NBMAX")) : 256;  // experiment knob
    while (nb_max >= 16 && tc_smem_bytes(2 * tc.ksteps * nb_max * 16, ring, row_pitch) > TC_SMEM_BUDGET) nb_max -= 16;
    if (nb_max < 16) return 0;
    tc.n_blocks = (n_tpl + nb_max - 1) / nb_max;
    tc.nb = ((n_tpl + tc.n_blocks - 1) / tc.n_blocks + 15) & ~15u;
    if (tc.nb > 128) tc.nb = (tc.nb + 31) & ~31u;  // issued as two MMAs of nb/2 (a multiple of 16) columns
    const size_t tile = (size_t)2 * tc.ksteps * tc.nb * 16;
    std::vector<uint8_t> bt(tile * tc.n_blocks, 0);
    std::vector<float2> cst((size_t)tc.n_blocks * tc.nb);
    std::vector<uint32_t> tof((size_t)tc.n_blocks * tc.nb, 0xFFFFFFFFu);
    const float inf = INFINITY;
    for (auto &c : cst) c = make_float2(inf, 0.f);
    for (uint32_t i = 0; i < n_tpl; i++) {
        const uint32_t blk = i / tc.nb, n = i % tc.nb;
        for (uint32_t kc = 0; kc < tc.kchunks; kc++) {
            const uint32_t row = np == 16 ? kc : kc / 2, boff = np == 16 ? 0 : (kc & 1) * 16;
            memcpy(&bt[blk * tile + ((size_t)kc * tc.nb + n) * 16], rows_host + ((size_t)i * n_h + row) * np + boff, 16);
        }
        const TplInfo &ti = info[i];
        // norm_n = sqrt(s2_n - s_n^2/n) = 1/rnorm_n ; constant (incl. all-zero) templates can never hit
        const double norm_n = 1.0 / ti.rnorm_n;
        const bool ok = std::isfinite(ti.rnorm_n) && ti.rnorm_n > 0 && std::isfinite(norm_n);
        cst[(size_t)blk * tc.nb + n] = make_float2(ok ? (float)norm_n : inf, (float)(ti.s_n * ti.n_recip));
        tof[(size_t)blk * tc.nb + n] = bank_index[i];
    }
    if (cudaMalloc(&tc.b_tiles, bt.size()) != cudaSuccess) return -1;
    if (cudaMalloc(&tc.tpl_of, tof.size() * 4) != cudaSuccess) return -1;
    if (cudaMemcpy(tc.b_tiles, bt.data(), bt.size(), cudaMemcpyHostToDevice) != cudaSuccess) return -1;
    if (cudaMemcpy(tc.tpl_of, tof.data(), tof.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) return -1;
    tc.consts_host = new std::vector<float2>(std::move(cst));
    tc.supported = true;
    return 0;
}

void tc_class_release(TcClass &tc)
{
    if (tc.b_tiles) cudaFree(tc.b_tiles);
    if (tc.consts) cudaFree(tc.consts);
    if (tc.tpl_of) cudaFree(tc.tpl_of);
    delete (std::vector<float2> *)tc.consts_host;
    tc.b_tiles = nullptr;
    tc.consts = nullptr;
    tc.tpl_of = nullptr;
    tc.consts_host = nullptr;
    tc.supported = false;
}

bool tc_class_supported(const TcClass &tc) { return tc.supported; }
void tc_workspace_release(TcWorkspace &) {}

cudaError_t launch_scan_tc(TcWorkspace &, const TcClass &tc, const ScanArgs &a, int n_pages, int sm_count,
                           cudaStream_t st, int *n_launches, uint32_t *dbg_acc, int dbg_pos)
{
    if (!tc.supported) return cudaErrorNotSupported;
    thread_local TcParams p;  // 2.3 KB: keep it off the stack
    p.inv = a.inv;
    p.inv_page_stride = a.inv_page_stride;
    p.pitch = a.pitch;
    p.r_w = a.r_w;
    p.r_h = a.r_h;
    p.n_w = tc.n_w;
    p.n_h = tc.n_h;
    p.np = tc.np;
    p.n_hp = tc.np == 16 ? (tc.n_h + 1) & ~1u : tc.n_h;
    p.ksteps = tc.ksteps;
    p.nb = tc.nb;
    // one MMA covers all nb columns: a tcgen05.mma carries ~100 cycles of fixed cost (measured), so fewer,
    // wider instructions win over finer TMEM buffering; FOCR_TC_NSUB2 re-enables the split for experiments
    p.nsub = 1;
    p.n_mma = tc.nb / p.nsub;
    p.nunits = (p.n_mma + 31) / 32;
    p.nbs = (p.n_mma + 31) & ~31;
    p.nbuf = std::min(512 / p.nbs, TC_MAX_BUF);
    p.ring_groups = (p.n_hp + 1 + TC_G - 1 + TC_G - 1) / TC_G + TC_LOOK_GROUPS;  // rows y..y+n_hp (two output rows in flight) may straddle one more group
    p.ring = p.ring_groups * TC_G;
    p.row_pitch = tc.np == 16 ? 2048 : 2304;
    p.n_entries = tc.np == 16 ? 128 : 144;
    p.btile_bytes = 2 * tc.ksteps * tc.nb * 16;
    p.sp = a.sp;
    p.pf = a.pf;
    p.spitch = a.spitch;
    p.plane_page_stride = a.plane_page_stride;
    p.cands = a.cands;
    p.cand_cap = a.cand_cap;
    p.cand_count = a.cand_count;
    p.n_pages = n_pages;
    const int xs = a.r_w - (int)tc.n_w + 1, ys = a.r_h - (int)tc.n_h;  // output rows 1 .. r_h-n_h
    if (xs <= 0 || ys <= 0) return cudaSuccess;
    p.n_xstrips = (xs + 127) / 128;
    p.n_ysegs = (ys + TC_YSEG - 1) / TC_YSEG;
    const size_t smem = tc_smem_bytes(p.btile_bytes, p.ring, p.row_pitch);
    cudaError_t e = cudaFuncSetAttribute(scan_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    const int items = n_pages * p.n_xstrips * p.n_ysegs;
    const int grid = std::min(items, sm_count);
    // prefilter margin: constants shrunk towards "more candidates" by 2^-12 (>= 500x the fp32 error)
    const float thr = (float)a.thr_d;
    const float up = 1.0f + 1.0f / 4096.0f, dn = 1.0f - 1.0f / 4096.0f;
    const std::vector<float2> &cst = *(const std::vector<float2> *)tc.consts_host;
    p.dbg_acc = dbg_acc;
    {
        const char *dm = getenv("FOCR_TC_DBG");
        p.dbg_mode = dm ? atoi(dm) : 0;
    }
    p.dbg_col = dbg_acc ? dbg_pos % (int)tc.nb : -1;
    for (uint32_t blk = 0; blk < tc.n_blocks; blk++) {
        if (dbg_acc && blk != (uint32_t)dbg_pos / tc.nb) continue;
        // table index = (sub * nunits + unit) * 32 + j  (each sub-block padded to whole 32-column units)
        for (int idx = 0; idx < 256; idx++) {
            float2 c = make_float2(-INFINITY, 0.f);  // padding: d = -inf -> never a candidate
            const int sub = idx / (p.nunits * 32), within = idx % (p.nunits * 32);
            const int col = sub * p.n_mma + within;
            if (sub < p.nsub && within < p.n_mma && col < (int)tc.nb) {
                const float2 s = cst[(size_t)blk * tc.nb + col];
                if (std::isfinite(s.x)) {
                    const float aa = thr * s.x;  // a = thr * norm_n
                    c.x = -(aa >= 0 ? aa * dn : aa * up);
                    c.y = -(s.y * dn);
                }
            }
            p.cst[idx] = c;
        }
        p.btile = tc.b_tiles + (size_t)blk * p.btile_bytes;
        p.col_base = blk * tc.nb;
        scan_tc_kernel<<<grid, TC_THREADS, smem, st>>>(p);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        if (n_launches) (*n_launches)++;
    }
    if (!dbg_acc) {
        CandArgs ca;
        ca.cands = a.cands;
        ca.cand_cap = a.cand_cap;
        ca.cand_count = a.cand_count;
        ca.cand_max = a.cand_max;
        ca.tpl_of = tc.tpl_of;
        ca.tpl = a.tpl;
        ca.sp = a.sp;
        ca.s2p = a.s2p;
        ca.spitch = a.spitch;
        ca.plane_page_stride = a.plane_page_stride;
        ca.n_d = (double)(tc.n_w * tc.n_h);
        ca.thr_d = a.thr_d;
        ca.sink = a.sink;
        cand_exact_kernel<<<sm_count * 4, 256, 0, st>>>(ca);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        if (n_launches) (*n_launches)++;
    }
    return cudaSuccess;
}

}  // namespace focr
