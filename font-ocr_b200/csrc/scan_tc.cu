// scan_tc.cu -- the NCC template scan as a tcgen05 GEMM (sm_100a) whose NORMALISATION also runs on the
// tensor core, followed by an exact pass over the few survivors.
//
// Replaces the hot loops of ncc_8_u8 / ncc_16_u8 (ncc.cpp:98-248, 302-393) for ALL templates of one
// box size at once.  For an output row y of a 128-window strip,
//
//     acc[m, t] = sum_{ny, j} page[y+ny][x0+m+j] * tpl[t][ny][j]          (exact, u8 x u8 -> s32)
//
// is a GEMM with M = 128 windows (TMEM lanes), N = templates (TMEM columns, <= 256) and K = 16*n_h.
// The reference then tests  sim = (acc - s_n*s_p/n) * rnorm_n * rnorm_p > thr  per (window, template)
// (ncc.cpp:212-220).  As a real inequality that is  d = acc - b_t*S_m - a_t*P_m > 0  with b_t = s_n/n,
// a_t = thr*norm_n (per template) and S_m = s_p, P_m = norm_p (per window): a RANK-2 correction.
//
// Design:
//  * A operand = Toeplitz (im2col) rows.  UMMA descriptors need core-matrix rows 16 B apart, windows
//    are 1 B apart, so each page row is expanded ONCE into a 128 x 16 B block in shared memory and
//    reused by all n_h vertical taps and all N templates (descriptor LBO points at ring slots y+2k, y+2k+1).
//    Raw page rows arrive by TMA bulk copies (cp.async.bulk + mbarrier complete_tx), 4 rows per barrier.
//  * 1 x tcgen05.mma kind::f16 (K = 16, fp32 result, accumulate OFF) first writes F = C0 - (b_t*S_m + a_t*P_m)
//    into the accumulator: A2[m] = {S_hi, S_hi, S_lo, P_1, P_1, P_2, V, 2^15} and
//    B2[t] = {-b_1, -b_2, -b_1, -a_1, -a_2, -a_1, -BIG, 510 | -BIG} are fp16 hi/lo splits (error < 8 in units of
//    acc); C0 = 2^15 * 510 = 2^24 - 2^16.  V = BIG marks windows that can never hit (constant / out of range),
//    the last B2 entry is -BIG for padding columns.  F lies in [2^23, 2^24) where fp32 has ulp 1, so its BITS are
//    0x4B000000 + (F - 2^23): an integer that is linear in F.
//  * 7 x tcgen05.mma kind::i8 then accumulate acc EXACTLY, as s32, onto those bits: the cell now holds the bits
//    of the fp32 number C0 + d (+- a few ulp; if the sum leaves the binade upwards it only over-estimates, fp32
//    bit patterns are monotonic).  Nothing has to be written back before the next row: the epilogue is ONE
//    3-input max per column pair and a compare against C0 - margin -- no conversion, no FMA, no per-column
//    constants, no tcgen05.st.  Windows whose b_max*S + a_max*P would push F below 2^23 (finer ulp, the bits
//    stop being linear and would UNDER-estimate) are flagged V = -BIG instead: they always survive the screen.
//  * survivors (d >= -margin; a few per thousand outputs) go to a candidate list; cand_exact_kernel
//    recomputes acc with integer arithmetic and replays the reference's f64 normalisation operation for
//    operation, so decisions and f32 scores are bit-identical to the CPU.  Nothing can be missed: every
//    approximation errs towards MORE candidates.
//
//  * A launch covers the templates of one box size or of TWO box sizes of the same height: the correlation GEMM only
//    sees zero-padded template rows, and the fp16 MMA's two 16-byte K chunks carry one set of window statistics each
//    (a column's B2 entry is non-zero only in the chunk of its own box size).
//  * A JOB = one output row x one sub-block of <= 256 columns (a launch has 1 or 2 sub-blocks per row).  Jobs walk a
//    RING of TMEM accumulators (512 / columns per sub-block of them): TWO converged warps issue alternate jobs (all
//    their state in uniform registers), two teams of 4 epilogue warps (one per TMEM lane quarter) drain alternate jobs and
//    hand the accumulator back as soon as their last tcgen05.ld has landed, before they screen.  With three accumulators the
//    hand-back chain (commit -> team wakes -> tcgen05.ld -> release -> issuer wakes) has two whole jobs of tensor time.
//  * Columns are ordered by template similarity (tc_class_build): a window that matches a glyph matches its look-alikes,
//    and clustering them into the same 32-column unit makes the epilogue's slow path run once for them, not once per unit.
//  * Boxes at most 8 wide pack two template rows per 16-byte K chunk (ring slot r = 8-byte windows of page rows r, r+1).
//
// Warp roles (one persistent CTA per SM over (page, x-strip, y-segment) items):
//   warp 0     TMA producer of raw page rows         warps 1-2   MMA issuers (converged warps, elect.sync-guarded tcgen05)
//   warp 3     TMEM alloc, otherwise idle            warps 4-7   Toeplitz expansion (one warp per row)
//   warps 8-11 A2 rows (window statistics -> fp16)   warps 12-19 epilogue (2 teams x 4 TMEM lane quarters)
#include <cooperative_groups.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

#include "scan_tc.cuh"

namespace cg = cooperative_groups;

namespace focr {

constexpr int TC_EPI_WARPS = 8;       // epilogue warps: 2 teams (alternate jobs) x 4 TMEM lane quarters
constexpr int TC_THREADS = 384 + 32 * TC_EPI_WARPS;
static_assert(TC_LISTS_PER_CTA == TC_EPI_WARPS, "scan_tc.cuh: candidate lists per CTA = epilogue warps");
constexpr int TC_EPI_UNITS = 4;       // 32-column units an epilogue warp holds in registers at a time
constexpr int TC_G = 4;               // rows per pipeline group: one mbarrier handshake per 4 rows
constexpr int TC_RAW_GROUPS = 4;      // raw page-row ring (TMA destination): 4 groups x 4 rows x 160 B
constexpr int TC_RAW_SLOTS = TC_RAW_GROUPS * TC_G;
constexpr int TC_RAW_BYTES = 288;     // per raw slot: one page row of 128 + 16 (+16: boxes wider than 16) bytes, or two rows of 144 (packed mode)
constexpr int TC_LOOK_GROUPS = 2;     // expanded row groups the producer side may run ahead of the MMA (1 for tall boxes)
constexpr int TC_RING_MAX = 12;       // max ring groups
constexpr int TC_A2_GROUPS = 4;       // A2 (statistics) ring: up to 4 groups x 4 output rows x 2 KB (2 for tall boxes)
constexpr int TC_MAX_BUF = 8;         // TMEM accumulator buffers
// setmaxnreg budget (the kernel is launched with 96 registers x 640 threads = 61440): warps 0-3 (TMA producer, MMA
// issuers: their state lives in uniform registers) keep 56, Toeplitz warps 40, A2 warps 48, and the 8 epilogue warps, which
// hold four 32-column units at a time, take 168: 128 x (56 + 40 + 48) + 256 x 168 = 61440
constexpr int TC_REGS_TOEPLITZ = 40;
constexpr int TC_REGS_A2 = 48;
constexpr int TC_REGS_EPILOGUE = 168;
constexpr int TC_REGS_ISSUE = 56;
constexpr size_t TC_SMEM_BUDGET = 220 * 1024;
constexpr float TC_C0 = 16711680.f;               // 2^15 * 510 = 2^24 - 2^16: the constant term of the fp16 MMA
constexpr float TC_KCAP = 8323072.f - 8192.f;     // C0 - 2^23 minus slack: largest b*S + a*P that keeps F in [2^23, 2^24)
constexpr float TC_BIG = 60000.f;                 // fp16-representable "never" marker (BIG*BIG = 3.6e9 >> any acc)
constexpr float TC_MARGIN = 256.f;                // absolute slack of the tensor-core normalisation, in units of acc (error budget < 60,
                                                  // + a * 2^-6 <= 64 for norm_p rounded to 1/32 in the packed statistics word)

struct TcParams {
    const uint8_t *inv;
    size_t inv_page_stride;
    int pitch, r_w, r_h;
    int n_w, n_w2, n_h, np;   // n_w2: width of the second box size (ncls == 2)
    int ncls;          // box sizes in this launch group (1 or 2)
    int n_hp;          // ring slots (page rows) an output row spans: n_h rounded up to 2 (np == 16), n_h (np == 32), 4*ksteps - 1 (packed)
    int packed;        // boxes at most 8 wide: ring slot r holds the 8-byte windows of page rows r AND r+1, a K chunk is two template rows
    int ksteps;        // tcgen05.mma kind::i8 per output row
    int nb;            // columns per launch = nsub * nbs
    int nsub;          // sub-blocks: jobs (accumulators) per output row
    int nunits;        // 32-column epilogue units per accumulator
    int nbs;           // columns per sub-block = TMEM column stride between accumulators = column stride of the B tile (multiple of 32)
    int nmma[2];       // N of the MMAs per sub-block: its real columns rounded up to 16 (<= nbs; the columns beyond are never written)
    int nbuf;          // accumulators in the ring: job k uses number k mod nbuf
    int ring;          // expanded-row ring slots = ring_groups * 4
    int ring_groups;
    int n_mirror;      // ring slots stored twice (see tc_mma_role); 0: the issue loop wraps every K step (np == 32)
    int a2_groups;     // groups of the A2 ring (<= TC_A2_GROUPS)
    int a2_slot;       // bytes per A2 row: 2048 per box size
    int sshift;        // the templates of the B tile are ceil(t / 2^sshift): S is split at bit 10 instead of 6 (see A2 rows)
    int row_pitch;     // bytes per expanded row slot
    int n_entries;     // 16-byte entries per expanded row (128, or 144 for np == 32)
    int yseg;          // output rows per work item
    const uint8_t *btile;     // [nsub][2*ksteps][nbs][16]
    uint32_t btile_bytes;
    const float4 *colconst;   // [nb] {norm_n, s_n/n, box size index, -}; norm_n = +inf for padding / constant templates
    float thr;
    float bmax[2], amax[2];   // per box size: max over this launch's columns of s_n/n and max(thr*norm_n, 0)
    uint32_t col_base;        // this launch's first column within the group (N-block * nb)
    const uint32_t *sp[2];    // window statistics planes per box size
    const float *pf[2];
    int pack;                 // sp holds `s_p | fix11.5(norm_p) << 16` per window (0xFFFF: constant window), pf is unused
    int spitch;
    size_t plane_page_stride;
    Hit *cands;               // candidate lists, one PRIVATE list per epilogue warp: [grid*TC_LISTS_PER_CTA][cand_cap] {group column, y<<16|x, -, page}
    uint32_t cand_cap;        // entries per warp list
    unsigned int *cand_count; // [grid*TC_LISTS_PER_CTA] entries each warp produced (may exceed cand_cap -> the host grows the lists and retries)
    int n_pages, n_xstrips, n_ysegs;
    unsigned int *wd;   // watchdog words (see mbar_wait): [0] raised, [1] tag, [2] info, [3] CTA, [4] warp, [5] parity
    long long *trace;   // timing experiments (env FOCR_TC_TRACE=file): CTA 0's roles add up the cycles they spend in each
                        // wait / phase, [64] (tools/tc_trace.py names the slots)
    int timeline;       // timing experiments: record the event timeline instead of the role time budget
    uint32_t tl0;       // first job of the timeline window
    int dbg_mode;       // timing experiments only (env FOCR_TC_DBG, a bit mask; results are wrong when non-zero):
                        // 1 epilogue skips the TMEM reads, 2 epilogue loads but does not screen, 4 no MMAs are issued,
                        // 8 A2 rows skip their global loads, 16 no Toeplitz expansion, 32 A2 rows skip their stores
    uint32_t *dbg_acc;  // parity probe: raw numerators of column dbg_col, [y*r_w+x]; disables the fp16 MMA
    int dbg_col, dbg_xlast;   // dbg_xlast = r_w - (box width of that column)
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
__device__ uint4 *g_wdlog = nullptr;
__device__ uint32_t g_wdprog[32];

// Wait for a phase of an mbarrier with a suspend-time hint: the hardware parks the thread (few issue
// slots taken from busy warps) and wakes it as soon as the phase completes.
// Watchdog: a wait that has timed out ~8000 times (>= 0.15 s) records who is stuck (wd[1..4] = tag, info, CTA,
// warp) and raises wd[0]; every waiting thread that sees wd[0] gives up, so a protocol bug ends the kernel with an
// error the host reports instead of hanging the GPU.
// slow path of the watchdog, OUT OF LINE: the wait loops sit in single-warp roles whose speed depends on how few
// instructions and instruction-cache lines they touch (inlining this at every wait cost 12 % of the kernel's speed).
// Returns true when the wait must give up.
__device__ __noinline__ bool mbar_watchdog(unsigned int *wd, uint32_t tries, uint32_t tag, uint32_t info, uint32_t parity,
                                           uint32_t addr, volatile uint32_t *prog)
{
    if (tries == 4096u && prog && atomicCAS(wd + 6, 0u, 1u) == 0u) {  // first long wait anywhere: snapshot of the CTA's progress
        for (int i = 0; i < 32; i++) g_wdprog[i] = prog[i];
        wd[7] = blockIdx.x, wd[8] = threadIdx.x >> 5, wd[9] = tag;
        __threadfence();
    }
    if (tries == 8192u && g_wdlog) {  // debugging aid (env FOCR_TC_WDLOG): what this warp has been stuck on
        uint4 *slot = g_wdlog + (size_t)blockIdx.x * 32 + (threadIdx.x >> 5);
        if (slot->x == 0) *slot = make_uint4(tag, info, parity, addr);
    }
    // give up after ~0.3 s, or ~0.15 s when another wait has already given up
    if (tries >= 16384u || (tries >= 8192u && *(volatile unsigned int *)wd != 0)) {
        if (atomicCAS(wd, 0u, 1u) == 0u) {
            wd[1] = tag, wd[2] = info, wd[3] = blockIdx.x, wd[4] = threadIdx.x >> 5;
            wd[5] = parity;
            __threadfence();
        }
        return true;
    }
    return false;
}

template <bool HINT = true>
__device__ __forceinline__ void mbar_wait_addr(const uint32_t addr, uint32_t parity, unsigned int *wd = nullptr, uint32_t tag = 0,
                                               uint32_t info = 0, volatile uint32_t *prog = nullptr)
{
    uint32_t done, tries = 0;
    for (;;) {
        if (HINT)
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(addr), "r"(parity), "r"(20000u)
                : "memory");
        else  // accumulator hand-offs: try_wait without a suspend-time hint answers ~30 cycles sooner (tools/pingpong.py)
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(addr), "r"(parity)
                : "memory");
        if (done) return;
        if (wd && (++tries & 63u) == 0 && mbar_watchdog(wd, tries, tag, info, parity, addr, prog)) return;
    }
}
template <bool HINT = true>
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity, unsigned int *wd = nullptr, uint32_t tag = 0,
                                          uint32_t info = 0, volatile uint32_t *prog = nullptr)
{
    mbar_wait_addr<HINT>(smem_u32(bar), parity, wd, tag, info, prog);
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

// The issuing WARP walks its role converged (every value it computes is warp-uniform, so ptxas keeps descriptors and
// loop state in uniform registers: no R2UR in the issue sequence); the tcgen05 instructions themselves are guarded by the
// `leader` flag from elect.sync, which ptxas turns into ONE uniform-datapath instruction per warp (UTCIMMA / UTCHMMA /
// UTCBAR).
__device__ __forceinline__ uint32_t elect_flag()
{
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred;
}
__device__ __forceinline__ void tc_commit(uint32_t leader, uint32_t bar_addr)
{
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "setp.ne.b32 q, %1, 0;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar_addr),
        "r"(leader)
        : "memory");
}
// nothing is issued when guard == 0
__device__ __forceinline__ void tc_mma_i8_if(uint32_t guard, uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.ne.b32 q, %5, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(guard)
        : "memory");
}
// D = A*B (accumulate off): the fp16 MMA opens every output row and overwrites the previous row's result
__device__ __forceinline__ void tc_mma_f16_overwrite_if(uint32_t guard, uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc)
{
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.b32 p, 0, 0;\n\t"
        "setp.ne.b32 q, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(guard)
        : "memory");
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
// tcgen05.wait::ld ordered against the USES of v through register dependencies (no memory clobber)
__device__ __forceinline__ void tc_wait_ld32(uint32_t (&v)[32])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                   "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
                   "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
                   "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31]));
}
// 16-column variant (the tail of a warp's column range): fills v[0..15]
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
// zero 16 columns of this warp's lane quarter
__device__ __forceinline__ void tc_st16_zero(uint32_t taddr)
{
    const uint32_t z = 0u;
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(z)
        : "memory");
}
__device__ __forceinline__ uint32_t tc_ld1(uint32_t taddr)
{
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
    return v;
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Experiment hooks (role time budget, FOCR_TC_DBG modes, progress words) cost instructions in loops whose speed is set
// by the instruction count of ONE thread, so they are compiled in only with -DFOCR_TC_EXPERIMENTS.
#ifdef FOCR_TC_EXPERIMENTS
constexpr bool TC_EXP = true;
#else
constexpr bool TC_EXP = false;
#endif
#define TC_PROG(slot, value)                  \
    do {                                      \
        if (TC_EXP) prog_[(slot)] = (value);  \
    } while (0)

// role time budget (timing experiments): TT(slot, stmt) runs stmt and, when tracing, adds its cycles to a counter
#define TT(k, ...)                                  \
    do {                                            \
        if (tron) {                                 \
            const long long t0_ = clock64();        \
            __VA_ARGS__;                            \
            tacc[k] += clock64() - t0_;             \
        } else {                                    \
            __VA_ARGS__;                            \
        }                                           \
    } while (0)
#define TT_BEGIN() const bool tron = TC_EXP && p.trace != nullptr && blockIdx.x == 0 && !p.timeline; long long tacc[6] = {0, 0, 0, 0, 0, 0}; const long long tstart_ = tron ? clock64() : 0
// event timeline (env FOCR_TC_TIMELINE=first job): CTA 0 stamps clock64 at event `ev` (0..3) of jobs [tl0, tl0+256) per role
constexpr int TC_TL_JOBS = 256;
#define TL(role, job, ev)                                                                                          \
    do {                                                                                                           \
        if (TC_EXP && p.timeline && blockIdx.x == 0 && (uint32_t)((job) - p.tl0) < (uint32_t)TC_TL_JOBS)         \
            p.trace[64 + (((role) * TC_TL_JOBS + ((job) - p.tl0)) * 4 + (ev))] = clock64();                       \
    } while (0)
#define TT_END(base)                                                         \
    do {                                                                     \
        if (tron) {                                                          \
            p.trace[(base)] = clock64() - tstart_;                           \
            for (int i_ = 0; i_ < 6; i_++) p.trace[(base) + 1 + i_] = tacc[i_]; \
        }                                                                    \
    } while (0)

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
    it.ys0 = 1 + ys * p.yseg;  // ncc.cpp:98: the scan starts at y = 1
    it.ys1 = min(it.ys0 + p.yseg, p.r_h - p.n_h + 1);
    return true;
}

// Survivors of the tensor-core screen go to a candidate list that is PRIVATE to the epilogue warp: slots
// are handed out with a ballot/popc prefix, so the hot kernel has no atomics at all (a global atomic with
// a return value costs the warp ~1 us, and ~12 % of all 32-column units contain a survivor).
__device__ __forceinline__ void append_candidates(Hit *list, uint32_t cap, uint32_t &count, uint32_t mask, uint32_t col0,
                                                  int page, int gx, int y)
{
    const unsigned lane_lt = (1u << (threadIdx.x & 31)) - 1u;
    uint32_t any = __reduce_or_sync(0xffffffffu, mask);
    while (any) {  // warp-uniform loop over the (few) columns in which some lane has a survivor
        const int j = __ffs(any) - 1;
        any &= any - 1;
        const bool mine = (mask >> j) & 1u;
        const unsigned vote = __ballot_sync(0xffffffffu, mine);
        if (mine) {
            const uint32_t slot = count + __popc(vote & lane_lt);
            if (slot < cap) {
                Hit h;
                h.t = col0 + j;
                h.yx = ((uint32_t)y << 16) | (uint32_t)gx;
                h.sim = 0.f;
                h.page = page;
                list[slot] = h;
            }
        }
        count += __popc(vote);
    }
}

// ---------------------------------------------------------------------------------------------- MMA issuers
struct TcSmem {   // shared-memory addresses (shared window, bytes) of the operands and barriers the issuing warp touches
    uint32_t btile, ring, a2ring, b2tile;
    uint32_t bar_btile, a_full, a_empty, a2_full, a2_empty, t_full, t_empty;
    volatile uint32_t *prog;
};

// TWO issuing warps (warps 1 and 2): warp `mw` issues the jobs whose index has parity mw; job k (output row k / nsub,
// sub-block k % nsub) goes to accumulator k mod nbuf and is drained by epilogue team k & 1.  tcgen05.mma blocks its issuer
// at the rate of the tensor pipe (measured: the 8 MMAs of a 160-column job take ~690 cycles to ISSUE, there is next to no
// queue), so whatever else an issuer does -- barrier waits (~100+ cycles even when complete), commits, ring hand-backs --
// is dead time for the pipe unless ANOTHER warp is issuing meanwhile.  The two warps touch different accumulators, so
// the order in which the pipe takes their MMAs is free.
// Barriers: every accumulator has one "full" and one "empty" barrier PER PIPELINE: t_full[acc][i] is committed by issuer
// i and waited on by team i; t_empty[acc][i] collects the arrivals of the team that drains the job which issuer i will
// overwrite (the job nbuf earlier).  Each barrier has exactly one waiter that sees every one of its phases, whatever
// the ring length (with an odd ring a shared barrier would be waited on by alternating parties, which can fall two phases
// behind -- a parity wait then aliases).  A waiter keeps the parity of its next phase per accumulator in a bit mask.
// A ring group goes back to its producer when BOTH warps have moved past it: each commits to the group's "empty" barrier
// (count 2) when it reaches a job that no longer reads the group -- tcgen05.commit only tracks the committing thread's MMAs.
// The warps run CONVERGED and everything they compute depends only on kernel parameters and loop counters, so ptxas keeps
// the whole role in uniform registers (descriptor arithmetic = UIADD3, no R2UR); `leader` (elect.sync) guards the
// tcgen05 instructions, which are uniform-datapath instructions executed once per warp.
// Because the ring stores its first n_hp-1 slots twice, the slots of an output row are consecutive (no wrap test fires
// inside a row when np == 16).
// LEAN = no parity probe (the common case): the fp16 MMA is always issued, no run-time switches in the issue sequence.
// MW = which of the two issuing warps, a compile-time constant (a run-time warp index would sit in a vector register and
// drag the barrier addresses and the job-ownership branch out of the uniform datapath).
template <bool LEAN, uint32_t MW>
__device__ __forceinline__ void tc_mma_role(const TcParams &p, const TcSmem &sm, const uint32_t leader)
{
    constexpr uint32_t mw = MW;
    // N of this warp's MMAs: with two sub-blocks per row the jobs of parity mw ARE sub-block mw; the real columns of a
    // sub-block rounded up to 16 (not to the 32 of the epilogue's units: an MMA takes N / 2 cycles)
    const uint32_t nmma = (uint32_t)(p.nsub == 2 ? p.nmma[mw] : p.nmma[0]);
    const uint32_t idesc8 = (2u << 4)                          // D format: S32
                            | (0u << 7) | (0u << 10)           // A, B: unsigned 8-bit
                            | (0u << 15) | (0u << 16)          // A, B: K-major
                            | ((nmma >> 3) << 17)              // N
                            | ((128u >> 4) << 24);             // M = 128
    const uint32_t idesc16 = (1u << 4)                         // D format: F32 (same TMEM columns, read as fp32)
                             | (0u << 7) | (0u << 10)          // A, B: F16
                             | ((nmma >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t ring_n = p.ring, ring_g = p.ring_groups, nbuf = p.nbuf, ksteps = p.ksteps, n_hp = p.n_hp, nbs = p.nbs,
                   nsub = p.nsub;
    const uint32_t b_lbo16 = (uint32_t)p.nbs, b_inc = 2 * b_lbo16;      // 16-byte units: K chunks of B are nbs columns apart
    const uint32_t b_sub = 2 * ksteps * b_lbo16, b2_sub = 2 * b_lbo16;   // sub-block strides of the B and B2 tiles
    // the two 16-byte K chunks of an MMA: ring slots s, s+1 (np == 16), s, s+2 (packed: a slot already holds two page rows), or
    // entries m, m+16 of one slot (np == 32)
    const uint32_t a_lbo16 = (p.np == 16 ? (uint32_t)p.row_pitch * (p.packed ? 2u : 1u) : 256u) >> 4;
    const uint32_t pitch16 = (uint32_t)p.row_pitch >> 4;
    const uint32_t a_step = p.np == 16 ? (p.packed ? 4u : 2u) : 1u;   // ring slots consumed per K step
    const uint64_t desc_hi = (uint64_t)((128u >> 4) | (1u << 14)) << 32;  // SBO = 128 B, version = 1
    const uint32_t a_lo0 = ((sm.ring & 0x3FFFFu) >> 4) | (a_lbo16 << 16);
    const uint32_t b_lo0 = ((sm.btile & 0x3FFFFu) >> 4) | (b_lbo16 << 16);
    const uint32_t a_inc = a_step * pitch16, a_wrap = ring_n * pitch16, a_end = a_lo0 + a_wrap;
    // A2 rows: one 2 KB block of 128 x 16 B per box size; the fp16 MMA's two K chunks read the two blocks (one box
    // size: both chunks read the same block, LBO = 0, and the second chunk of B2 is all zeros)
    const uint32_t a2_slot16 = (uint32_t)p.a2_slot >> 4;
    const uint32_t a2_lbo = (p.ncls == 2 ? (2048u >> 4) : 0u) << 16;
    const uint32_t a2_addr16 = (sm.a2ring & 0x3FFFFu) >> 4, a2_wrap16 = (uint32_t)p.a2_groups * TC_G * a2_slot16;
    const uint32_t a2_end16 = a2_addr16 + a2_wrap16;
    const uint32_t b2_lo = ((sm.b2tile & 0x3FFFFu) >> 4) | (b_lbo16 << 16);
    const bool corr = LEAN || p.dbg_acc == nullptr;
    const uint32_t a2_groups = p.a2_groups;
    const bool wrap = p.n_mirror == 0;       // no mirror slots: a row's K steps may run past the end of the ring
    const uint32_t mma_on = (TC_EXP && (p.dbg_mode & 4)) ? 0u : leader;
    const uint32_t f16_on = corr ? mma_on : 0u;
    const uint32_t t_full = sm.t_full + 8 * mw, t_empty = sm.t_empty + 8 * mw;   // this pipeline's barriers: [acc][2]
    volatile uint32_t *const prog_ = sm.prog;
    mbar_wait_addr(sm.bar_btile, 0, p.wd, 10 + mw, 0);
    TT_BEGIN();
    uint32_t o = 0;                          // global output-row index of the current row
    uint32_t g = 0;                          // global page-row index of that row's first page row
    uint32_t a_first = a_lo0;                // A descriptor (low word) of that page row's ring slot
    uint32_t o_slot16 = a2_addr16;           // A2 ring slot address (>> 4) of that output row
    uint32_t rel_g = 0, rel_rows = TC_G;     // page-row groups: next to hand back / rows covered once it is
    uint32_t new_g = 0, new_par = 0, rows_ready = 0;
    uint32_t rel2_g = 0, rel2_rows = TC_G, new2_g = 0, new2_par = 0, rows2_ready = 0;
    uint32_t job = 0, acc = 0;               // the current job and its accumulator (job mod nbuf)
    uint32_t empty_par = 0;                  // bit a: parity of this warp's next phase of t_empty[a][mw]
    Item it;
    for (int idx = blockIdx.x; get_item(p, idx, it); idx += gridDim.x) {
        const uint32_t n_out_rows = it.ys1 - it.ys0;
        for (uint32_t j = 0; j < n_out_rows; j++) {
            // one job per sub-block: the row's operands are shared, the templates (B, B2) and the accumulator differ
            uint32_t b_lo = b_lo0, b2 = b2_lo;
            for (uint32_t sb = 0; sb < nsub; sb++, b_lo += b_sub, b2 += b2_sub, job++) {
                if ((job & 1u) == mw) {
                    TC_PROG(1 + mw, (1u << 24) | o);
                    // ring groups that lie entirely below this row are no longer read by this warp
                    while (rel_rows <= g) {
                        TT(4, tc_commit(leader, sm.a_empty + 8 * rel_g));
                        rel_rows += TC_G;
                        if (++rel_g == ring_g) rel_g = 0;
                    }
                    while (rel2_rows <= o) {
                        TT(4, tc_commit(leader, sm.a2_empty + 8 * rel2_g));
                        rel2_rows += TC_G;
                        if (++rel2_g == a2_groups) rel2_g = 0;
                    }
                    // operands: page rows g .. g+n_hp-1 and the A2 row of this output
                    while (rows_ready < g + n_hp) {
                        TT(0, mbar_wait_addr(sm.a_full + 8 * new_g, new_par, p.wd, 12 + mw, o, sm.prog));
                        rows_ready += TC_G;
                        if (++new_g == ring_g) new_g = 0, new_par ^= 1;
                    }
                    while (corr && rows2_ready <= o) {
                        TT(1, mbar_wait_addr(sm.a2_full + 8 * new2_g, new2_par, p.wd, 14 + mw, o, sm.prog));
                        rows2_ready += TC_G;
                        if (++new2_g == a2_groups) new2_g = 0, new2_par ^= 1;
                    }
                    const uint32_t d0 = acc * nbs;
                    TC_PROG(1 + mw, (2u << 24) | o);
                    if (leader) TL(mw, job, 0);
                    if (job >= nbuf) {   // the accumulator's previous job has been drained
                        TT(2, mbar_wait_addr<false>(t_empty + 16 * acc, (empty_par >> acc) & 1u, p.wd, 16 + mw, o, sm.prog));
                        empty_par ^= 1u << acc;
                    }
                    tc_fence_after();
                    if (leader) TL(mw, job, 1);
                    const long long ti_ = tron ? clock64() : 0;
                    // F = A2 . B2^T in fp32 (accumulate off): K = 16 fp16 = two 16-byte chunks, one per box size
                    tc_mma_f16_overwrite_if(f16_on, d0, desc_hi | (o_slot16 | a2_lbo), desc_hi | b2, idesc16);
                    // K steps: consecutive ring slots, consecutive B chunks.  A rolled loop: in uniform registers a step is two
                    // UIADD3, a compare and the UTCIMMA, far below the ~80 cycles the pipe takes per MMA (and a few small
                    // instantiations keep ptxas' uniform-register allocation intact; four unrolled variants did not).
                    {
                        uint32_t al = a_first, bl = b_lo;
#pragma unroll 1
                        for (uint32_t k = 0; k < ksteps; k++) {
                            tc_mma_i8_if(mma_on, d0, desc_hi | al, desc_hi | bl, idesc8, (corr || k) ? 1u : 0u);
                            al += a_inc;
                            bl += b_inc;
                            if (wrap && al >= a_end) al -= a_wrap;
                        }
                    }
                    if (tron) tacc[3] += clock64() - ti_;
                    if (leader) TL(mw, job, 2);
                    TT(4, tc_commit(leader, t_full + 16 * acc));   // accumulator ready for this pipeline's epilogue team
                    if (leader) TL(mw, job, 3);
                }
                if (++acc == nbuf) acc = 0;
            }
            // next row
            o++;
            g++;
            o_slot16 += a2_slot16;
            if (o_slot16 >= a2_end16) o_slot16 -= a2_wrap16;
            a_first += pitch16;
            if (a_first >= a_end) a_first -= a_wrap;
            if (j + 1 == n_out_rows) {   // next item: its first row follows this item's last page row
                g += n_hp - 1;
                a_first += (n_hp - 1) * pitch16;
                if (a_first >= a_end) a_first -= a_wrap;
            }
        }
    }
    TC_PROG(1 + mw, 9u << 24);
    // groups this warp never passed explicitly (the last rows): nothing waits for them
    if (leader && mw == 0) TT_END(0);
}

// NUNITS = 32-column units per accumulator (p.nunits) as a compile-time constant: the epilogue warps run a latency-bound
// instruction stream (two of them per SM sub-partition), and every run-time "is there a unit b" test in it costs a
// constant load, a compare, a branch and often an instruction-cache miss (ncu: ~400 instructions per job with run-time
// tests against ~150 without).
template <int NUNITS>
__global__ void __launch_bounds__(TC_THREADS, 1) scan_tc_kernel(const __grid_constant__ TcParams p)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    // ---- shared memory carve-up (all blocks multiples of 128 B)
    uint8_t *btile = smem;
    uint8_t *ring = btile + ((p.btile_bytes + 127) & ~127u);
    uint8_t *raw = ring + (size_t)(p.ring + p.n_mirror) * p.row_pitch;  // + mirror slots: slot ring+i repeats slot i
    uint8_t *a2ring = raw + ((TC_RAW_SLOTS * TC_RAW_BYTES + 127) & ~127);  // [a2_groups*4][ncls][128][16 B] fp16 x 8
    uint8_t *b2tile = a2ring + (size_t)p.a2_groups * TC_G * p.a2_slot;    // [nsub][2][nbs][16 B]: one K chunk per box size
    uint64_t *bars = (uint64_t *)(b2tile + (size_t)2 * p.nb * 16);   // b2tile: [nsub][2][nbs][16]
    uint64_t *bar_btile = bars;                       // 1
    uint64_t *raw_full = bars + 1;                    // TC_RAW_GROUPS
    uint64_t *raw_empty = raw_full + TC_RAW_GROUPS;   // TC_RAW_GROUPS
    uint64_t *a_full = raw_empty + TC_RAW_GROUPS;     // TC_RING_MAX
    uint64_t *a_empty = a_full + TC_RING_MAX;         // TC_RING_MAX
    uint64_t *a2_full = a_empty + TC_RING_MAX;        // TC_A2_GROUPS
    uint64_t *a2_empty = a2_full + TC_A2_GROUPS;      // TC_A2_GROUPS
    uint64_t *t_full = a2_empty + TC_A2_GROUPS;       // [TC_MAX_BUF][2]: per accumulator and pipeline (issuing warp / epilogue team)
    uint64_t *t_empty = t_full + 2 * TC_MAX_BUF;      // [TC_MAX_BUF][2]
    uint32_t *tmem_ptr = (uint32_t *)(t_empty + 2 * TC_MAX_BUF);
    volatile uint32_t *prog = (volatile uint32_t *)(tmem_ptr + 2);   // [32] per-warp progress (debugging aid, see mbar_wait)
    // (73 barriers + tmem_ptr[2] + prog[32] = 720 bytes after a 128-byte aligned start: the scratch is 16-byte aligned)
    uint8_t *epi_scratch = (uint8_t *)(tmem_ptr + 2 + 32);   // [TC_EPI_WARPS][128 B] candidate extraction
    if (threadIdx.x < 32) prog[threadIdx.x] = 0;
    volatile uint32_t *const prog_ = prog;

    // warp index through a broadcast: ptxas then knows it is warp-uniform, so the role branches are uniform branches and the
    // issuing warp can keep its state in uniform registers (see tc_mma_role)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

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
        for (int i = 0; i < TC_A2_GROUPS; i++) {
            mbar_init(a2_full + i, 4);     // one arrival per A2 warp
            mbar_init(a2_empty + i, 2);
        }
        for (int i = 0; i < 2 * TC_MAX_BUF; i++) {
            mbar_init(t_full + i, 1);
            mbar_init(t_empty + i, TC_EPI_WARPS / 2);  // one arrival per epilogue warp (lane quarter) of the team that drains the job
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // B2 tile: B2[t] = {-b1, -b2, -b1, -a1, -a2, -a1, -BIG, 510 | -BIG} as fp16 hi/lo splits of b = s_n/n and
    // a = thr*norm_n, in the K chunk of the column's box size (the other chunk is zero); +inf norm marks padding /
    // constant templates (last entry -BIG: never a candidate)
    for (int t = threadIdx.x; t < p.nb; t += TC_THREADS) {
        const float4 c = p.colconst[t];
        const bool pad = !(c.x < __int_as_float(0x7f800000));
        const float a = pad ? 0.f : p.thr * c.x, b = pad ? 0.f : c.y;
        const __half a1 = __float2half_rn(a), b1 = __float2half_rn(b);
        const __half a2 = __float2half_rn(a - __half2float(a1)), b2 = __float2half_rn(b - __half2float(b1));
        const __half big = __float2half_rn(-TC_BIG), padh = __float2half_rn(pad ? -TC_BIG : 510.f);
        // sshift > 0: the A2 rows carry S_hi/16 (S may exceed the fp16 range), so the factors of S_hi carry the 16
        const __half b1h = p.sshift ? __float2half_rn(16.f * __half2float(b1)) : b1;
        const __half b2h = p.sshift ? __float2half_rn(16.f * __half2float(b2)) : b2;
        __align__(16) __half h[8] = {__hneg(b1h), __hneg(b2h), __hneg(b1), __hneg(a1), __hneg(a2), __hneg(a1), big, padh};
        const int sb = t / p.nbs, n = t - sb * p.nbs;      // layout [sub][K chunk][column][16 B]
        const int chunk = (!pad && c.z != 0.f) ? 1 : 0;    // second box size -> second K chunk
        uint8_t *dst = b2tile + ((size_t)sb * 2 * p.nbs + n) * 16;
        *(uint4 *)(dst + (size_t)chunk * p.nbs * 16) = *(const uint4 *)h;
        *(uint4 *)(dst + (size_t)(1 - chunk) * p.nbs * 16) = make_uint4(0, 0, 0, 0);
    }
    if (warp == 3) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)),
                     "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();  // zero blocks / B2 were written through the generic proxy, the MMA reads them through the async proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // the CTA owns all 512 TMEM columns of its SM: the allocation starts at lane 0, column 0.  Using the literal
    // keeps every TMEM address warp-uniform for the compiler.
    if (*tmem_ptr != 0) __trap();
    constexpr uint32_t tmem_base = 0;

    // rows this CTA streams through the pipeline (all its items)
    uint32_t total_rows = 0, total_out = 0;
    if (warp >= 4 && warp < 12) {
        Item it;
        for (int idx = blockIdx.x; get_item(p, idx, it); idx += gridDim.x) {
            total_rows += (it.ys1 - it.ys0) + p.n_hp - 1;
            total_out += it.ys1 - it.ys0;
        }
    }

    // Register re-partition (setmaxnreg, warpgroup granular, first statement of a role): the Toeplitz and A2
    // warps give registers away, the epilogue warps, which hold two 32-column units at a time, take them.
    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TC_REGS_ISSUE));
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
        TT_BEGIN();
        Item it;
        for (int idx = blockIdx.x; get_item(p, idx, it); idx += gridDim.x) {
            const uint8_t *src = p.inv + (size_t)it.page * p.inv_page_stride + it.x0 + (size_t)it.ys0 * p.pitch;
            const int n_rows = (it.ys1 - it.ys0) + p.n_hp - 1;
            for (int r = 0; r < n_rows; r++, src += p.pitch) {
                if (in_group == 0 && !first_round) TT(0, mbar_wait(raw_empty + rg, rgpar, p.wd, 1, slot, prog));
                if (elect_one()) {
                    mbar_expect_tx(raw_full + rg, p.packed ? 2 * row_bytes : row_bytes);
                    tma_bulk_g2s(raw + slot * TC_RAW_BYTES, src, row_bytes, raw_full + rg);
                    // packed mode: the slot of page row r also needs row r+1 (PAGE_PAD_ROWS zero rows follow the page)
                    if (p.packed) tma_bulk_g2s(raw + slot * TC_RAW_BYTES + 144, src + p.pitch, row_bytes, raw_full + rg);
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
        if (lane == 0) TT_END(32);
    } else if (warp == 1 || warp == 2) {
        // ================================================================== MMA issuers (tc_mma_role): whole warps, converged
        const uint32_t sbase = smem_u32(smem);
        const TcSmem sm = {sbase + (uint32_t)(btile - smem), sbase + (uint32_t)(ring - smem), sbase + (uint32_t)(a2ring - smem),
                           sbase + (uint32_t)(b2tile - smem), sbase + (uint32_t)((uint8_t *)bar_btile - smem),
                           sbase + (uint32_t)((uint8_t *)a_full - smem), sbase + (uint32_t)((uint8_t *)a_empty - smem),
                           sbase + (uint32_t)((uint8_t *)a2_full - smem), sbase + (uint32_t)((uint8_t *)a2_empty - smem),
                           sbase + (uint32_t)((uint8_t *)t_full - smem), sbase + (uint32_t)((uint8_t *)t_empty - smem), prog};
        const uint32_t leader = elect_flag();
        const bool lean = p.dbg_acc == nullptr;
        if (warp == 1) {
            if (lean) tc_mma_role<true, 0>(p, sm, leader);
            else tc_mma_role<false, 0>(p, sm, leader);
        } else {
            if (lean) tc_mma_role<true, 1>(p, sm, leader);
            else tc_mma_role<false, 1>(p, sm, leader);
        }
        __syncwarp();
    }
    } else if (warp < 8) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TC_REGS_TOEPLITZ));
        // ================================================================== Toeplitz expansion
        // one warp per page row, four rows (one group) in flight per handshake
        const int w = warp - 4;
        const uint32_t n_mirror = p.n_mirror;   // slots 0 .. n_mirror-1 are stored twice, so an output row never wraps
        const uint32_t ring_g = p.ring_groups;
        uint32_t rg = 0, rgpar = 0, ag = 0, agpar = 1;
        bool first_round = true;
        TT_BEGIN();
        for (uint32_t g0 = 0; g0 < total_rows; g0 += TC_G) {
            if (lane == 0) TC_PROG(warp, g0);
            if (w == 0) {  // one warp polls, the other three wait on a named barrier (see the epilogue)
                TT(0, mbar_wait(raw_full + rg, rgpar, p.wd, 2, g0, prog));
                if (!first_round) TT(1, mbar_wait(a_empty + ag, agpar, p.wd, 3, g0, prog));
            }
            asm volatile("bar.sync 3, 128;" ::: "memory");
            if (g0 + w < total_rows && !(TC_EXP && (p.dbg_mode & 16))) {
                const uint32_t *rw = (const uint32_t *)(raw + (rg * TC_G + w) * TC_RAW_BYTES);
                const uint32_t s = ag * TC_G + w;
                uint8_t *dst = ring + (size_t)s * p.row_pitch;
                for (int ee = lane; ee < p.n_entries; ee += 32) {
                    const uint32_t *wp = rw + (ee >> 2);
                    const int sh = (ee & 3) * 8;
                    uint4 o4;
                    if (p.packed) {   // 8 bytes of this page row, 8 bytes of the next one (second half of the raw slot)
                        const uint32_t *wq = wp + 36;
                        const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2], v0 = wq[0], v1 = wq[1], v2 = wq[2];
                        o4.x = __funnelshift_r(w0, w1, sh);
                        o4.y = __funnelshift_r(w1, w2, sh);
                        o4.z = __funnelshift_r(v0, v1, sh);
                        o4.w = __funnelshift_r(v1, v2, sh);
                    } else {
                        const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2], w3 = wp[3], w4 = wp[4];
                        o4.x = __funnelshift_r(w0, w1, sh);
                        o4.y = __funnelshift_r(w1, w2, sh);
                        o4.z = __funnelshift_r(w2, w3, sh);
                        o4.w = __funnelshift_r(w3, w4, sh);
                    }
                    *(uint4 *)(dst + ee * 16) = o4;
                    if (s < n_mirror) *(uint4 *)(dst + (size_t)p.ring * p.row_pitch + ee * 16) = o4;
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
        if (w == 0 && lane == 0) TT_END(16);
    } else if (warp < 12) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TC_REGS_A2));
        // ================================================================== A2 rows: window statistics -> fp16 operand
        // One warp per output row, four rows per handshake.  For window m of output row y and each box size:
        //   A2 = {S_hi, S_hi, S_lo, P_1, P_1, P_2, V, 2^15}   (S = s_p split at bit 6; P = norm_p split hi/lo)
        // V = BIG when the window can never hit: x outside [1, r_w-n_w] (ncc.rs:281) or a constant window
        // (norm_p = +inf marker: rnorm_p = inf in the reference, ncc.cpp:216-220).  V = -BIG (always a
        // candidate, S = P = 0) when b_max*S + a_max*P could push F below 2^23, where its bits stop being linear.
        const int w = warp - 8;
        uint32_t ag = 0, agpar = 1;
        bool first_round = true;
        int cur_idx = blockIdx.x;
        uint32_t item_o0 = 0;  // global output index of the current item's first row
        Item it;
        bool have = get_item(p, cur_idx, it);
        TT_BEGIN();
        for (uint32_t o0 = 0; o0 < total_out; o0 += TC_G) {
            if (lane == 0) TC_PROG(warp, o0);
            if (!first_round) {
                if (w == 0) TT(0, mbar_wait(a2_empty + ag, agpar, p.wd, 4, o0, prog));
                asm volatile("bar.sync 4, 128;" ::: "memory");
            }
            const uint32_t o = o0 + w;
            if (o < total_out) {
                while (have && o >= item_o0 + (uint32_t)(it.ys1 - it.ys0)) {  // advance to the item that owns row o
                    item_o0 += it.ys1 - it.ys0;
                    cur_idx += gridDim.x;
                    have = get_item(p, cur_idx, it);
                }
                const int y = it.ys0 + (int)(o - item_o0);
                const size_t rowoff = (size_t)it.page * p.plane_page_stride + (size_t)y * p.spitch + it.x0;
                // ALL global loads of the row first (both box sizes: up to 16 in flight per thread): the role is bound by their
                // latency, one round trip per row instead of one per box size (or, with a branch in between, per load)
                uint32_t wv[2][4];
                float pw[2][4];
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    if (c == 1 && p.ncls != 2) break;
                    const uint32_t *spc = p.sp[c] + rowoff;
                    const int x_last = p.r_w - (c ? p.n_w2 : p.n_w);
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const int m = lane + 32 * i, gx = it.x0 + m;
                        const bool ok = gx >= 1 && gx <= x_last && !(TC_EXP && (p.dbg_mode & 8));
                        wv[c][i] = ok ? __ldg(spc + m) : (p.pack ? 0xFFFF0000u : 0u);
                    }
                }
                if (!p.pack) {
#pragma unroll
                    for (int c = 0; c < 2; c++) {
                        if (c == 1 && p.ncls != 2) break;
                        const float *pfc = p.pf[c] + rowoff;
                        const int x_last = p.r_w - (c ? p.n_w2 : p.n_w);
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            const int m = lane + 32 * i, gx = it.x0 + m;
                            const bool ok = gx >= 1 && gx <= x_last && !(TC_EXP && (p.dbg_mode & 8));
                            pw[c][i] = ok ? __ldg(pfc + m) : __int_as_float(0x7f800000);
                        }
                    }
                }
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    if (c == 1 && p.ncls != 2) break;
                    uint8_t *dst = a2ring + (size_t)(ag * TC_G + w) * p.a2_slot + c * 2048;
                    const float bmax = p.bmax[c], amax = p.amax[c];
                    uint32_t sv[4];
                    float pv[4];
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        if (p.pack) {   // one word per window: s_p in the low half, norm_p in 11.5 fixed point in the high half
                            sv[i] = wv[c][i] & 0xFFFFu;
                            pv[i] = (wv[c][i] >> 16) == 0xFFFFu ? __int_as_float(0x7f800000) : (float)(wv[c][i] >> 16) * 0.03125f;
                        } else {
                            sv[i] = wv[c][i];
                            pv[i] = pw[c][i];
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 4 && !(TC_EXP && (p.dbg_mode & 32)); i++) {
                        const int m = lane + 32 * i;
                        const bool valid = pv[i] < __int_as_float(0x7f800000);
                        const bool lin = valid && (bmax * (float)sv[i] + amax * pv[i] <= TC_KCAP);
                        const uint32_t s = lin ? sv[i] : 0u;
                        const float P = lin ? pv[i] : 0.f;
                        // S = S_hi + S_lo, both exact in fp16: split at bit 6 (S < 2^16), or at bit 10 with S_hi stored /16 for
                        // boxes of more than 256 pixels (S < 2^19); the factor 16 is in B2
                        const float s_hi = p.sshift ? (float)((s & ~1023u) >> 4) : (float)(s & ~63u);
                        const float s_lo = p.sshift ? (float)(s & 1023u) : (float)(s & 63u);
                        const __half p1 = __float2half_rn(P);
                        const __half p2 = __float2half_rn(P - __half2float(p1));
                        const __half shi = __float2half_rn(s_hi), slo = __float2half_rn(s_lo);
                        const __half v = __float2half_rn(valid ? (lin ? 0.f : -TC_BIG) : TC_BIG), k = __float2half_rn(32768.f);
                        __align__(16) __half h[8] = {shi, shi, slo, p1, p1, p2, v, k};
                        *(uint4 *)(dst + m * 16) = *(const uint4 *)h;
                    }
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(a2_full + ag);
            if (++ag == (uint32_t)p.a2_groups) ag = 0, agpar ^= 1, first_round = false;
        }
        if (w == 0 && lane == 0) TT_END(24);
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(TC_REGS_EPILOGUE));
        // ================================================================== epilogue (8 warps = 2 teams x 4 lane quarters)
        // The accumulator holds the bits of fp32 C0 + d: one 3-input max per column pair (a tree, so the
        // FMNMX3s are independent), one vote per 32 columns.  Jobs alternate between two TEAMS of 4 warps (one per TMEM
        // lane quarter); a warp takes ALL columns of its job's lane quarter, up to four 32-column units in registers at
        // once, and hands the accumulator back to the MMA warps as soon as its last unit has landed -- before it screens
        // (jobs of more than four units: the first ones are screened while the last ones load).  The warps of a team
        // do not synchronise with each other: each polls the job's "full" barrier itself, so a warp that is busy
        // extracting candidates delays nobody else.
        const int e = warp - 12;
        const int q = e & 3;                      // TMEM lane quarter this warp may access (warp % 4)
        const int team = e >> 2;                  // which jobs
        const uint32_t nbuf = p.nbuf, nsub = p.nsub;
        constexpr int nunits = NUNITS, n2 = nunits - TC_EPI_UNITS;   // n2 > 0: units screened before the release
        const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
        const float T = p.dbg_acc ? 3.0e38f : TC_C0 - TC_MARGIN;
        uint32_t acc = 0;                         // the next job's accumulator (as in tc_mma_role)
        uint32_t job = 0;
        uint32_t full_par = 0;                    // bit a: parity of this team's next phase of t_full[a][team]
        const uint32_t next_pipe = (uint32_t)(team + (int)p.nbuf) & 1u;   // pipeline of the job that overwrites an accumulator this team drains
        Hit *my_list = p.cands + (size_t)(blockIdx.x * TC_LISTS_PER_CTA + e) * p.cand_cap;
        uint32_t my_count = 0;
        float *const scratch = (float *)(epi_scratch + e * 128);   // one lane's 32 values at a time (candidate extraction)
        const unsigned lane_lt = (1u << lane) - 1u;
        // columns of the accumulators that no MMA of this launch writes (N is rounded up to 16, the units to 32): zero them
        // once, or whatever an earlier kernel left there would be screened row after row
        {
            const int n_min = min(p.nmma[0], p.nsub == 2 ? p.nmma[1] : p.nmma[0]);
            for (uint32_t a = 0; a < p.nbuf; a++)
                for (int c0 = n_min; c0 < p.nbs; c0 += 16) tc_st16_zero(tlane + a * p.nbs + c0);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        const bool no_ld = TC_EXP && (p.dbg_mode & 1), no_screen = TC_EXP && (p.dbg_mode & 2);
        const int dbg_sub = p.dbg_acc ? p.dbg_col / p.nbs : -1, dbg_c = p.dbg_acc ? p.dbg_col % p.nbs : 0;
        TT_BEGIN();
        Item it;
        for (int idx = blockIdx.x; get_item(p, idx, it); idx += gridDim.x) {
            const int gx0 = it.x0 + q * 32;       // window of lane 0
            for (int y = it.ys0; y < it.ys1; y++) {
              for (uint32_t sb = 0; sb < nsub; sb++, job++) {
                if ((job & 1u) == (uint32_t)team) {
                if (lane == 0) TC_PROG(warp, job);
                TT(0, mbar_wait<false>(t_full + 2 * acc + team, (full_par >> acc) & 1u, p.wd, 20 + team, y, prog));
                full_par ^= 1u << acc;
                tc_fence_after();
                const int tlrole = e == 0 ? 2 : e == 4 ? 3 : 0;
                if (tlrole && lane == 0) TL(tlrole, job, 0);
                const long long tseen_ = tron ? clock64() : 0;
                const uint32_t tb = tlane + acc * p.nbs;
                const uint32_t cbase = p.col_base + sb * p.nbs;
                auto screen = [&](uint32_t (&v)[32], int u) {
                    const long long ts_ = tron ? clock64() : 0;
                    float t[11];
#pragma unroll
                    for (int j = 0; j < 10; j++)
                        asm("max.f32 %0, %1, %2, %3;"
                            : "=f"(t[j])
                            : "f"(__uint_as_float(v[3 * j])), "f"(__uint_as_float(v[3 * j + 1])), "f"(__uint_as_float(v[3 * j + 2])));
                    t[10] = fmaxf(__uint_as_float(v[30]), __uint_as_float(v[31]));
                    float a, b, c, mx;
                    asm("max.f32 %0, %1, %2, %3;" : "=f"(a) : "f"(t[0]), "f"(t[1]), "f"(t[2]));
                    asm("max.f32 %0, %1, %2, %3;" : "=f"(b) : "f"(t[3]), "f"(t[4]), "f"(t[5]));
                    asm("max.f32 %0, %1, %2, %3;" : "=f"(c) : "f"(t[6]), "f"(t[7]), "f"(t[8]));
                    asm("max.f32 %0, %1, %2, %3;" : "=f"(mx) : "f"(t[9]), "f"(t[10]), "f"(a));
                    asm("max.f32 %0, %1, %2, %3;" : "=f"(mx) : "f"(mx), "f"(b), "f"(c));
                    uint32_t hits = __ballot_sync(0xffffffffu, mx >= T);   // lanes (windows) with a survivor among these 32 columns
                    // ~12 % of the units have one: the survivor lanes (few) are handled ONE AT A TIME, warp-wide: the lane
                    // parks its 32 values in shared memory, every lane then tests one column, and the ballot is the
                    // column mask -- each set lane appends its own candidate record (slots by popc prefix, no atomics)
                    while (hits) {
                        const int L = __ffs(hits) - 1;
                        hits &= hits - 1;
                        if (lane == L) {
#pragma unroll
                            for (int j = 0; j < 8; j++)
                                ((uint4 *)scratch)[j] = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                        }
                        __syncwarp();
                        const bool mine = scratch[lane] >= T;
                        const unsigned vote = __ballot_sync(0xffffffffu, mine);
                        if (mine) {
                            const uint32_t slot = my_count + __popc(vote & lane_lt);
                            if (slot < p.cand_cap) {
                                Hit h;
                                h.t = cbase + u * 32 + lane;
                                h.yx = ((uint32_t)y << 16) | (uint32_t)(gx0 + L);
                                h.sim = 0.f;
                                h.page = it.page;
                                my_list[slot] = h;
                            }
                        }
                        my_count += __popc(vote);
                        __syncwarp();
                        if (tron) tacc[2]++;
                    }
                    if (tron) tacc[5] += clock64() - ts_;
                };
                if (p.dbg_acc && (int)sb == dbg_sub) {  // parity probe: the raw numerator of one column (the fp16 MMA is off)
                    const uint32_t a = tc_ld1(tb + dbg_c);
                    tc_wait_ld();
                    const int gx = gx0 + lane;
                    if (gx >= 1 && gx <= p.dbg_xlast) p.dbg_acc[(size_t)y * p.r_w + gx] = a;
                }
                uint32_t v[TC_EPI_UNITS][32];
#pragma unroll
                for (int b = 0; b < TC_EPI_UNITS; b++)
                    if (b < nunits && !no_ld) tc_ld32(tb + b * 32, v[b]);
#pragma unroll
                for (int b = 0; b < TC_EPI_UNITS; b++)
                    if (b < nunits && !no_ld) tc_wait_ld32(v[b]);
                if (tlrole && lane == 0) TL(tlrole, job, 1);
                if (n2 > 0) {   // more than four units: screen the first ones, reload their registers with the last ones
#pragma unroll
                    for (int b = 0; b < TC_EPI_UNITS; b++)
                        if (b < n2) {
                            if (!no_screen) screen(v[b], b);
                            if (!no_ld) tc_ld32(tb + (b + TC_EPI_UNITS) * 32, v[b]);
                        }
#pragma unroll
                    for (int b = 0; b < TC_EPI_UNITS; b++)
                        if (b < n2 && !no_ld) tc_wait_ld32(v[b]);
                }
                // every tcgen05.ld of this job has completed: hand the accumulator back, then screen what is in registers
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(t_empty + 2 * acc + next_pipe);
                if (tlrole && lane == 0) TL(tlrole, job, 2);
                if (tron) tacc[3] += clock64() - tseen_;   // time from "accumulator full" seen to "accumulator released"
                if (!no_screen) {
#pragma unroll
                    for (int b = 0; b < TC_EPI_UNITS; b++)
                        if (b < nunits) screen(v[b], b < n2 ? b + TC_EPI_UNITS : b);
                }
                if (tron) tacc[4] += clock64() - tseen_;       // ... to the end of the job (screens, candidates)
                if (tlrole && lane == 0) TL(tlrole, job, 3);
                }
                if (++acc == nbuf) acc = 0;
              }
            }
        }
        if (lane == 0) p.cand_count[blockIdx.x * TC_LISTS_PER_CTA + e] = my_count;
        if (e == 0 && lane == 0) TT_END(8);
    }

    // ---- teardown
    tc_fence_before();
    __syncthreads();
    if (warp == 3) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------- exact pass
// Every survivor of the tensor-core screen gets the reference's own arithmetic: the numerator is
// recomputed with integer multiply-adds (ncc.cpp:108-166: same integers, any order), then the f64
// normalisation of ncc.cpp:212-220 via ncc_exact and ncc.rs:309-311 via patch_rnorm.  The decision
// `sim > threshold` and the f32 score are bit-identical to the CPU.  Runs right after the MMA launches
// of a box size, while its statistic planes are live.
struct CandArgs {
    const Hit *cands;        // [n_lists][cand_cap]
    uint32_t cand_cap, n_lists;
    uint32_t split;          // blocks per list (each takes every split-th group of blockDim candidates)
    const unsigned int *cand_count;  // [n_lists]
    unsigned int *cand_max;  // high-water mark of a list's count (overflow detection on the host)
    const TcColInfo *col_info;   // [n_blocks*nb] per group column: template constants, bank index (0xFFFFFFFF = padding), box size
    const uint8_t *rows;         // [n_blocks*nb][n_h][np] zero-padded template rows in column order
    const uint8_t *inv;
    size_t inv_page_stride;
    int pitch, n_h, np;
    int n_w[2];              // box widths of the group (one or two box sizes of the same height)
    double n_d[2], thr_d;
    HitSink sink;
};

__global__ void __launch_bounds__(256) cand_exact_kernel(CandArgs a)
{
    for (uint32_t bl = blockIdx.x; bl < a.n_lists * a.split; bl += gridDim.x) {
        const uint32_t l = bl / a.split, part = bl - l * a.split;
        const unsigned total = a.cand_count[l];
        if (threadIdx.x == 0 && part == 0 && total > a.cand_cap) atomicMax(a.cand_max, total);
        const unsigned n = min(total, a.cand_cap);
        const Hit *list = a.cands + (size_t)l * a.cand_cap;
        for (unsigned i = part * blockDim.x + threadIdx.x; i < n; i += blockDim.x * a.split) {
            const Hit c = list[i];
            // everything the candidate needs hangs off its record directly (column -> constants and template rows, window ->
            // page rows): ONE dependent round trip after the list read, then the rows stream
            const TcColInfo ti = a.col_info[c.t];
            const uint32_t t = ti.bank_t;
            if (t == 0xFFFFFFFFu) continue;
            const uint32_t y = c.yx >> 16, x = c.yx & 0xFFFFu;
            // exact numerator: u8 x u8 -> u32 with __dp4a on byte-shifted page words; the template rows are
            // zero padded to np bytes, so bytes beyond n_w contribute nothing
            const uint32_t *trow = (const uint32_t *)(a.rows + (size_t)c.t * a.n_h * a.np);
            const uint8_t *p0 = a.inv + (size_t)c.page * a.inv_page_stride + (size_t)y * a.pitch + x;
            const int sh = (int)((uintptr_t)p0 & 3) * 8, nw4 = a.np >> 2;
            const uint32_t *prow = (const uint32_t *)((uintptr_t)p0 & ~(uintptr_t)3);
            // ... and the window's sum and sum of squares (ncc.rs:307-308) from the same words, bytes beyond n_w masked off:
            // the exact pass reads no statistics plane
            const int bs = (int)ti.bs;
            const int n_w = a.n_w[bs];
            uint32_t acc = 0, s2_p = 0, s_p = 0;
            const int pitch4 = a.pitch >> 2;
            auto rows = [&](auto nw4_c) {   // nw4_c: words per template row as a compile-time constant (0: run-time nw4)
                constexpr int NW4 = decltype(nw4_c)::value;
                const int n4 = NW4 ? NW4 : nw4;
#pragma unroll 4
                for (int ny = 0; ny < a.n_h; ny++, trow += n4, prow += pitch4) {
                    uint32_t lo = __ldg(prow);
#pragma unroll
                    for (int q0 = 0; q0 < n4; q0 += 4) {
                        const uint4 tv = __ldg((const uint4 *)(trow + q0));   // template rows are 16-byte aligned (np = 16 or 32)
                        const uint32_t tw[4] = {tv.x, tv.y, tv.z, tv.w};
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const int q = q0 + k;
                            const uint32_t hi = __ldg(prow + q + 1);
                            const uint32_t w = __funnelshift_r(lo, hi, sh);
                            acc = __dp4a(w, tw[k], acc);
                            const int valid = n_w - 4 * q;   // bytes of this word inside the window
                            const uint32_t wm = valid >= 4 ? w : (valid <= 0 ? 0u : (w & ((1u << (8 * valid)) - 1u)));
                            s2_p = __dp4a(wm, wm, s2_p);
                            s_p = __dp4a(wm, 0x01010101u, s_p);
                            lo = hi;
                        }
                    }
                }
            };
            if (nw4 == 4) rows(std::integral_constant<int, 4>{});
            else if (nw4 == 8) rows(std::integral_constant<int, 8>{});
            else rows(std::integral_constant<int, 0>{});
            const double rn_p = patch_rnorm(s_p, s2_p, a.n_d[bs]);
            float sim;
            if (ncc_exact(acc, s_p, rn_p, ti.s_n, ti.n_recip, ti.rnorm_n, a.thr_d, &sim)) {
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
}

// ---------------------------------------------------------------------------------------------- host side
static size_t tc_smem_bytes(uint32_t btile_bytes, int ring, int n_mirror, int row_pitch, int nb, int a2_groups, int a2_slot)
{
    return ((btile_bytes + 127) & ~127u) + (size_t)(ring + n_mirror) * row_pitch + ((TC_RAW_SLOTS * TC_RAW_BYTES + 127) & ~127) +
           (size_t)a2_groups * TC_G * a2_slot + (size_t)2 * nb * 16 +
           (1 + 2 * TC_RAW_GROUPS + 2 * TC_RING_MAX + 2 * TC_A2_GROUPS + 4 * TC_MAX_BUF) * 8 + 64 + 128 + 16 + TC_EPI_WARPS * 128;
}

// pos[i] = place of the group's i-th template (first box size first) in a column order that keeps similar templates
// together, by the mean-free normalised correlations of the zero-padded template rows: the epilogue's column units
// (unit_caps: their sizes in column order) are filled with clusters, or -- without unit sizes -- a greedy nearest-neighbour chain
static std::vector<uint32_t> similarity_order(const TcClassSrc *src, uint32_t ncls, uint32_t n_tpl0, uint32_t n_tpl, uint32_t n_h,
                                              uint32_t np, const std::vector<uint32_t> &unit_caps = {})
{
    std::vector<uint32_t> pos(n_tpl);
    for (uint32_t i = 0; i < n_tpl; i++) pos[i] = i;
    if (n_tpl < 3 || n_tpl > 4096 || getenv("FOCR_TC_NOREORDER")) return pos;
    (void)ncls;
    // (large banks of large boxes: every `stride`-th pixel is enough to tell look-alikes apart and bounds the n^2 * K work)
    const size_t K0 = (size_t)n_h * np, stride = (n_tpl > 512 && K0 > 256) ? K0 / 256 : 1, K = (K0 + stride - 1) / stride;
    std::vector<float> vec((size_t)n_tpl * K);
    for (uint32_t i = 0; i < n_tpl; i++) {
        const uint32_t bs = i < n_tpl0 ? 0 : 1, li = bs ? i - n_tpl0 : i;
        const uint8_t *t = src[bs].rows_host + (size_t)li * K0;
        double mean = 0;
        for (size_t k = 0; k < K; k++) mean += t[k * stride];
        mean /= (double)K;
        double nrm = 0;
        for (size_t k = 0; k < K; k++) nrm += (t[k * stride] - mean) * (t[k * stride] - mean);
        const double inv = nrm > 0 ? 1.0 / std::sqrt(nrm) : 0.0;
        for (size_t k = 0; k < K; k++) vec[(size_t)i * K + k] = (float)((t[k * stride] - mean) * inv);
    }
    std::vector<float> sim((size_t)n_tpl * n_tpl);
    auto rows = [&](uint32_t i0, uint32_t step) {
        for (uint32_t i = i0; i < n_tpl; i += step)
            for (uint32_t j = 0; j <= i; j++) {
                const float *a = &vec[(size_t)i * K], *b = &vec[(size_t)j * K];
                float d = 0;
                for (size_t k = 0; k < K; k++) d += a[k] * b[k];
                sim[(size_t)i * n_tpl + j] = sim[(size_t)j * n_tpl + i] = d;
            }
    };
    const unsigned nt = n_tpl > 512 ? std::max(1u, std::min(8u, std::thread::hardware_concurrency())) : 1u;
    std::vector<std::thread> th;
    for (unsigned t = 1; t < nt; t++) th.emplace_back(rows, t, nt);
    rows(0, nt);
    for (auto &t : th) t.join();
    std::vector<char> used(n_tpl, 0);
    if (!unit_caps.empty() && !getenv("FOCR_TC_ORDER_CHAIN")) {
        // fill the epilogue's 32-column units one by one: seed = the unplaced template with the closest unplaced neighbour,
        // members = the unplaced templates most similar to the seed (measured 1 % faster than the plain chain below)
        uint32_t k = 0;
        for (uint32_t cap : unit_caps) {
            if (k >= n_tpl) break;
            int seed = -1;
            float best = -3.f;
            for (uint32_t i = 0; i < n_tpl; i++) {
                if (used[i]) continue;
                float nn = -2.f;
                for (uint32_t j = 0; j < n_tpl; j++)
                    if (!used[j] && j != i) nn = std::max(nn, sim[(size_t)i * n_tpl + j]);
                if (nn > best) best = nn, seed = (int)i;
            }
            used[seed] = 1;
            pos[seed] = k++;
            for (uint32_t m = 1; m < cap && k < n_tpl; m++) {
                int pick = -1;
                float bs = -3.f;
                for (uint32_t j = 0; j < n_tpl; j++)
                    if (!used[j] && sim[(size_t)seed * n_tpl + j] > bs) bs = sim[(size_t)seed * n_tpl + j], pick = (int)j;
                used[pick] = 1;
                pos[pick] = k++;
            }
        }
        return pos;
    }
    uint32_t cur = 0;
    used[0] = 1;
    pos[0] = 0;
    for (uint32_t k = 1; k < n_tpl; k++) {   // the unplaced template most similar to the last placed one
        uint32_t best = 0;
        float bs = -2.f;
        for (uint32_t j = 0; j < n_tpl; j++)
            if (!used[j] && sim[(size_t)cur * n_tpl + j] > bs) bs = sim[(size_t)cur * n_tpl + j], best = j;
        used[best] = 1;
        pos[best] = k;
        cur = best;
    }
    return pos;
}

int tc_class_build(TcClass &tc, const TcClassSrc *src, uint32_t ncls, uint32_t n_h, uint32_t np)
{
    tc.supported = false;
    if (ncls < 1 || ncls > 2) return 0;
    tc.ncls = ncls;
    tc.n_w = src[0].n_w;
    tc.n_w2 = ncls == 2 ? src[1].n_w : 0;
    tc.n_h = n_h;
    tc.np = np;
    tc.n_tpl0 = src[0].n_tpl;
    const uint32_t n_tpl = src[0].n_tpl + (ncls == 2 ? src[1].n_tpl : 0);
    tc.n_tpl = n_tpl;
    const uint32_t n_w_max = std::max(tc.n_w, tc.n_w2);
    // boxes at most 8 wide: two template rows per 16-byte K chunk (ncc_8_u8's class, ncc.cpp:48-251)
    tc.packed = np == 16 && n_w_max <= 8 && !getenv("FOCR_TC_NOPACK");
    tc.kchunks = tc.packed ? (n_h + 1) / 2 : n_h * (np / 16);
    tc.ksteps = (tc.kchunks + 1) / 2;
    const uint32_t n_hp = tc.packed ? 4 * tc.ksteps - 1 : (np == 16 ? (n_h + 1) & ~1u : n_h);
    tc.n_hp = n_hp;
    // tall boxes: less look-ahead and a shorter A2 ring leave room for the B tile; boxes wider than 16 (np == 32) take
    // one ring slot per K step, so the issue loop can wrap and the mirror slots are not needed
    const bool tall = n_hp > 16;
    tc.look_groups = tall ? 1 : TC_LOOK_GROUPS;
    tc.n_mirror = np == 16 ? n_hp - 1 : 0;
    tc.ring_groups = (n_hp + TC_G - 1 + TC_G - 1) / TC_G + tc.look_groups;
    const int ring = tc.ring_groups * TC_G;
    const int row_pitch = np == 16 ? 2048 : 2304;
    const int a2_slot = 2048 * (int)ncls;
    if (tc.ring_groups > (uint32_t)TC_RING_MAX) return 0;  // unsupported shape -> SIMT kernel
    // Boxes of more than 256 pixels: b*S + a*P outgrows the 23-bit linear range of the fp32 trick, so the SCREEN runs at
    // 1/2^sshift scale: the B tile holds ceil(t / 2^sshift) (acc' >= acc / 2^sshift: still errs towards more
    // candidates), b and a are divided by 2^sshift; the exact pass uses the true templates.
    tc.sshift = 0;
    while ((n_w_max * n_h) > (256u << tc.sshift)) tc.sshift++;
    if (tc.sshift > 3) return 0;
    // A2 ring: 4 groups of 4 rows when they fit next to a useful B tile, else 2 (tall boxes, two box sizes)
    tc.a2_groups = (tall || ncls == 2) ? 2 : TC_A2_GROUPS;
    // largest column count per launch (multiple of 32, <= 512 = two sub-blocks of 256) whose B tiles fit next to the rings
    int nb_max = 512;
    while (nb_max >= 32 &&
           tc_smem_bytes(2 * tc.ksteps * nb_max * 16, ring, (int)tc.n_mirror, row_pitch, nb_max, (int)tc.a2_groups, a2_slot) > TC_SMEM_BUDGET)
        nb_max -= 32;
    if (nb_max < 32) return 0;
    // Column blocks (launches) and sub-blocks (jobs per output row).  A sub-block is one accumulator: <= 256 columns, a
    // multiple of 32 (the epilogue works in 32-column units and every column of a unit must be written by the MMA; padding
    // columns carry -BIG in B2 and can never survive the screen).  More than 256 columns per launch are two sub-blocks that
    // share the row's operands; the accumulators then form a ring of 512 / columns-per-sub-block entries.
    // FOCR_TC_NBMAX / FOCR_TC_SPLIT: experiments (cap the columns per launch / always two sub-blocks).
    if (const char *e = getenv("FOCR_TC_NBMAX")) nb_max = std::max(32, std::min(nb_max, atoi(e) & ~31));
    tc.n_blocks = (n_tpl + nb_max - 1) / nb_max;
    const uint32_t per_blk = (n_tpl + tc.n_blocks - 1) / tc.n_blocks;   // templates per launch
    tc.nsub = (per_blk > 256 || (per_blk > 64 && getenv("FOCR_TC_SPLIT"))) ? 2 : 1;
    tc.nbsub = ((per_blk + tc.nsub - 1) / tc.nsub + 31) & ~31u;
    tc.nb = tc.nsub * tc.nbsub;
    // columns of a launch over its sub-blocks: the first one full (a whole number of 32-column units), the rest in the
    // second -- whose MMAs then only span its real columns rounded up to 16 (296 columns: 160 + 136 -> N = 160 and 144)
    auto cols0_of = [&](uint32_t cnt) { return std::min(cnt, tc.nbsub); };
    for (int c = 0; c < 2; c++) tc.blk_nmma[c].assign(tc.n_blocks, 16);
    for (uint32_t b = 0; b < tc.n_blocks; b++) {
        const uint32_t cnt = std::min(per_blk, n_tpl - b * per_blk), c0 = cols0_of(cnt), c1 = cnt - c0;
        tc.blk_nmma[0][b] = std::max(16u, (c0 + 15) & ~15u);
        tc.blk_nmma[1][b] = std::max(16u, (c1 + 15) & ~15u);
    }
    const size_t subtile = (size_t)2 * tc.ksteps * tc.nbsub * 16, tile = subtile * tc.nsub;
    std::vector<uint8_t> bt(tile * tc.n_blocks, 0);
    std::vector<float4> cst((size_t)tc.n_blocks * tc.nb);
    std::vector<uint8_t> rows_all((size_t)tc.n_blocks * tc.nb * n_h * np, 0);   // column order, zero for padding columns
    std::vector<TcColInfo> cinfo((size_t)tc.n_blocks * tc.nb, TcColInfo{0., 0., 0., 0xFFFFFFFFu, 0u});
    const float inf = INFINITY;
    for (auto &c : cst) c = make_float4(inf, 0.f, 0.f, 0.f);
    for (int c = 0; c < 2; c++) {
        tc.blk_bmax[c].assign(tc.n_blocks, 0.f);
        tc.blk_normmax[c].assign(tc.n_blocks, 0.f);
    }
    tc.col_of.assign(n_tpl, 0);
    // Column order.  The epilogue's slow path runs once per (window, 32-column unit) that holds a survivor, and a window that
    // matches a glyph matches its look-alikes too -- the same letter at the neighbouring subpixel shifts first of all, which
    // the bank order (offset index, letter) puts a whole alphabet apart.  Similar templates are therefore placed in
    // the same 32-column unit (clusters over the templates' normalised correlations), so that a matching window's survivors
    // fall into one unit instead of several (config 3: correlation kernel 0.414 -> 0.383 ms/page).  Results are indexed through the per-column records (TcColInfo): the order of
    // the columns is invisible outside the kernel.
    std::vector<uint32_t> unit_caps;   // sizes of the epilogue's column units in column order
    for (uint32_t b = 0; b < tc.n_blocks; b++)
        for (uint32_t sb = 0; sb < tc.nsub; sb++)
        {
            const uint32_t cnt = std::min(per_blk, n_tpl - b * per_blk), in_sub = sb == 0 ? cols0_of(cnt) : cnt - cols0_of(cnt);
            for (uint32_t c0 = 0; c0 < in_sub; c0 += 32) unit_caps.push_back(std::min(32u, in_sub - c0));
        }
    const std::vector<uint32_t> pos = similarity_order(src, ncls, tc.n_tpl0, n_tpl, n_h, np, unit_caps);
    for (uint32_t i = 0; i < n_tpl; i++) {
        const uint32_t bs = i < tc.n_tpl0 ? 0 : 1, li = bs ? i - tc.n_tpl0 : i;   // box size, index within it
        const uint8_t *trows = src[bs].rows_host + (size_t)li * n_h * np;
        const uint32_t ip = pos[i];                                               // place in the column order
        const uint32_t blk = ip / per_blk, r = ip % per_blk;
        const uint32_t cols0 = cols0_of(std::min(per_blk, n_tpl - blk * per_blk)), sub = r < cols0 ? 0 : 1, n = r - sub * cols0;
        const uint32_t col = sub * tc.nbsub + n;
        for (uint32_t kc = 0; kc < tc.kchunks; kc++) {
            uint8_t *dst = &bt[blk * tile + sub * subtile + ((size_t)kc * tc.nbsub + n) * 16];
            if (tc.packed) {   // rows 2kc and 2kc+1, 8 bytes each (a row beyond n_h is zero)
                for (int half = 0; half < 2; half++) {
                    const uint32_t row = 2 * kc + half;
                    for (int q = 0; q < 8; q++) dst[8 * half + q] = row < n_h ? trows[(size_t)row * np + q] : 0;
                }
                continue;
            }
            const uint32_t row = np == 16 ? kc : kc / 2, boff = np == 16 ? 0 : (kc & 1) * 16;
            const uint8_t *sp = trows + (size_t)row * np + boff;
            for (int q = 0; q < 16; q++) dst[q] = (uint8_t)((sp[q] + (1u << tc.sshift) - 1u) >> tc.sshift);
        }
        const TplInfo &ti = src[bs].info[li];
        // norm_n = sqrt(s2_n - s_n^2/n) = 1/rnorm_n ; constant (incl. all-zero) templates can never hit
        const double scale = 1.0 / (double)(1u << tc.sshift);
        const double norm_n = scale / ti.rnorm_n, b_t = scale * ti.s_n * ti.n_recip;
        const bool ok = std::isfinite(ti.rnorm_n) && ti.rnorm_n > 0 && std::isfinite(norm_n);
        cst[(size_t)blk * tc.nb + col] = make_float4(ok ? (float)norm_n : inf, (float)b_t, (float)bs, 0.f);
        memcpy(&rows_all[((size_t)blk * tc.nb + col) * n_h * np], trows, (size_t)n_h * np);
        cinfo[(size_t)blk * tc.nb + col] = TcColInfo{ti.rnorm_n, ti.n_recip, ti.s_n, src[bs].bank_index[li], bs};
        tc.col_of[i] = blk * tc.nb + col;
        if (ok) {
            tc.blk_bmax[bs][blk] = std::max(tc.blk_bmax[bs][blk], (float)b_t);
            tc.blk_normmax[bs][blk] = std::max(tc.blk_normmax[bs][blk], (float)norm_n);
        }
    }
    if (cudaMalloc(&tc.b_tiles, bt.size()) != cudaSuccess) return -1;
    if (cudaMalloc(&tc.consts, cst.size() * sizeof(float4)) != cudaSuccess) return -1;
    if (cudaMalloc(&tc.rows, rows_all.size()) != cudaSuccess) return -1;
    if (cudaMalloc(&tc.col_info, cinfo.size() * sizeof(TcColInfo)) != cudaSuccess) return -1;
    if (cudaMemcpy(tc.col_info, cinfo.data(), cinfo.size() * sizeof(TcColInfo), cudaMemcpyHostToDevice) != cudaSuccess) return -1;
    if (cudaMemcpy(tc.b_tiles, bt.data(), bt.size(), cudaMemcpyHostToDevice) != cudaSuccess) return -1;
    if (cudaMemcpy(tc.consts, cst.data(), cst.size() * sizeof(float4), cudaMemcpyHostToDevice) != cudaSuccess) return -1;
    if (cudaMemcpy(tc.rows, rows_all.data(), rows_all.size(), cudaMemcpyHostToDevice) != cudaSuccess) return -1;
    tc.supported = true;
    return 0;
}

void tc_class_release(TcClass &tc)
{
    if (tc.b_tiles) cudaFree(tc.b_tiles);
    if (tc.consts) cudaFree(tc.consts);
    if (tc.rows) cudaFree(tc.rows);
    if (tc.col_info) cudaFree(tc.col_info);
    tc.col_info = nullptr;
    tc.b_tiles = nullptr;
    tc.consts = nullptr;
    tc.rows = nullptr;
    tc.supported = false;
}

bool tc_class_supported(const TcClass &tc) { return tc.supported; }

cudaError_t launch_scan_tc(const TcClass &tc, const ScanArgs &a, int n_pages, int sm_count,
                           cudaStream_t st, int *n_launches, uint32_t *dbg_acc, int dbg_pos, TcHook *hook)
{
    if (!tc.supported) return cudaErrorNotSupported;
    if (dbg_acc && tc.sshift) return cudaErrorNotSupported;  // the screen of a > 256-pixel box runs on scaled templates: no raw numerators
    TcParams p;
    memset(&p, 0, sizeof(p));
    p.inv = a.inv;
    p.inv_page_stride = a.inv_page_stride;
    p.pitch = a.pitch;
    p.r_w = a.r_w;
    p.r_h = a.r_h;
    p.n_w = tc.n_w;
    p.n_w2 = tc.n_w2;
    p.ncls = tc.ncls;
    p.n_h = tc.n_h;
    p.np = tc.np;
    p.n_hp = tc.n_hp;
    p.packed = tc.packed ? 1 : 0;
    p.ksteps = tc.ksteps;
    p.nb = tc.nb;
    p.nsub = tc.nsub;
    p.nbs = tc.nbsub;
    p.nunits = tc.nbsub / 32;
    p.nmma[0] = p.nmma[1] = tc.nbsub;   // (set per launch below)
    p.nbuf = std::min(512 / p.nbs, TC_MAX_BUF);
    if (const char *e = getenv("FOCR_TC_NBUF")) p.nbuf = std::max(1, std::min(p.nbuf, atoi(e)));   // experiments
    p.ring_groups = tc.ring_groups;   // rows y..y+n_hp-1 may straddle one more group; + look-ahead
    p.n_mirror = tc.n_mirror;
    p.a2_groups = tc.a2_groups;
    p.a2_slot = 2048 * (int)tc.ncls;
    p.sshift = tc.sshift;
    p.ring = p.ring_groups * TC_G;
    p.row_pitch = tc.np == 16 ? 2048 : 2304;
    p.n_entries = tc.np == 16 ? 128 : 144;
    p.btile_bytes = 2 * tc.ksteps * tc.nb * 16;
    // the screen only has to pass a SUPERSET of the hits: |sim| <= 1 up to rounding, so thresholds beyond +-2 are
    // clamped (keeps thr*norm_n inside the fp16 range); cand_exact_kernel applies the real threshold
    p.thr = (float)std::min(std::max(a.thr_d, -2.0), 2.0);
    p.sp[0] = a.sp;
    p.pf[0] = a.pf;
    p.sp[1] = tc.ncls == 2 ? a.sp2 : a.sp;
    p.pf[1] = tc.ncls == 2 ? a.pf2 : a.pf;
    p.pack = a.pack;
    if (a.pack && tc.sshift) return cudaErrorInvalidValue;   // packed words need s_p < 2^16
    if (tc.ncls == 2 && (!a.sp2 || (!a.pf2 && !a.pack))) return cudaErrorInvalidValue;
    p.spitch = a.spitch;
    p.plane_page_stride = a.plane_page_stride;
    p.cands = a.cands;
    p.cand_cap = a.cand_cap;
    p.cand_count = a.cand_count;
    p.wd = a.cand_max ? a.cand_max + 1 : nullptr;   // api.cu: flags[4..9] follow the candidate high-water mark
    p.n_pages = n_pages;
    // windows of the NARROWEST box size: the strips cover x = 1 .. r_w - min n_w (the wider box flags its last columns invalid)
    const int n_w_min = tc.ncls == 2 ? (int)std::min(tc.n_w, tc.n_w2) : (int)tc.n_w;
    const int xs = a.r_w - n_w_min + 1, ys = a.r_h - (int)tc.n_h;  // output rows 1 .. r_h-n_h
    if (xs <= 0 || ys <= 0) return cudaSuccess;
    p.n_xstrips = (xs + 127) / 128;
    // Work items = (page, 128-window strip, y-segment), dealt round-robin to one persistent CTA per SM.  The segment height is
    // chosen per launch: the kernel takes (rounds of items per CTA) x (rows per item + a per-item cost: the n_hp-1 extra page
    // rows an item expands and the pipeline refill at its start, ~3 + n_hp/4 rows' worth), so the number of segments that
    // minimises that product wins -- long segments amortise the per-item cost, but the last round must not be half empty
    // (16 pages of 2480x3508: 18 segments of 195 rows, 39 full rounds, 1.3 % faster than 28 x 125; a single 608x800 page: 29
    // segments so that 145 items fill the 148 SMs once).
    {
        const long long cols = (long long)n_pages * p.n_xstrips;
        const long long item_cost = 3 + p.n_hp / 4;
        const long long segs_lo = std::max<long long>(1, (ys + 255) / 256), segs_hi = std::max<long long>(segs_lo, (ys + 7) / 8);
        long long best = -1, best_segs = segs_lo;
        for (long long segs = segs_lo; segs <= segs_hi; segs++) {
            const long long rows = (ys + segs - 1) / segs, rounds = (cols * ((ys + rows - 1) / rows) + sm_count - 1) / sm_count;
            const long long cost = rounds * (rows + item_cost);
            if (best < 0 || cost < best) best = cost, best_segs = segs;
        }
        p.yseg = (int)((ys + best_segs - 1) / best_segs);
    }
    if (const char *e = getenv("FOCR_TC_YSEG")) p.yseg = std::max(1, atoi(e));   // experiments
    p.n_ysegs = (ys + p.yseg - 1) / p.yseg;
    const size_t smem = tc_smem_bytes(p.btile_bytes, p.ring, p.n_mirror, p.row_pitch, p.nb, p.a2_groups, p.a2_slot);
    void (*kernel)(const TcParams) = nullptr;
    switch (p.nunits) {
        case 1: kernel = scan_tc_kernel<1>; break;
        case 2: kernel = scan_tc_kernel<2>; break;
        case 3: kernel = scan_tc_kernel<3>; break;
        case 4: kernel = scan_tc_kernel<4>; break;
        case 5: kernel = scan_tc_kernel<5>; break;
        case 6: kernel = scan_tc_kernel<6>; break;
        case 7: kernel = scan_tc_kernel<7>; break;
        case 8: kernel = scan_tc_kernel<8>; break;
        default: return cudaErrorInvalidValue;
    }
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    const int items = n_pages * p.n_xstrips * p.n_ysegs;
    const int grid = std::min(items, sm_count);
    p.dbg_acc = dbg_acc;
    const uint32_t dbg_colidx = dbg_acc ? tc.col_of[dbg_pos] : 0;   // dbg_pos = template index within the group
    p.dbg_col = dbg_acc ? (int)(dbg_colidx % tc.nb) : -1;
    p.dbg_xlast = a.r_w - (int)((dbg_acc && (uint32_t)dbg_pos >= tc.n_tpl0) ? tc.n_w2 : tc.n_w);
    {
        const char *dm = getenv("FOCR_TC_DBG");
        p.dbg_mode = dm ? atoi(dm) : 0;
    }
    uint4 *wdlog = nullptr;
    if (getenv("FOCR_TC_WDLOG") && cudaMalloc((void **)&wdlog, (size_t)grid * 32 * sizeof(uint4)) == cudaSuccess) {
        cudaMemsetAsync(wdlog, 0, (size_t)grid * 32 * sizeof(uint4), st);
        cudaMemcpyToSymbolAsync(g_wdlog, &wdlog, sizeof(wdlog), 0, cudaMemcpyHostToDevice, st);
    }
    const char *trace_path = dbg_acc ? nullptr : getenv("FOCR_TC_TRACE");
    for (uint32_t blk = 0; blk < tc.n_blocks; blk++) {
        if (dbg_acc && blk != dbg_colidx / tc.nb) continue;
        const size_t trace_words = 64 + (size_t)4 * TC_TL_JOBS * 4;
        if (trace_path && cudaMalloc((void **)&p.trace, trace_words * 8) == cudaSuccess) {
            cudaMemsetAsync(p.trace, 0, trace_words * 8, st);
            const char *tl = getenv("FOCR_TC_TIMELINE");
            p.timeline = tl != nullptr;
            p.tl0 = tl ? (uint32_t)atoi(tl) : 0;
        }
        p.nmma[0] = (int)tc.blk_nmma[0][blk];
        p.nmma[1] = (int)tc.blk_nmma[1][blk];
        if (getenv("FOCR_TC_FULLN")) p.nmma[0] = p.nmma[1] = p.nbs;   // experiments: MMAs over the whole accumulator stride
        p.btile = tc.b_tiles + (size_t)blk * p.btile_bytes;
        p.colconst = tc.consts + (size_t)blk * tc.nb;
        p.col_base = blk * tc.nb;
        for (int c = 0; c < 2; c++) {
            p.bmax[c] = tc.blk_bmax[c][blk] * 1.001f;
            p.amax[c] = std::max(p.thr * tc.blk_normmax[c][blk], 0.f) * 1.001f;
        }
        kernel<<<grid, TC_THREADS, smem, st>>>(p);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        if (n_launches) (*n_launches)++;
        if (wdlog) {
            std::vector<uint4> h((size_t)grid * 32);
            cudaStreamSynchronize(st);
            cudaMemcpy(h.data(), wdlog, h.size() * sizeof(uint4), cudaMemcpyDeviceToHost);
            unsigned int wdh[10] = {0};
            if (p.wd) cudaMemcpy(wdh, p.wd, sizeof(wdh), cudaMemcpyDeviceToHost);
            if (wdh[0]) {
                const unsigned cta = wdh[3];
                fprintf(stderr, "[focr] scan_tc stalled: box %ux%u nb %d nbuf %d ring %d; first report tag %u info %u CTA %u warp %u\n", tc.n_w,
                        tc.n_h, p.nb, p.nbuf, p.ring, wdh[1], wdh[2], cta, wdh[4]);
                uint32_t pg[32];
                cudaMemcpyFromSymbol(pg, g_wdprog, sizeof(pg));
                fprintf(stderr, "   progress snapshot of CTA %u (taken by warp %u, tag %u): issuer stage %u o %u | toeplitz g0", wdh[7], wdh[8],
                        wdh[9], pg[1] >> 24, pg[1] & 0xFFFFFF);
                for (int w = 4; w < 8; w++) fprintf(stderr, " %u", pg[w]);
                fprintf(stderr, " | a2 o0");
                for (int w = 8; w < 12; w++) fprintf(stderr, " %u", pg[w]);
                fprintf(stderr, " | epilogue rows");
                for (int w = 12; w < 28; w++) fprintf(stderr, " %u", pg[w]);
                fprintf(stderr, "\n");
                for (int w = 0; w < 28 && getenv("FOCR_TC_WDLOG")[0] == '2'; w++) {
                    const uint4 v = h[(size_t)cta * 32 + w];
                    fprintf(stderr, "   warp %2d: tag %2u info %6u parity %u bar@%u\n", w, v.x, v.y, v.z, v.w);
                }
            }
            cudaMemsetAsync(wdlog, 0, (size_t)grid * 32 * sizeof(uint4), st);
        }
        if (p.trace) {  // dump CTA 0's per-row timestamps: one file per (box size, N-block) launch, last launch wins
            std::vector<long long> h(trace_words);
            cudaStreamSynchronize(st);
            cudaMemcpy(h.data(), p.trace, h.size() * 8, cudaMemcpyDeviceToHost);
            cudaFree(p.trace);
            p.trace = nullptr;
            const std::string fn = std::string(trace_path) + "." + std::to_string(tc.n_w) + "x" + std::to_string(tc.n_h) + "." + std::to_string(blk);
            if (FILE *f = fopen(fn.c_str(), "wb")) {
                fwrite(h.data(), 8, h.size(), f);
                fclose(f);
            }
        }
        if (dbg_acc) continue;
        // the per-warp candidate lists are rewritten by every launch: run the exact pass right away
        CandArgs ca;
        ca.cands = a.cands;
        ca.cand_cap = a.cand_cap;
        ca.n_lists = (uint32_t)grid * TC_LISTS_PER_CTA;
        ca.cand_count = a.cand_count;
        ca.cand_max = a.cand_max;
        ca.col_info = (const TcColInfo *)tc.col_info;
        ca.rows = tc.rows;
        ca.inv = a.inv;
        ca.inv_page_stride = a.inv_page_stride;
        ca.pitch = a.pitch;
        ca.n_h = tc.n_h;
        ca.np = tc.np;
        ca.n_w[0] = tc.n_w;
        ca.n_w[1] = tc.ncls == 2 ? tc.n_w2 : tc.n_w;
        ca.n_d[0] = (double)(ca.n_w[0] * tc.n_h);
        ca.n_d[1] = (double)(ca.n_w[1] * tc.n_h);
        ca.thr_d = a.thr_d;
        ca.sink = a.sink;
        if (hook) hook->exact_begin();
        {
            static const int split = [] { const char *e = getenv("FOCR_EXACT_SPLIT"); return e ? std::max(1, std::min(16, atoi(e))) : 4; }();   // 4 blocks per list balance the uneven lists (measured 0.038 -> 0.0335 ms/page)
            ca.split = (uint32_t)split;
            cand_exact_kernel<<<grid * TC_LISTS_PER_CTA * split, 256, 0, st>>>(ca);
        }
        e = cudaGetLastError();
        if (hook) hook->exact_end();
        if (e != cudaSuccess) return e;
        if (n_launches) (*n_launches)++;
    }
    return cudaSuccess;
}

}  // namespace focr
