// scan_tc.cuh -- host interface of the tcgen05 correlation kernel (scan_tc.cu).
#pragma once
#include <vector>

#include "common.cuh"
#include "kernels.cuh"

namespace focr {

constexpr int TC_LISTS_PER_CTA = 8;   // epilogue warps per CTA, each with a private candidate list

// A LAUNCH GROUP of the tcgen05 kernel: the templates of one box size, or of TWO box sizes of the same height and
// padded row width (np).  The correlation GEMM does not see the box width (template rows are zero padded to np
// bytes); only the rank-2 normalisation does, and the fp16 MMA that applies it has room for two sets of window
// statistics (its two 16-byte K chunks).  The templates are re-laid out as the B operand of tcgen05.mma
// (K-major, no swizzle).
struct TcClass {
    bool supported = false;
    uint32_t ncls = 1;      // box sizes in this group (1 or 2)
    uint32_t n_w = 0, n_w2 = 0, n_h = 0, np = 0;   // n_w2: width of the second box size (ncls == 2)
    uint32_t n_tpl = 0;     // real templates of the group (first box size first)
    uint32_t n_tpl0 = 0;    // ... of which the first n_tpl0 have the first box size
    uint32_t nb = 0;        // columns per launch = nsub * nbsub
    uint32_t nsub = 0;      // sub-blocks per launch: every output row is nsub jobs (one accumulator each) sharing its operands
    uint32_t nbsub = 0;     // columns per sub-block = N of the MMAs (multiple of 32, <= 256)
    uint32_t n_blocks = 0;  // launches per chunk of pages
    bool packed = false;    // boxes at most 8 wide: a K chunk holds TWO template rows of 8 bytes (half the MMA work of one row per chunk)
    uint32_t n_hp = 0;      // ring slots (page rows) an output row spans
    uint32_t sshift = 0;    // boxes with more than 256 pixels: the screen runs on templates scaled by 2^-sshift (rounded up)
    uint32_t n_mirror = 0;  // ring slots stored twice (np == 16: an output row never wraps); 0 = the issue loop wraps
    uint32_t look_groups = 0, a2_groups = 0, ring_groups = 0;  // ring sizing of this group
    uint32_t kchunks = 0;   // 16-byte K chunks per template = n_h * np/16
    uint32_t ksteps = 0;    // tcgen05.mma instructions per output tile = ceil(kchunks/2) (K = 32 each)
    uint8_t *b_tiles = nullptr;   // device [n_blocks][nsub][2*ksteps][nbsub][16]
    float4 *consts = nullptr;     // device [n_blocks*nb] {norm_n, s_n/n, box size index as float, -}; norm_n = +inf for padding / constant templates
    uint8_t *rows = nullptr;      // device [n_blocks*nb][n_h][np] zero-padded template rows in COLUMN order (exact pass; zero for padding columns)
    void *col_info = nullptr;     // device [n_blocks*nb] TcColInfo: what the exact pass needs per column, one 32-byte record
    std::vector<uint32_t> col_of; // host, per group-local template: its column (launch * nb + column)
    std::vector<uint32_t> blk_nmma[2];   // host, per N-block and sub-block: N of the MMAs (real columns rounded up to 16)
    std::vector<float> blk_bmax[2], blk_normmax[2];  // host, per N-block and box size: max s_n/n and max norm_n over its real columns
};

// per column of a launch group: the template's constants for the exact pass (ncc.cpp:73-86), its bank index and box size
struct __align__(16) TcColInfo {
    double rnorm_n, n_recip, s_n;
    uint32_t bank_t;   // 0xFFFFFFFF: padding column
    uint32_t bs;       // which box size of the group (0 or 1)
};

// one box size going into a launch group
struct TcClassSrc {
    const uint8_t *rows_host;    // [n_tpl][n_h][np] zero-padded rows
    uint32_t n_w, n_tpl;
    const uint32_t *bank_index;  // [n_tpl]
    const TplInfo *info;         // [n_tpl]
};

// optional instrumentation around the exact pass (api.cu times it as its own stage)
struct TcHook {
    virtual void exact_begin() {}
    virtual void exact_end() {}
    virtual ~TcHook() {}
};

int tc_class_build(TcClass &tc, const TcClassSrc *src, uint32_t ncls, uint32_t n_h, uint32_t np);
void tc_class_release(TcClass &tc);
bool tc_class_supported(const TcClass &tc);
// a.sp / a.pf (and a.sp2 / a.pf2 for the second box size) are the window statistics planes of the group's box sizes.
// dbg_acc/dbg_pos: parity probe -- store the raw numerators of the group's dbg_pos-th template
cudaError_t launch_scan_tc(const TcClass &tc, const ScanArgs &a, int n_pages, int sm_count,
                           cudaStream_t st, int *n_launches, uint32_t *dbg_acc = nullptr, int dbg_pos = -1,
                           TcHook *hook = nullptr);

}  // namespace focr
