// scan_tc.cuh -- host interface of the tcgen05 correlation kernel (scan_tc.cu).
#pragma once
#include <vector>

#include "common.cuh"
#include "kernels.cuh"

namespace focr {

constexpr int TC_LISTS_PER_CTA = 16;  // epilogue warps per CTA, each with a private candidate list

// per box size: the template bank re-laid out as the B operand of tcgen05.mma (K-major, no swizzle)
struct TcClass {
    bool supported = false;
    uint32_t n_w = 0, n_h = 0, np = 0;
    uint32_t n_tpl = 0;     // real templates of this class
    uint32_t nb = 0;        // columns per launch = nsub * nbsub
    uint32_t nsub = 0;      // sub-blocks per launch: every output row is nsub jobs (one accumulator each) sharing its operands
    uint32_t nbsub = 0;     // columns per sub-block = N of the MMAs (multiple of 32, <= 128 when nsub > 1)
    uint32_t n_blocks = 0;  // launches per chunk of pages
    uint32_t sshift = 0;    // boxes with more than 256 pixels: the screen runs on templates scaled by 2^-sshift (rounded up)
    uint32_t n_mirror = 0;  // ring slots stored twice (np == 16: an output row never wraps); 0 = the issue loop wraps
    uint32_t look_groups = 0, a2_groups = 0, ring_groups = 0;  // ring sizing of this class
    uint32_t kchunks = 0;   // 16-byte K chunks per template = n_h * np/16
    uint32_t ksteps = 0;    // tcgen05.mma instructions per output tile = ceil(kchunks/2) (K = 32 each)
    uint8_t *b_tiles = nullptr;   // device [n_blocks][2*ksteps][nb][16]
    float2 *consts = nullptr;     // device [n_blocks*nb] {norm_n, s_n/n}; norm_n = +inf for padding / constant templates
    uint32_t *tpl_of = nullptr;   // device [n_blocks*nb] bank index (0xFFFFFFFF for padding)
    uint32_t *cls_of = nullptr;   // device [n_blocks*nb] index within the class's template rows (0xFFFFFFFF for padding)
    std::vector<uint32_t> col_of; // host, per class-local template: its column (launch * nb + column)
    std::vector<float> blk_bmax, blk_normmax;  // host, per N-block: max s_n/n and max norm_n over its real columns
};

// optional instrumentation around the exact pass (api.cu times it as its own stage)
struct TcHook {
    virtual void exact_begin() {}
    virtual void exact_end() {}
    virtual ~TcHook() {}
};

int tc_class_build(TcClass &tc, const uint8_t *rows_host, uint32_t n_w, uint32_t n_h, uint32_t np, uint32_t n_tpl,
                   const uint32_t *bank_index, const TplInfo *info);
void tc_class_release(TcClass &tc);
bool tc_class_supported(const TcClass &tc);
// dbg_acc/dbg_pos: parity probe -- store the raw numerators of the class's dbg_pos-th template
cudaError_t launch_scan_tc(const TcClass &tc, const ScanArgs &a, int n_pages, int sm_count,
                           cudaStream_t st, int *n_launches, uint32_t *dbg_acc = nullptr, int dbg_pos = -1,
                           TcHook *hook = nullptr);

}  // namespace focr
