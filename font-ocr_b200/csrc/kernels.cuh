// kernels.cuh -- argument blocks and host launchers of every kernel in libfocr_b200.so.
#pragma once
#include "common.cuh"

namespace focr {

struct StatsArgs {
    const uint8_t *inv;      // inverted pages
    size_t inv_page_stride;  // bytes between pages
    int pitch;               // bytes per page row
    int r_w, r_h, n_w, n_h;
    float inv_n_f;           // 1/(n_w*n_h) in f32
    // output planes, [page][y][x] with row pitch `spitch` and page stride `plane_page_stride` (elements)
    uint32_t *sp, *s2p;
    float *pf;
    double *rn;  // may be NULL (only the SIMT scan and the parity probe need the f64 plane)
    int spitch;
    size_t plane_page_stride;
    // optional second box of the SAME height (n_w2 != 0): its sp / pf planes come out of the same pass over the page
    // (the vertical sums and the row prefix sums do not depend on the width)
    int n_w2;
    float inv_n_f2;
    uint32_t *sp2;
    float *pf2;
    // pack != 0 (boxes of at most 256 pixels, tcgen05 path): ONE word per window, `s_p | fix(norm_p) << 16` with norm_p in
    // 11.5 fixed point (0xFFFF = constant window), written to sp / sp2; the pf planes are not touched.  Half the bytes.
    int pack;
};

// what every scan kernel appends to
struct HitSink {
    Hit *hits;
    uint32_t hit_cap;               // per launch, all pages together
    unsigned int *hit_count;        // total appended (may exceed hit_cap -> host grows and retries)
    unsigned int *rowcount;         // [page][T][r_h]
    uint32_t T, r_h;
};

struct ScanArgs {
    const uint8_t *inv;
    size_t inv_page_stride;
    int pitch;
    int r_w, r_h;
    SizeClassDev cls;
    const TplInfo *tpl;      // [T] bank-wide
    const uint32_t *sp;
    const uint32_t *s2p;
    const float *pf;
    const double *rn;
    const uint32_t *sp2 = nullptr;   // tcgen05 launch groups of two box sizes: the planes of the second one
    const float *pf2 = nullptr;
    int pack = 0;                    // the sp planes hold packed words (StatsArgs::pack)
    int spitch;
    size_t plane_page_stride;
    double thr_d;
    float thr_f;
    HitSink sink;
    // tcgen05 path only: prefilter survivors wait here for the exact f64 pass
    Hit *cands;
    uint32_t cand_cap;
    unsigned int *cand_count, *cand_max;
    uint32_t *acc_out;       // parity probe: raw numerators of ONE template (cls.n_tpl == 1), [y*r_w+x]
};

struct FinalizeArgs {
    const Hit *hits;
    const unsigned int *hit_count;
    uint32_t hit_cap;
    const unsigned int *rowcount;   // [PT][r_h]
    uint32_t *y_cut;                // [PT]
    unsigned int *sel_count;        // [2][PT]: hits kept above the cut row / in the cut row
    unsigned long long *sel;        // [PT][sel_cap]: [0, n_out) above the cut row, [n_out, sel_cap) the cut row
    uint32_t sel_cap;
    unsigned int *overflow;         // set when a selection list overflowed (cannot happen by construction)
    uint32_t T, r_h, n_pages, n_out;
    focr_match *out;                // [PT][n_out]
    uint32_t *counts;               // [PT]
};

cudaError_t launch_stage_invert(const uint8_t *src, size_t src_page_stride, size_t src_pitch, uint8_t *dst,
                                size_t dst_page_stride, int dst_pitch, int r_w, int r_h, int n_pages, int invert,
                                cudaStream_t st);
cudaError_t launch_window_stats(const StatsArgs &a, int n_pages, cudaStream_t st);
cudaError_t launch_scan_simt(const ScanArgs &a, int n_pages, cudaStream_t st, int *n_launches);
cudaError_t launch_finalize(const FinalizeArgs &a, cudaStream_t st, int *n_launches);
size_t finalize_sel_cap(uint32_t r_w, uint32_t n_out);  // 0 = unsupported

}  // namespace focr
