// common.cuh -- shared device/host declarations for libfocr_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/focr_b200.h"

namespace focr {

// ---------------------------------------------------------------- data layout in HBM
//
// page      : inverted u8, row pitch `pitch` (multiple of 128 B so TMA/128-bit loads are aligned),
//             PAGE_PAD_ROWS zero rows after the last row so that any tile may over-read.
// stats     : one set of planes per (page, box size), pitch `spitch` elements:
//               sp  u32  window sum            (ncc.rs:307  s_p)
//               s2p u32  window sum of squares (ncc.rs:308  s2_p; < 2^32 for boxes <= 64x64)
//               rn  f64  1/sqrt(s2p - sp^2/n)  (ncc.rs:309-311 patch_rnorm, bit-identical)
//               sf,pf f32 prefilter operands of the tcgen05 epilogue (scan_tc.cu)
// hit list  : unordered {t, y<<16|x, sim, page} records appended by the scan kernels
// rowcount  : u32 [page][t][r_h]  hits per row, used to find the row where the n_out cap is reached
// out       : focr_match [page][t][n_out] + counts [page][t]

constexpr int PAGE_PAD_ROWS = 4;   // zero rows after a page: the scan's row pipeline reads up to 3 rows past the last one
constexpr int MAX_TPL_W = 32;
constexpr int MAX_TPL_H = 64;

struct Hit {  // 16 B
    uint32_t t;     // template index within the bank
    uint32_t yx;    // y << 16 | x
    float sim;
    uint32_t page;
};

// per-template constants (ncc.cpp:73-86), computed on the host in IEEE f64
struct TplInfo {
    double rnorm_n;   // 1/sqrt(s2_n - s_n^2/n)
    double n_recip;   // 1/n, n = n_w*n_h UNPADDED (ncc.cpp:70)
    double s_n;       // (double)s_n
    uint32_t n_w, n_h;
    uint32_t cls;     // size class
    uint32_t data_off;  // byte offset of the padded rows inside the class's template block
};

// one box size (n_w, n_h) and the templates that have it
struct SizeClassDev {
    uint32_t n_w, n_h;
    uint32_t np;          // padded row width in bytes: 16 or 32
    uint32_t n_tpl;       // templates in this class
    const uint32_t *tpl_index;  // [n_tpl] bank index of each
    const uint8_t *rows;  // [n_tpl][n_h][np] zero-padded rows (copy_needle_n_u8, ncc.rs:925-935)
};

// the exact f64 similarity of the reference, operation for operation (ncc.cpp:212-220, 237-240):
//   num = fnmadd(s_n*s_p, n_recip, acc);  den = rnorm_n*rnorm_p;  sim = num*den
// All double-precision CUDA ops used here are IEEE round-to-nearest; the intrinsics forbid
// contraction so that ONLY the reference's own fused multiply-add is fused.
__device__ __forceinline__ bool ncc_exact(uint32_t acc, uint32_t s_p, double rnorm_p, double s_n,
                                          double n_recip, double rnorm_n, double thr_d, float *sim_out)
{
    double prod = __dmul_rn(s_n, (double)s_p);              // exact (< 2^53)
    double num = __fma_rn(-prod, n_recip, (double)(int32_t)acc);
    double den = __dmul_rn(rnorm_n, rnorm_p);
    double sim = __dmul_rn(num, den);
    *sim_out = (float)sim;                                   // cvt.rn.f32.f64
    return (sim != __longlong_as_double(0x7ff0000000000000LL)) && (sim > thr_d);
}

// ncc.rs:309-311: patch_rnorm from exact window sums
__device__ __forceinline__ double patch_rnorm(uint32_t s_p, uint32_t s2_p, double n_d)
{
    double sq = (double)((unsigned long long)s_p * (unsigned long long)s_p);
    double norm = __dsub_rn((double)s2_p, __ddiv_rn(sq, n_d));
    return __ddiv_rn(1.0, __dsqrt_rn(norm));
}

}  // namespace focr
