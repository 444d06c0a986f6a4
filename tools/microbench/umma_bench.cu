// umma_bench.cu -- tcgen05 micro-benchmarks (measurement aids; built into tools/microbench/libfocr_microbench.so,
// NOT part of libfocr_b200.so or its header).
//
// MEASURED_PEAKS.json has no integer tensor peak, and the correlation kernel's design hinges on how
// long ONE tcgen05.mma of a given N takes (scan_tc.cu issues many small ones).  Every SM runs one CTA
// that issues `iters` x `ksteps` MMAs (M = 128, K = 32, u8 x u8 -> s32, K-major no-swizzle operands in
// shared memory exactly like scan_tc.cu) into `nacc` rotating accumulators and times them with clock64.
#include <algorithm>
#include <cstdint>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "focr_microbench.h"

// measurement aids live outside the product library (libfocr_b200.so) and its header; they only share the error codes
enum { FOCR_OK = 0, FOCR_ERR_CUDA = 1, FOCR_ERR_ARG = 2 };
static thread_local std::string g_mb_err;
static int focr_internal_fail(int code, const std::string &msg)
{
    g_mb_err = msg;
    return code;
}
extern "C" const char *focr_microbench_last_error(void) { return g_mb_err.c_str(); }

namespace focr {

__device__ __forceinline__ uint32_t ub_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// PAIR: launched as clusters of two CTAs; the leader issues cta_group::2 MMAs (M = 256, each CTA holds N/2 columns of B)
template <bool PAIR>
__global__ void __launch_bounds__(128, 1) umma_i8_bench_kernel(int n, int ksteps, int iters, int nacc, long long *cycles)
{
    uint32_t rank = 0;
    if (PAIR) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    const int warp = threadIdx.x >> 5;
    // operands: A = ksteps*2 chunks of 128 x 16 B, B = ksteps*2 chunks of n x 16 B (contents irrelevant)
    for (int i = threadIdx.x; i < (ksteps * 2 * (128 + n) * 16) / 4; i += blockDim.x) ((uint32_t *)smem)[i] = 0x01010101u;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(ub_smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ub_smem_u32(&tmem_ptr)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ub_smem_u32(&tmem_ptr)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (PAIR) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    else __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_ptr;
    if (warp == 1) {
        const uint32_t idesc = (2u << 4) | ((uint32_t)(n >> 3) << 17) | (((PAIR ? 256u : 128u) >> 4) << 24);
        const uint32_t desc_hi = (128u >> 4) | (1u << 14);
        const uint32_t a0 = ((ub_smem_u32(smem) & 0x3FFFFu) >> 4) | ((2048u >> 4) << 16);
        const uint32_t bbase = ub_smem_u32(smem) + ksteps * 2 * 2048;
        const uint32_t b0 = ((bbase & 0x3FFFFu) >> 4) | (((uint32_t)n * 16u >> 4) << 16);
        const int nbs = (n + 31) & ~31;
        long long t0 = 0, t1 = 0;
        uint32_t elected;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(elected));
        __syncwarp();
        t0 = clock64();
        if (elected && rank == 0) {
            for (int it = 0; it < iters; it++) {
                const uint32_t d = tmem_base + (uint32_t)((it % nacc) * nbs);
                for (int k = 0; k < ksteps; k++) {
                    const uint64_t ad = ((uint64_t)desc_hi << 32) | (a0 + (uint32_t)k * 2 * (2048 >> 4));
                    const uint64_t bd = ((uint64_t)desc_hi << 32) | (b0 + (uint32_t)k * 2 * ((uint32_t)n * 16u >> 4));
                    if (PAIR)
                        asm volatile(
                            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                            "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                            "l"(ad), "l"(bd), "r"(idesc), "r"(k)
                            : "memory");
                    else
                        asm volatile(
                            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                            "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                            "l"(ad), "l"(bd), "r"(idesc), "r"(k)
                            : "memory");
                }
            }
            const long long ti = clock64();
            if (PAIR)
                asm volatile("{\n\t.reg .b16 m;\n\tmov.b16 m, 3;\n\t"
                             "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}" ::"r"(ub_smem_u32(&bar)) : "memory");
            else
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(ub_smem_u32(&bar)) : "memory");
            cycles[gridDim.x + blockIdx.x] = ti - t0;  // when the issuing thread got past the last tcgen05.mma
        }
        __syncwarp();
        uint32_t done;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(ub_smem_u32(&bar)), "r"(0) : "memory");
        } while (!done);
        t1 = clock64();
        if ((threadIdx.x & 31) == 0) cycles[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (PAIR) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    else __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// TMEM <-> register traffic micro-benchmark: `nw` warps (multiple of 4) each run `iters` rounds of
// tcgen05.ld 32x32b.x32 (+ wait) and/or tcgen05.st 32x32b.x32 over rotating 32-column units of their lane
// quarter, optionally while one more warp keeps the tensor core busy with `mma_count` x 7 MMAs of N = mma_n.
// out[2*blk] = slowest ld/st warp's cycles, out[2*blk+1] = the MMA warp's cycles.
__global__ void __launch_bounds__(32 * 17, 1) tmem_bench_kernel(int nw, int mode, int iters, int mma_n, int mma_count, long long *out)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    __shared__ long long wc[16];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ksteps = 7, n = mma_n > 0 ? mma_n : 16;
    for (int i = threadIdx.x; i < (ksteps * 2 * (128 + n) * 16) / 4; i += blockDim.x) ((uint32_t *)smem)[i] = 0x01010101u;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(ub_smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ub_smem_u32(&tmem_ptr)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_ptr;
    if (warp < nw) {
        const int q = warp & 3, grp = warp >> 2, ngrp = nw >> 2;
        const uint32_t tl = tmem_base + ((uint32_t)(q * 32) << 16);
        uint32_t sink = 0;
        int u = grp;
        const long long t0 = clock64();
        for (int it = 0; it < iters; it++) {
            const uint32_t ta = tl + (uint32_t)u * 32;
            if (mode != 2) {
                uint32_t v[32];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                    "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                      "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                      "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                      "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(ta));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (mode == 3) {
                    float mx = __uint_as_float(v[0]);
#pragma unroll
                    for (int j = 1; j < 31; j += 2)
                        asm("max.f32 %0, %1, %2, %3;" : "=f"(mx) : "f"(mx), "f"(__uint_as_float(v[j])), "f"(__uint_as_float(v[j + 1])));
                    mx = fmaxf(mx, __uint_as_float(v[31]));
                    if (__any_sync(0xffffffffu, mx >= 3.0e38f)) sink++;
                } else {
                    sink += v[3] ^ v[29];
                }
            }
            if (mode == 1 || mode == 2) {
                asm volatile(
                    "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
                    "{%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(ta),
                    "r"(0x4B000000u)
                    : "memory");
            }
            u += ngrp;
            if (u >= 16) u -= 16;
        }
        if (mode == 1 || mode == 2) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        const long long t1 = clock64();
        if (lane == 0) wc[warp] = (t1 - t0) + (sink == 0x12345678u ? 1 : 0);
    } else if (warp == nw && mma_n > 0) {
        const uint32_t idesc = (2u << 4) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t desc_hi = (128u >> 4) | (1u << 14);
        const uint32_t a0 = ((ub_smem_u32(smem) & 0x3FFFFu) >> 4) | ((2048u >> 4) << 16);
        const uint32_t bbase = ub_smem_u32(smem) + ksteps * 2 * 2048;
        const uint32_t b0 = ((bbase & 0x3FFFFu) >> 4) | (((uint32_t)n * 16u >> 4) << 16);
        const int nbs = (n + 31) & ~31, nacc = 512 / nbs;
        uint32_t elected;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(elected));
        __syncwarp();
        const long long t0 = clock64();
        if (elected) {
            for (int it = 0; it < mma_count; it++) {
                const uint32_t d = tmem_base + (uint32_t)((it % nacc) * nbs);
                for (int k = 0; k < ksteps; k++) {
                    const uint64_t ad = ((uint64_t)desc_hi << 32) | (a0 + (uint32_t)k * 2 * (2048 >> 4));
                    const uint64_t bd = ((uint64_t)desc_hi << 32) | (b0 + (uint32_t)k * 2 * ((uint32_t)n * 16u >> 4));
                    asm volatile(
                        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                        "l"(ad), "l"(bd), "r"(idesc), "r"(1)
                        : "memory");
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(ub_smem_u32(&bar)) : "memory");
        }
        __syncwarp();
        uint32_t done;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(ub_smem_u32(&bar)), "r"(0) : "memory");
        } while (!done);
        const long long t1 = clock64();
        if (lane == 0) out[2 * blockIdx.x + 1] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        long long m = 0;
        for (int i = 0; i < nw; i++) m = wc[i] > m ? wc[i] : m;
        out[2 * blockIdx.x] = m;
        if (mma_n <= 0) out[2 * blockIdx.x + 1] = 0;
    }
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// Hand-shake round trip: one thread signals barrier A (tcgen05.commit or a plain arrive), a whole warp waits for A and
// its lane 0 arrives on barrier B, the first thread waits for B -- `iters` times.  wait_kind 0 = mbarrier.try_wait with a
// 20 us suspend hint (what scan_tc.cu uses), 1 = try_wait without hint, 2 = test_wait polling.  nwait = waiting warps.
__device__ __forceinline__ void pp_wait(uint32_t addr, uint32_t parity, int kind)
{
    uint32_t done;
    do {
        if (kind == 0)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(addr), "r"(parity), "r"(20000u) : "memory");
        else if (kind == 1)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        else
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}
__global__ void __launch_bounds__(32 * 9, 1) pingpong_kernel(int wait_kind, int use_commit, int nwait, int iters, long long *out)
{
    __shared__ uint64_t bars[2];
    __shared__ uint32_t tmem_ptr;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(ub_smem_u32(&bars[0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ub_smem_u32(&bars[1])), "r"(nwait) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ub_smem_u32(&tmem_ptr)), "r"(32) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    const uint32_t a = ub_smem_u32(&bars[0]), b = ub_smem_u32(&bars[1]);
    if (warp == 0) {
        if (lane == 0) {
            const long long t0 = clock64();
            for (int it = 0; it < iters; it++) {
                if (use_commit)
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(a) : "memory");
                else
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory");
                pp_wait(b, it & 1, wait_kind);
            }
            out[blockIdx.x] = clock64() - t0;
        }
    } else if (warp <= nwait) {
        for (int it = 0; it < iters; it++) {
            pp_wait(a, it & 1, wait_kind);
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b) : "memory");
        }
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_ptr), "r"(32) : "memory");
}

}  // namespace focr


static double g_last_issue_cycles = 0;
// cycles (median over SMs) the issuing thread of the last focr_bench_umma_i8 run needed to get past its last
// tcgen05.mma: with the total this shows how far ahead of the tensor pipe an issuing thread may run
extern "C" double focr_bench_umma_issue_cycles(void) { return g_last_issue_cycles; }

// n: MMA N (multiple of 16, <= 256); returns the median over SMs of cycles per tcgen05.mma and the wall time
static int umma_i8_impl(int device, int n, int ksteps, int iters, int nacc, double *cycles_per_mma, double *ms_total, bool pair);
extern "C" int focr_bench_umma_i8(int device, int n, int ksteps, int iters, int nacc, double *cycles_per_mma,
                                  double *ms_total)
{
    return umma_i8_impl(device, n, ksteps, iters, nacc, cycles_per_mma, ms_total, false);
}
// the same stream issued as cta_group::2 MMAs (M = 256) by the leader of every CTA pair; cycles per MMA on the leader
extern "C" int focr_bench_umma_i8_2cta(int device, int n, int ksteps, int iters, int nacc, double *cycles_per_mma,
                                       double *ms_total)
{
    return umma_i8_impl(device, n, ksteps, iters, nacc, cycles_per_mma, ms_total, true);
}
static int umma_i8_impl(int device, int n, int ksteps, int iters, int nacc, double *cycles_per_mma, double *ms_total, bool pair)
{
    using namespace focr;
    if (device < 0 || !cycles_per_mma || !ms_total || n < 16 || n > 256 || (n & 15) || ksteps < 1 || ksteps > 16 || iters < 1 ||
        nacc < 1 || nacc * ((n + 31) & ~31) > 512)
        return focr_internal_fail(FOCR_ERR_ARG, "focr_bench_umma_i8: bad argument");
    if (cudaSetDevice(device) != cudaSuccess) return focr_internal_fail(FOCR_ERR_CUDA, "cudaSetDevice");
    cudaStream_t st = nullptr;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    long long *d = nullptr;
    if (cudaMalloc((void **)&d, sms * 16) != cudaSuccess) return focr_internal_fail(FOCR_ERR_CUDA, "cudaMalloc");
    const size_t smem = (size_t)ksteps * 2 * (128 + n) * 16 + 1024;
    void (*kernel)(int, int, int, int, long long *) = pair ? umma_i8_bench_kernel<true> : umma_i8_bench_kernel<false>;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return focr_internal_fail(FOCR_ERR_CUDA, "cudaFuncSetAttribute");
    if (pair) sms &= ~1;
    auto launch = [&](int its) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(sms);
        cfg.blockDim = dim3(128);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = pair ? 2 : 1;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, kernel, n, ksteps, its, nacc, d);
    };
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    launch(8);  // warm-up
    cudaEventRecord(e0, st);
    launch(iters);
    cudaEventRecord(e1, st);
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return focr_internal_fail(FOCR_ERR_CUDA, std::string("umma bench: ") + cudaGetErrorString(e));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> h(sms * 2);
    cudaMemcpy(h.data(), d, sms * 16, cudaMemcpyDeviceToHost);
    cudaFree(d);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    std::sort(h.begin(), h.begin() + sms);
    std::sort(h.begin() + sms, h.end());
    *cycles_per_mma = (double)h[sms / 2] / ((double)iters * ksteps);
    if (iters < 0) *cycles_per_mma = 0;
    g_last_issue_cycles = (double)h[sms + sms / 2];
    *ms_total = ms;
    return FOCR_OK;
}

// TMEM load/store micro-benchmark (see tmem_bench_kernel): cycles per 32x32b.x32 round per warp (median over SMs of the
// slowest warp) and cycles per MMA of the concurrent tensor-core stream (0 when mma_n == 0).
extern "C" int focr_bench_tmem(int device, int nw, int mode, int iters, int mma_n, int mma_count, double *cycles_per_round,
                               double *cycles_per_mma)
{
    using namespace focr;
    if (device < 0 || !cycles_per_round || !cycles_per_mma || nw < 4 || nw > 16 || (nw & 3) || mode < 0 || mode > 3 || iters < 1 ||
        mma_n < 0 || mma_n > 256 || (mma_n & 15))
        return focr_internal_fail(FOCR_ERR_ARG, "focr_bench_tmem: bad argument");
    if (cudaSetDevice(device) != cudaSuccess) return focr_internal_fail(FOCR_ERR_CUDA, "cudaSetDevice");
    cudaStream_t st = nullptr;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    long long *d = nullptr;
    if (cudaMalloc((void **)&d, sms * 16) != cudaSuccess) return focr_internal_fail(FOCR_ERR_CUDA, "cudaMalloc");
    const size_t smem = (size_t)7 * 2 * (128 + 256) * 16 + 1024;
    if (cudaFuncSetAttribute(tmem_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return focr_internal_fail(FOCR_ERR_CUDA, "cudaFuncSetAttribute");
    tmem_bench_kernel<<<sms, 32 * (nw + 1), smem, st>>>(nw, mode, 8, mma_n, 8, d);
    tmem_bench_kernel<<<sms, 32 * (nw + 1), smem, st>>>(nw, mode, iters, mma_n, mma_count, d);
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return focr_internal_fail(FOCR_ERR_CUDA, std::string("tmem bench: ") + cudaGetErrorString(e));
    std::vector<long long> h(sms * 2);
    cudaMemcpy(h.data(), d, sms * 16, cudaMemcpyDeviceToHost);
    cudaFree(d);
    std::vector<long long> a(sms), b(sms);
    for (int i = 0; i < sms; i++) a[i] = h[2 * i], b[i] = h[2 * i + 1];
    std::sort(a.begin(), a.end());
    std::sort(b.begin(), b.end());
    *cycles_per_round = (double)a[sms / 2] / iters;
    *cycles_per_mma = mma_n > 0 ? (double)b[sms / 2] / ((double)mma_count * 7) : 0.0;
    return FOCR_OK;
}

// cycles per hand-shake round trip (see pingpong_kernel), median over SMs
extern "C" int focr_bench_pingpong(int device, int wait_kind, int use_commit, int nwait, int iters, double *cycles_per_round)
{
    using namespace focr;
    if (device < 0 || !cycles_per_round || wait_kind < 0 || wait_kind > 2 || nwait < 1 || nwait > 8 || iters < 1)
        return focr_internal_fail(FOCR_ERR_ARG, "focr_bench_pingpong: bad argument");
    if (cudaSetDevice(device) != cudaSuccess) return focr_internal_fail(FOCR_ERR_CUDA, "cudaSetDevice");
    cudaStream_t st = nullptr;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    long long *d = nullptr;
    if (cudaMalloc((void **)&d, sms * 8) != cudaSuccess) return focr_internal_fail(FOCR_ERR_CUDA, "cudaMalloc");
    pingpong_kernel<<<sms, 32 * 9, 0, st>>>(wait_kind, use_commit, nwait, iters, d);
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return focr_internal_fail(FOCR_ERR_CUDA, std::string("pingpong: ") + cudaGetErrorString(e));
    std::vector<long long> h(sms);
    cudaMemcpy(h.data(), d, sms * 8, cudaMemcpyDeviceToHost);
    cudaFree(d);
    std::sort(h.begin(), h.end());
    *cycles_per_round = (double)h[sms / 2] / iters;
    return FOCR_OK;
}
