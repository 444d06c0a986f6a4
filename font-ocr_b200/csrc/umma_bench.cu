// umma_bench.cu -- tcgen05.mma kind::i8 micro-benchmark (measurement aid, not on the product path).
//
// MEASURED_PEAKS.json has no integer tensor peak, and the correlation kernel's design hinges on how
// long ONE tcgen05.mma of a given N takes (scan_tc.cu issues many small ones).  Every SM runs one CTA
// that issues `iters` x `ksteps` MMAs (M = 128, K = 32, u8 x u8 -> s32, K-major no-swizzle operands in
// shared memory exactly like scan_tc.cu) into `nacc` rotating accumulators and times them with clock64.
#include <algorithm>
#include <cstdint>
#include <string>
#include <vector>

#include "common.cuh"

int focr_internal_fail(int code, const std::string &msg);
int focr_internal_device(const focr_ctx *ctx);
extern "C" void *focr_ctx_stream(focr_ctx *ctx);

namespace focr {

__device__ __forceinline__ uint32_t ub_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) umma_i8_bench_kernel(int n, int ksteps, int iters, int nacc, long long *cycles)
{
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
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ub_smem_u32(&tmem_ptr)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_ptr;
    if (warp == 1) {
        const uint32_t idesc = (2u << 4) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
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
        if (elected) {
            for (int it = 0; it < iters; it++) {
                const uint32_t d = tmem_base + (uint32_t)((it % nacc) * nbs);
                for (int k = 0; k < ksteps; k++) {
                    const uint64_t ad = ((uint64_t)desc_hi << 32) | (a0 + (uint32_t)k * 2 * (2048 >> 4));
                    const uint64_t bd = ((uint64_t)desc_hi << 32) | (b0 + (uint32_t)k * 2 * ((uint32_t)n * 16u >> 4));
                    asm volatile(
                        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                        "l"(ad), "l"(bd), "r"(idesc), "r"(k)
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
        t1 = clock64();
        if ((threadIdx.x & 31) == 0) cycles[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

}  // namespace focr

// n: MMA N (multiple of 16, <= 256); returns the median over SMs of cycles per tcgen05.mma and the wall time
extern "C" int focr_bench_umma_i8(focr_ctx *ctx, int n, int ksteps, int iters, int nacc, double *cycles_per_mma,
                                  double *ms_total)
{
    using namespace focr;
    if (!ctx || !cycles_per_mma || !ms_total || n < 16 || n > 256 || (n & 15) || ksteps < 1 || ksteps > 16 || iters < 1 ||
        nacc < 1 || nacc * ((n + 31) & ~31) > 512)
        return focr_internal_fail(FOCR_ERR_ARG, "focr_bench_umma_i8: bad argument");
    if (cudaSetDevice(focr_internal_device(ctx)) != cudaSuccess) return focr_internal_fail(FOCR_ERR_CUDA, "cudaSetDevice");
    cudaStream_t st = (cudaStream_t)focr_ctx_stream(ctx);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, focr_internal_device(ctx));
    long long *d = nullptr;
    if (cudaMalloc((void **)&d, sms * 8) != cudaSuccess) return focr_internal_fail(FOCR_ERR_CUDA, "cudaMalloc");
    const size_t smem = (size_t)ksteps * 2 * (128 + n) * 16 + 1024;
    if (cudaFuncSetAttribute(umma_i8_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return focr_internal_fail(FOCR_ERR_CUDA, "cudaFuncSetAttribute");
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    umma_i8_bench_kernel<<<sms, 128, smem, st>>>(n, ksteps, 8, nacc, d);  // warm-up
    cudaEventRecord(e0, st);
    umma_i8_bench_kernel<<<sms, 128, smem, st>>>(n, ksteps, iters, nacc, d);
    cudaEventRecord(e1, st);
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return focr_internal_fail(FOCR_ERR_CUDA, std::string("umma bench: ") + cudaGetErrorString(e));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> h(sms);
    cudaMemcpy(h.data(), d, sms * 8, cudaMemcpyDeviceToHost);
    cudaFree(d);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    std::sort(h.begin(), h.end());
    *cycles_per_mma = (double)h[sms / 2] / ((double)iters * ksteps);
    *ms_total = ms;
    return FOCR_OK;
}
