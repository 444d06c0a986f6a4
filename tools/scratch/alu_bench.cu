// throughput of candidate "screen" instructions: cycles per warp-instruction per SMSP with 1, 2, 4 warps per SMSP
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(long long *out, float seed, int iters)
{
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = seed * (i + 1) + threadIdx.x;
    uint32_t u[32];
#pragma unroll
    for (int i = 0; i < 32; i++) u[i] = __float_as_uint(v[i]);
    float acc = 0.f;
    uint32_t iacc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        if (OP == 0) {   // 16 x FMNMX3 tree over 32 values
            float t[11];
#pragma unroll
            for (int j = 0; j < 10; j++) asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(t[j]) : "f"(v[3 * j]), "f"(v[3 * j + 1]), "f"(v[3 * j + 2]));
            t[10] = fmaxf(v[30], v[31]);
            float a, b, c, mx;
            asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(a) : "f"(t[0]), "f"(t[1]), "f"(t[2]));
            asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(b) : "f"(t[3]), "f"(t[4]), "f"(t[5]));
            asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(c) : "f"(t[6]), "f"(t[7]), "f"(t[8]));
            asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(mx) : "f"(t[9]), "f"(t[10]), "f"(a));
            asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(mx) : "f"(mx), "f"(b), "f"(c));
            acc += mx;
            v[0] = acc;
        } else if (OP == 1) {  // 31 x FMNMX (2-input)
            float m = v[0];
#pragma unroll
            for (int j = 0; j < 16; j++) { float a; asm volatile("max.f32 %0, %1, %2;" : "=f"(a) : "f"(v[2 * j]), "f"(v[2 * j + 1])); v[j] = a; }
#pragma unroll
            for (int j = 0; j < 8; j++) { float a; asm volatile("max.f32 %0, %1, %2;" : "=f"(a) : "f"(v[2 * j]), "f"(v[2 * j + 1])); v[16 + j] = a; }
#pragma unroll
            for (int j = 0; j < 7; j++) { float a; asm volatile("max.f32 %0, %1, %2;" : "=f"(a) : "f"(v[16 + j]), "f"(v[17 + j])); m = fmaxf(m, a); }
            acc += m;
            v[0] = acc;
        } else if (OP == 2) {  // 16 x VIMNMX3 (signed int max3)
            int t[11];
#pragma unroll
            for (int j = 0; j < 10; j++) t[j] = __vimax3_s32((int)u[3 * j], (int)u[3 * j + 1], (int)u[3 * j + 2]);
            t[10] = max((int)u[30], (int)u[31]);
            int a = __vimax3_s32(t[0], t[1], t[2]), b = __vimax3_s32(t[3], t[4], t[5]), c = __vimax3_s32(t[6], t[7], t[8]);
            int mx = __vimax3_s32(t[9], t[10], a);
            mx = __vimax3_s32(mx, b, c);
            iacc += mx;
            u[0] = iacc;
        } else if (OP == 3) {  // 32 x FADD (v - T) then 16 x LOP3 and-tree on sign
            uint32_t d[32];
#pragma unroll
            for (int j = 0; j < 32; j++) d[j] = __float_as_uint(v[j] - 16711424.f);
            uint32_t r = 0xffffffffu;
#pragma unroll
            for (int j = 0; j < 32; j += 2) asm volatile("lop3.b32 %0, %1, %2, %3, 0x80;" : "=r"(r) : "r"(r), "r"(d[j]), "r"(d[j + 1]));
            iacc += r;
            v[0] = __uint_as_float(iacc);
        } else if (OP == 4) {  // 16 x add.f32x2 + 16 x LOP3
            uint32_t d[32];
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                unsigned long long a = ((unsigned long long)__float_as_uint(v[j + 1]) << 32) | __float_as_uint(v[j]), b = 0xCB7EFF00CB7EFF00ull, c;
                asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(c) : "l"(a), "l"(b));
                d[j] = (uint32_t)c, d[j + 1] = (uint32_t)(c >> 32);
            }
            uint32_t r = 0xffffffffu;
#pragma unroll
            for (int j = 0; j < 32; j += 2) asm volatile("lop3.b32 %0, %1, %2, %3, 0x80;" : "=r"(r) : "r"(r), "r"(d[j]), "r"(d[j + 1]));
            iacc += r;
            v[0] = __uint_as_float(iacc);
        } else if (OP == 5) {  // 32 x IADD3-pairs: (T-1-v) or-tree: 32 IADD + 16 LOP3
            uint32_t r = 0;
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                const uint32_t a = 0x4B7EFEFFu - u[j], b = 0x4B7EFEFFu - u[j + 1];
                asm volatile("lop3.b32 %0, %1, %2, %3, 0xfe;" : "=r"(r) : "r"(r), "r"(a), "r"(b));
            }
            iacc += r;
            u[0] = iacc;
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    if (acc == 123.456f || iacc == 0x12345u) out[1000] = 1;
}
int main()
{
    long long *d;
    cudaMalloc(&d, 8192 * 8);
    const int iters = 20000;
    const char *names[] = {"16 FMNMX3 tree", "31 FMNMX tree", "16 VIMNMX3 tree", "32 FADD + 16 LOP3", "16 FADD2 + 16 LOP3", "32 IADD + 16 LOP3"};
    for (int op = 0; op < 6; op++)
        for (int warps = 4; warps <= 16; warps *= 2) {   // warps per CTA = warps per SM (1, 2, 4 per SMSP)
            for (int rep = 0; rep < 2; rep++) {
                switch (op) {
                    case 0: k<0><<<148, warps * 32>>>(d, 1.5f, iters); break;
                    case 1: k<1><<<148, warps * 32>>>(d, 1.5f, iters); break;
                    case 2: k<2><<<148, warps * 32>>>(d, 1.5f, iters); break;
                    case 3: k<3><<<148, warps * 32>>>(d, 1.5f, iters); break;
                    case 4: k<4><<<148, warps * 32>>>(d, 1.5f, iters); break;
                    case 5: k<5><<<148, warps * 32>>>(d, 1.5f, iters); break;
                }
                cudaDeviceSynchronize();
            }
            long long h[148];
            cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            double avg = 0;
            for (int i = 0; i < 148; i++) avg += h[i];
            avg /= 148;
            printf("%-22s warps/SMSP %d: %.1f cycles per 32-value screen per warp, %.1f per SMSP-screen\n", names[op], warps / 4, avg / iters, avg / iters / (warps / 4));
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
