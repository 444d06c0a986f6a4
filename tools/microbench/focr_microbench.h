/*
 * focr_microbench.h -- measurement aids for DESIGN.md section 4.2 (tools/microbench/libfocr_microbench.so).
 * NOT part of the product ABI (include/focr_b200.h): tcgen05 / TMEM / mbarrier micro-benchmarks that bench.py
 * and tools/ use to measure the roofline denominators on the box they run on.  Every entry returns 0 on success
 * (1 = CUDA error, 2 = bad argument; message in focr_microbench_last_error()).
 */
#ifndef FOCR_MICROBENCH_H
#define FOCR_MICROBENCH_H

#ifdef __cplusplus
extern "C" {
#endif

const char *focr_microbench_last_error(void);

/* tcgen05.mma kind::i8 (u8 x u8 -> s32): every SM issues iters*ksteps MMAs of M=128, N=n, K=32 into nacc rotating
 * accumulators; cycles_per_mma is the median over SMs, ms_total the CUDA-event time of the timed launch.
 * bench.py turns it into the measured int8 dense peak (MEASURED_PEAKS.json has no integer tensor peak). */
int focr_bench_umma_i8(int device, int n, int ksteps, int iters, int nacc, double *cycles_per_mma, double *ms_total);
/* the same stream as cta_group::2 MMAs (M = 256) issued by the leader of every CTA pair */
int focr_bench_umma_i8_2cta(int device, int n, int ksteps, int iters, int nacc, double *cycles_per_mma, double *ms_total);
double focr_bench_umma_issue_cycles(void);

/* cycles per hand-shake round trip "signal (tcgen05.commit or arrive) -> nwait warps wait and answer -> the signaller
 * waits" with the three ways of waiting on an mbarrier (0 try_wait + suspend hint, 1 try_wait, 2 test_wait) */
int focr_bench_pingpong(int device, int wait_kind, int use_commit, int nwait, int iters, double *cycles_per_round);

/* TMEM <-> register traffic (tcgen05.ld / tcgen05.st 32x32b.x32) from nw warps, optionally while another warp streams
 * MMAs of N = mma_n; mode 0 = ld, 1 = ld + st, 2 = st, 3 = ld + the screen's max tree */
int focr_bench_tmem(int device, int nw, int mode, int iters, int mma_n, int mma_count, double *cycles_per_round,
                    double *cycles_per_mma);

#ifdef __cplusplus
}
#endif
#endif
