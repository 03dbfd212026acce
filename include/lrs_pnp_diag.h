/* Diagnostics of the tcgen05 engine — NOT part of the product ABI.
 *
 * These entry points exist only in lrs_pnp_dip_b200/csrc/liblrs_pnp_diag.so, the diagnostics build of the library (the
 * same sources compiled with -DLRS_DIAGNOSTICS plus tc_probe.cu).  The product library liblrs_pnp.so (include/lrs_pnp.h)
 * exports none of them and carries no timing instrumentation.  They have no counterpart in the reference; tests/ and
 * scripts/ use them to pin hardware behaviour (operand layouts, rounding), to replay the fused kernel's tile order on
 * the host, and to read the barrier-wait counters of an instrumented launch. */
#ifndef LRS_PNP_DIAG_H
#define LRS_PNP_DIAG_H

#include "lrs_pnp.h"

#ifdef __cplusplus
extern "C" {
#endif

/* C[128,N] = A[128,Kd] * B[N,Kd]^T through one CTA's tensor core with the operand layouts of the fused
 * kernel: A from shared memory (a_in_tmem = 0) or TMEM (1); B K-major (b_mn_major = 0) or MN-major (1),
 * SWIZZLE_NONE; f16 = 0: kind::tf32 on the fp32 data, 1: kind::f16 on the data rounded to fp16. */
int lrs_tc_probe_f32(const float* A_dev, const float* B_dev, float* C_dev, int N, int Kd, int a_in_tmem,
                     int b_mn_major, int f16, lrs_stream_t stream);
/* Barrier-wait cycle counters of the fused tcgen05 kernel (block 0), filled when the environment variable
 * LRS_TC_TIMING is set at the first launch; out32_host = uint64 [32] in HOST memory. */
int lrs_tc_timing_read(unsigned long long* out32_host);
/* HOST-side replay (no GPU needed) of the tile order the fused tcgen05 kernel uses for the patch range
 * [p_begin, p_end) on a device with `sms` SMs: visits_host[p - p_begin] is incremented once per lane that owns patch p
 * (the caller zeroes it; a correct walk leaves every entry at 1), tiles_host receives the number of tiles.  bb = 8. */
int lrs_debug_tile_walk(int64_t R, int64_t C, int bb, int s, int64_t p_begin, int64_t p_end, int sms, int* visits_host,
                        int64_t* tiles_host);
/* HOST-side replay (no GPU needed) of ONE sweep of the parallel order of lrs_sym_eig_jacobi_f64 (csrc/jacobi_eig.cu) for C
 * columns: meets_host[p*C + q] (p < q; caller zeroes C*C ints) counts how often columns p and q are paired — a correct order
 * leaves every p < q at 1 —, conflicts_host the sub-rounds in which a column was used twice, subrounds_host their number. */
int lrs_debug_jacobi_schedule(int C, int* meets_host, int* conflicts_host, int* subrounds_host);
/* Cycle counts of MMA issue chains / TMEM load-store streams; out_dev = int64 [blocks*16]. */
int lrs_tc_microbench(int do_mma, int f16, int ts, int N, int nacc, int ldst, int depth, int reps, int blocks,
                      long long* out_dev, lrs_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* LRS_PNP_DIAG_H */
