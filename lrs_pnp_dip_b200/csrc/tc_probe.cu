// Bring-up probes for the tcgen05 engine.  lrs_tc_probe_f32 runs one 128 x N x Kd TF32 GEMM through
// exactly the descriptor encoders / operand layouts the fused kernel uses (A from shared memory or
// from TMEM, B K-major or MN-major, SWIZZLE_NONE), so the layouts and the TF32 operand rounding of
// the hardware can be pinned against a CPU model in tests/.  lrs_tc_microbench reports cycle counts
// of MMA issue chains and TMEM load/store streams (DESIGN.md uses them to budget the pipeline).
#include "common.cuh"
#include "tc_common.cuh"

namespace lrs {
using namespace tc;

// C[128,N] = A[128,Kd] * B[N,Kd]^T
__global__ void __launch_bounds__(128, 1) tc_probe_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                          float* __restrict__ Cout, int N, int Kd, int a_in_tmem,
                                                          int b_mn_major) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    float* Bs = reinterpret_cast<float*>(smem);                 // N*Kd floats
    float* As = reinterpret_cast<float*>(smem + (size_t)N * Kd * 4);  // 128*Kd floats (SS mode)
    const int tid = threadIdx.x, warp = tid >> 5;

    // B -> canonical no-swizzle layout
    const uint32_t b_lbo = b_mn_major ? (uint32_t)(N / 4) * 128u : 128u;
    const uint32_t b_sbo = b_mn_major ? 128u : (uint32_t)(Kd / 4) * 128u;
    for (int e = tid; e < N * Kd; e += 128) {
        int n = e / Kd, k = e % Kd;
        uint32_t off = b_mn_major ? (uint32_t)((n % 4) * 4 + (k % 8) * 16) + (n / 4) * b_sbo + (k / 8) * b_lbo
                                  : (uint32_t)((k % 4) * 4 + (n % 8) * 16) + (n / 8) * b_sbo + (k / 4) * b_lbo;
        Bs[off / 4] = B[e];
    }
    const uint32_t a_lbo = 128u, a_sbo = (uint32_t)(Kd / 4) * 128u;
    if (!a_in_tmem) {
        for (int e = tid; e < 128 * Kd; e += 128) {
            int m = e / Kd, k = e % Kd;
            uint32_t off = (uint32_t)((k % 4) * 4 + (m % 8) * 16) + (m / 8) * a_sbo + (k / 4) * a_lbo;
            As[off / 4] = A[e];
        }
    }
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmem_base_s;
    const uint32_t lane_addr = tbase + ((uint32_t)(warp * 32) << 16);
    const uint32_t a_col = 0, c_col = 256;  // A in columns [0,Kd), C in [256, 256+N)

    if (a_in_tmem) {
        for (int k0 = 0; k0 < Kd; k0 += 8) {
            uint32_t r[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) r[i] = __float_as_uint(A[(size_t)tid * Kd + k0 + i]);
            tmem_st8(lane_addr + a_col + k0, r);
        }
        tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (tid == 0) {
        const uint32_t idesc = make_idesc_tf32(128, N, b_mn_major != 0);
        const uint32_t bs = smem_u32(Bs), as = smem_u32(As);
        for (int ks = 0; ks < Kd / 8; ++ks) {
            uint64_t bdesc = b_mn_major ? make_smem_desc(bs + ks * b_lbo, b_lbo, b_sbo)
                                        : make_smem_desc(bs + 2 * ks * b_lbo, b_lbo, b_sbo);
            if (a_in_tmem) {
                mma_tf32_ts(tbase + c_col, tbase + a_col + 8 * ks, bdesc, idesc, ks > 0);
            } else {
                uint64_t adesc = make_smem_desc(as + 2 * ks * a_lbo, a_lbo, a_sbo);
                mma_tf32_ss(tbase + c_col, adesc, bdesc, idesc, ks > 0);
            }
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    for (int n0 = 0; n0 < N; n0 += 8) {
        uint32_t r[8];
        tmem_ld8(lane_addr + c_col + n0, r);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 8; ++i) Cout[(size_t)tid * N + n0 + i] = __uint_as_float(r[i]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

// mode: 0 TS N=64 | 1 TS N=256 | 2 SS N=64 | 3 SS N=256 | 4 tmem ld x32 (8 warps) | 5 tmem st x32 (8 warps)
//       6 = mode 0 with concurrent ld/st traffic from the 8 epilogue warps
__global__ void __launch_bounds__(384, 1) tc_microbench_kernel(int mode, int reps, long long* __restrict__ out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int e = tid; e < 48 * 1024; e += 384) reinterpret_cast<float*>(smem)[e] = 0.001f * (e & 255);
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmem_base_s;
    long long t0 = clock64(), t1 = t0;
    const bool do_mma = (mode <= 3 || mode == 6), do_ldst = (mode >= 4);
    if (do_mma && tid == 0) {
        const bool ts = (mode == 0 || mode == 1 || mode == 6);
        const int N = (mode == 1 || mode == 3) ? 256 : 64;
        const uint32_t idesc = make_idesc_tf32(128, N, false);
        const uint32_t sb = smem_u32(smem);
        t0 = clock64();
        for (int i = 0; i < reps; ++i) {
            uint64_t bdesc = make_smem_desc(sb + (i & 7) * 256, 128, 1024);
            if (ts) {
                mma_tf32_ts(tbase + 256, tbase + (i & 15) * 8, bdesc, idesc, i > 0);
            } else {
                uint64_t adesc = make_smem_desc(sb + 64 * 1024 + (i & 7) * 256, 128, 1024);
                mma_tf32_ss(tbase + 256, adesc, bdesc, idesc, i > 0);
            }
        }
        mma_commit(&bar);
        mbar_wait(&bar, 0);
        t1 = clock64();
        out[blockIdx.x * 16 + 0] = t1 - t0;
    }
    if (do_ldst && warp >= 4) {
        const uint32_t lane_addr = tbase + ((uint32_t)((warp & 3) * 32) << 16) + ((warp >= 8) ? 64u : 0u);
        uint32_t r[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = i;
        long long s0 = clock64();
        for (int i = 0; i < reps; ++i) {
            if (mode == 4 || mode == 6) {
                tmem_ld32(lane_addr + (i & 1) * 32, r);
                tmem_wait_ld();
            }
            if (mode == 5 || mode == 6) {
                tmem_st32(lane_addr + (i & 1) * 32, r);
                tmem_wait_st();
            }
        }
        long long s1 = clock64();
        if ((tid & 31) == 0) out[blockIdx.x * 16 + warp] = s1 - s0 + (r[3] == 0x7fffffff ? 1 : 0);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

}  // namespace lrs

using namespace lrs;

extern "C" {

int lrs_tc_probe_f32(const float* A_dev, const float* B_dev, float* C_dev, int N, int Kd, int a_in_tmem, int b_mn_major,
                     lrs_stream_t stream) {
    const char* fn = "lrs_tc_probe_f32";
    if (!A_dev || !B_dev || !C_dev) return fail_arg(fn, "null pointer");
    if (N < 16 || N > 256 || N % 16 || Kd < 8 || Kd > 256 || Kd % 8) return fail_arg(fn, "need 16<=N<=256 (mult of 16), 8<=Kd<=256 (mult of 8)");
    size_t smem = (size_t)N * Kd * 4 + (size_t)128 * Kd * 4;
    if (smem > 200 * 1024) return fail_arg(fn, "operands do not fit shared memory");
    int rc = check_cuda(fn, cudaFuncSetAttribute(tc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (rc != LRS_OK) return rc;
    tc_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A_dev, B_dev, C_dev, N, Kd, a_in_tmem, b_mn_major);
    LRS_CHECK_LAUNCH(fn);
    return LRS_OK;
}

int lrs_tc_microbench(int mode, int reps, int blocks, long long* out_dev, lrs_stream_t stream) {
    const char* fn = "lrs_tc_microbench";
    if (mode < 0 || mode > 6 || reps <= 0 || blocks <= 0 || !out_dev) return fail_arg(fn, "bad arguments");
    size_t smem = 192 * 1024;
    int rc = check_cuda(fn, cudaFuncSetAttribute(tc_microbench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (rc != LRS_OK) return rc;
    tc_microbench_kernel<<<blocks, 384, smem, (cudaStream_t)stream>>>(mode, reps, out_dev);
    LRS_CHECK_LAUNCH(fn);
    return LRS_OK;
}

}  // extern "C"
