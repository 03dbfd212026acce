// Bring-up probes for the tcgen05 engine.  lrs_tc_probe_f32 runs one 128 x N x Kd GEMM through exactly
// the descriptor encoders / operand layouts the fused kernel uses (A from shared memory or from TMEM,
// B K-major or MN-major, SWIZZLE_NONE; kind::tf32 on fp32 data or kind::f16 on data rounded to fp16), so
// the layouts and the operand rounding of the hardware are pinned against a CPU model in tests/.
// lrs_tc_microbench reports cycle counts of MMA issue chains and TMEM load/store streams.
//
// Facts established with these probes on B200:
//   * kind::tf32 drops the 13 low mantissa bits of its fp32 operands (truncation);
//   * tf32 operands in shared memory work K-major with SWIZZLE_NONE, but MN-major tf32 needs the
//     SWIZZLE_128B_BASE32B layout (SWIZZLE_NONE MN-major silently produces zeros);
//   * an MMA with A in TMEM costs >= 88 cycles, with A in shared memory >= 60 cycles and
//     >= (A bytes + B bytes)/128 cycles, whatever N is.
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/lrs_pnp_diag.h"

namespace lrs {
using namespace tc;

// Canonical SWIZZLE_NONE offsets (bytes) for an operand with EB-byte elements; chunk = 16 bytes.
template <int EB>
__device__ __forceinline__ uint32_t off_kmajor(int r, int k, uint32_t lbo, uint32_t sbo) {
    constexpr int T = 16 / EB;
    return (uint32_t)((k % T) * EB + (r % 8) * 16) + (uint32_t)(r / 8) * sbo + (uint32_t)(k / T) * lbo;
}
template <int EB>
__device__ __forceinline__ uint32_t off_mnmajor(int r, int k, uint32_t lbo, uint32_t sbo) {
    constexpr int T = 16 / EB;
    return (uint32_t)((r % T) * EB + (k % 8) * 16) + (uint32_t)(r / T) * sbo + (uint32_t)(k / 8) * lbo;
}

// C[128,N] = A[128,Kd] * B[N,Kd]^T.  F16: operands are converted to fp16 (round to nearest) first.
template <bool F16>
__global__ void __launch_bounds__(128, 1) tc_probe_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                          float* __restrict__ Cout, int N, int Kd, int a_in_tmem,
                                                          int b_mn_major) {
    constexpr int EB = F16 ? 2 : 4;       // element bytes
    constexpr int T = 16 / EB;            // elements per 16-byte chunk
    constexpr int KI = 32 / EB;           // K per MMA instruction (8 tf32 / 16 fp16)
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    uint8_t* Bs = smem;
    uint8_t* As = smem + (size_t)N * Kd * EB;
    const int tid = threadIdx.x, warp = tid >> 5;

    auto put = [&](uint8_t* base, uint32_t off, float v) {
        if (F16) *reinterpret_cast<__half*>(base + off) = __float2half_rn(v);
        else *reinterpret_cast<float*>(base + off) = v;
    };
    // K-major: 16-byte chunks along K are LBO apart, 8-row groups SBO apart.
    // MN-major: 16-byte chunks along N are SBO apart, 8-deep K groups LBO apart.
    const uint32_t b_lbo = b_mn_major ? (uint32_t)(N / T) * 128u : 128u;
    const uint32_t b_sbo = b_mn_major ? 128u : (uint32_t)(Kd / T) * 128u;
    for (int e = tid; e < N * Kd; e += 128) {
        int n = e / Kd, k = e % Kd;
        uint32_t off = b_mn_major ? off_mnmajor<EB>(n, k, b_lbo, b_sbo) : off_kmajor<EB>(n, k, b_lbo, b_sbo);
        put(Bs, off, B[e]);
    }
    const uint32_t a_lbo = 128u, a_sbo = (uint32_t)(Kd / T) * 128u;
    if (!a_in_tmem) {
        for (int e = tid; e < 128 * Kd; e += 128) {
            int m = e / Kd, k = e % Kd;
            put(As, off_kmajor<EB>(m, k, a_lbo, a_sbo), A[e]);
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
    const uint32_t a_col = 0, c_col = 256;  // A in columns [0, Kd*EB/4), C in [256, 256+N)

    if (a_in_tmem) {
        // 8 TMEM columns = one MMA K-slice (8 tf32 or 16 fp16, element 2c in the low half of column c)
        for (int k0 = 0; k0 < Kd; k0 += KI) {
            uint32_t r[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (F16) {
                    __half2 h = __floats2half2_rn(A[(size_t)tid * Kd + k0 + 2 * i], A[(size_t)tid * Kd + k0 + 2 * i + 1]);
                    r[i] = *reinterpret_cast<uint32_t*>(&h);
                } else {
                    r[i] = __float_as_uint(A[(size_t)tid * Kd + k0 + i]);
                }
            }
            tmem_st8(lane_addr + a_col + (k0 / KI) * 8, r);
        }
        tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp == 0) {
        const uint32_t idesc = F16 ? make_idesc_f16(128, N, b_mn_major != 0) : make_idesc_tf32(128, N, b_mn_major != 0);
        const uint32_t bs = smem_u32(Bs), as = smem_u32(As);
        const uint32_t leader = elect_one();
        for (int ks = 0; ks < Kd / KI; ++ks) {
            // one instruction consumes 2 chunks along K (K-major) / 32-byte-deep K (MN-major: KI/8 groups)
            uint64_t bdesc = b_mn_major ? make_smem_desc(bs + ks * (KI / 8) * b_lbo, b_lbo, b_sbo)
                                        : make_smem_desc(bs + 2 * ks * b_lbo, b_lbo, b_sbo);
            uint64_t adesc = make_smem_desc(as + 2 * ks * a_lbo, a_lbo, a_sbo);
            if (leader) {
                if (F16) {
                    if (a_in_tmem) mma_f16_ts(tbase + c_col, tbase + a_col + 8 * ks, bdesc, idesc, ks > 0);
                    else mma_f16_ss(tbase + c_col, adesc, bdesc, idesc, ks > 0);
                } else {
                    if (a_in_tmem) mma_tf32_ts(tbase + c_col, tbase + a_col + 8 * ks, bdesc, idesc, ks > 0);
                    else mma_tf32_ss(tbase + c_col, adesc, bdesc, idesc, ks > 0);
                }
            }
        }
        if (leader) mma_commit(&bar);
        __syncwarp();
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

// MMA issue-chain / TMEM stream micro-benchmark.
//   f16: kind::f16 (1) or kind::tf32 (0); ts: A from TMEM (1) or shared memory (0); N: MMA N; nacc: rotating
//   accumulators (power of two); ldst: 0 none | 1 epilogue warps stream tcgen05.ld | 2 tcgen05.st | 3 both;
//   depth: 32-column transfers in flight before the wait
__global__ void __launch_bounds__(384, 1) tc_microbench_kernel(int do_mma, int f16, int ts, int N, int nacc, int ldst,
                                                               int depth, int reps, long long* __restrict__ out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int e = tid; e < 48 * 1024; e += 384) reinterpret_cast<uint32_t*>(smem)[e] = 0x2c002c00u + (e & 255);
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
    if (do_mma && warp == 0) {
        const uint32_t idesc = f16 ? make_idesc_f16(128, N, false) : make_idesc_tf32(128, N, false);
        const uint32_t sb = smem_u32(smem);
        const uint32_t leader = elect_one();
        const uint32_t amask = (uint32_t)nacc - 1;
        __syncwarp();
        long long t0 = clock64();
#pragma unroll 8
        for (int i = 0; i < reps; ++i) {
            uint64_t bdesc = make_smem_desc(sb + (i & 7) * 256, 128, 1024);
            uint32_t acc = tbase + 128 + ((uint32_t)i & amask) * (uint32_t)N;  // accumulators in columns [128, 128+nacc*N)
            uint64_t adesc = make_smem_desc(sb + 64 * 1024 + (i & 7) * 256, 128, 1024);
            uint32_t atm = tbase + (i & 15) * 8;
            if (leader) {
                if (f16) {
                    if (ts) mma_f16_ts(acc, atm, bdesc, idesc, i >= nacc);
                    else mma_f16_ss(acc, adesc, bdesc, idesc, i >= nacc);
                } else {
                    if (ts) mma_tf32_ts(acc, atm, bdesc, idesc, i >= nacc);
                    else mma_tf32_ss(acc, adesc, bdesc, idesc, i >= nacc);
                }
            }
        }
        if (leader) mma_commit(&bar);
        __syncwarp();
        mbar_wait(&bar, 0);
        if (leader) out[blockIdx.x * 16 + 0] = clock64() - t0;
    }
    if (ldst && warp >= 4) {
        // epilogue-style traffic on columns [384, 512)
        const uint32_t lane_addr = tbase + ((uint32_t)((warp & 3) * 32) << 16) + 384u + ((warp >= 8) ? 64u : 0u);
        uint32_t r[32], q[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = q[i] = i;
        long long s0 = clock64();
        for (int i = 0; i < reps; ++i) {
            if (ldst & 1) {
                tmem_ld32(lane_addr, r);
                if (depth > 1) tmem_ld32(lane_addr + 32, q);
                tmem_wait_ld();
            }
            if (ldst & 2) {
                tmem_st32(lane_addr, r);
                if (depth > 1) tmem_st32(lane_addr + 32, q);
                tmem_wait_st();
            }
        }
        long long s1 = clock64();
        if ((tid & 31) == 0) out[blockIdx.x * 16 + warp] = s1 - s0 + ((r[3] ^ q[5]) == 0x7fffffff ? 1 : 0);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

}  // namespace lrs

using namespace lrs;

extern "C" {

int lrs_tc_probe_f32(const float* A_dev, const float* B_dev, float* C_dev, int N, int Kd, int a_in_tmem, int b_mn_major,
                     int f16, lrs_stream_t stream) {
    const char* fn = "lrs_tc_probe_f32";
    if (!A_dev || !B_dev || !C_dev) return fail_arg(fn, "null pointer");
    if (N < 16 || N > 256 || N % 16 || Kd < 16 || Kd > 256 || Kd % 16)
        return fail_arg(fn, "need 16<=N<=256 (mult of 16), 16<=Kd<=256 (mult of 16)");
    size_t eb = f16 ? 2 : 4;
    size_t smem = ((size_t)N * Kd + (size_t)128 * Kd) * eb;
    if (smem > 200 * 1024) return fail_arg(fn, "operands do not fit shared memory");
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    if (f16) {
        rc = check_cuda(fn, cudaFuncSetAttribute(tc_probe_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (rc != LRS_OK) return rc;
        tc_probe_kernel<true><<<1, 128, smem, st>>>(A_dev, B_dev, C_dev, N, Kd, a_in_tmem, b_mn_major);
    } else {
        rc = check_cuda(fn, cudaFuncSetAttribute(tc_probe_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (rc != LRS_OK) return rc;
        tc_probe_kernel<false><<<1, 128, smem, st>>>(A_dev, B_dev, C_dev, N, Kd, a_in_tmem, b_mn_major);
    }
    LRS_CHECK_LAUNCH(fn);
    return LRS_OK;
}

int lrs_tc_microbench(int do_mma, int f16, int ts, int N, int nacc, int ldst, int depth, int reps, int blocks,
                      long long* out_dev, lrs_stream_t stream) {
    const char* fn = "lrs_tc_microbench";
    if (reps <= 0 || blocks <= 0 || !out_dev || N < 16 || N > 256 || N % 16 || nacc < 1 || nacc * N > 256 ||
        (nacc & (nacc - 1)))
        return fail_arg(fn, "bad arguments");
    size_t smem = 192 * 1024;
    int rc = check_cuda(fn, cudaFuncSetAttribute(tc_microbench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (rc != LRS_OK) return rc;
    tc_microbench_kernel<<<blocks, 384, smem, (cudaStream_t)stream>>>(do_mma, f16, ts, N, nacc, ldst, depth, reps, out_dev);
    LRS_CHECK_LAUNCH(fn);
    return LRS_OK;
}

}  // extern "C"
