// Fused sparse-coding step on the implicit patch set — tcgen05 / TMEM engine (bb = 8, n = 64, K = 256).
//
// Same contract as the FFMA engine (sparse_fused_simt.cu): for every selected 8x8 window of the unfolded
// matrix run Nit soft-ISTA iterations against D and write Phi_z = D alpha (main_LRS_PnP.py:259-303,
// ista.m:13-24).  Here the two contractions of every iteration run on the 5th-generation tensor cores:
//
//   tile        128 consecutive patches (reference order) = the 128 TMEM lanes = MMA M
//   GEMM-B      G[128 x 256]  = alpha + r D        accumulated IN PLACE onto the fp32 state in TMEM
//   epilogue    alpha = soft(G, T)                 tcgen05.ld -> registers -> tcgen05.st
//   GEMM-A      Da[128 x 64]  = alpha D^T          A operand = alpha pieces staged in TMEM (TS form)
//   epilogue    r = m .* (y - Da) / a              written back to TMEM as the next A operand
//
// so alpha (128 KB per tile) never leaves TMEM during the Nit iterations; HBM sees 3 gathers and one
// Phi_z store per patch.
//
// fp32 accuracy on the tensor cores: every fp32 operand x is split into two fp16 pieces x = x1 + x2
// (x1 = fp16(x), x2 = fp16(x - x1): 22 significant bits) and a product uses three MMAs
// a1 b1 + a2 b1 + a1 b2 accumulated in fp32 — the 3-pass split-precision scheme, here on kind::f16 rather
// than kind::tf32 because (measured on B200, tests/test_gpu_tc.py):
//   * D is the B operand of both GEMMs, K-major in one and MN-major in the other.  With 16-bit elements
//     the two SWIZZLE_NONE canonical layouts are transposes of each other, so ONE 64 KB copy of the two
//     D pieces serves both GEMMs.  MN-major tf32 operands only exist in the SWIZZLE_128B_BASE32B layout,
//     which would need a second copy: 2 x 128 KB > 227 KB of shared memory.
//   * an fp16 MMA covers K = 16 per instruction (tf32: 8) at the same cost: half the tensor-pipe time.
// Each patch is normalised by an exact power of two (max |y| -> [0.5, 1)) so the fp16 pieces never leave
// their range whatever the scale of the data; the result is scaled back exactly.
//
// Warp roles (384 threads, 1 CTA / SM, persistent over tiles):
//   warp 0      MMA issuer (whole warp runs the loop, one elected lane issues)
//   warp 1      TMEM allocator
//   warps 4-11  epilogue: warp w owns TMEM lanes 32*(w%4).., column half (w-4)/4 of every chunk
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace lrs {
using namespace tc;

namespace {

constexpr int KATOMS = 256;
constexpr int TILE = 128;
constexpr int NTHREADS = 384;
constexpr uint32_t COL_ALPHA = 0;    // [0,256)   fp32 alpha / GEMM-B accumulator
constexpr uint32_t COL_ACC = 256;    // [256,320) alpha1 D1 + alpha2 D1 ; [320,384) alpha1 D2
constexpr uint32_t COL_STG0 = 384;   // staging buffers: piece 1 in [+0,+32), piece 2 in [+32,+64)
constexpr uint32_t COL_STG1 = 448;   //   buffer 0 also carries the residual pieces between GEMM-A and GEMM-B
// The state kept in TMEM is at = a * alpha' (alpha' = alpha / 2^ex, the patch-normalised coefficients), so that the
// residual operand r = m .* (y' - D alpha') is O(1) whatever the step constant a is:
//     at <- soft(at + r D, lambda' / 2),      D alpha' = (D at) / a
// (soft is positively homogeneous: a * soft(g, T) = soft(a g, a T), and a T = lambda / 2).
constexpr float S_ALPHA = 4.0f;      // state pieces    = fp16(4 at)
constexpr float S_D = 4.0f;          // D pieces        = fp16(4 D)
constexpr float S_R = 0.25f;         // residual pieces = fp16(r / 4)  (S_R * S_D = 1: GEMM-B lands in the state's units)

// D pieces in shared memory, one copy for both GEMMs (bytes):
//   (k%8)*2 + (i%8)*16 + (8*piece + i/8)*128 + (k/8)*2048       i = pixel (row of D), k = atom
constexpr uint32_t D_SMEM_BYTES = 64 * 1024;
constexpr uint32_t D_SK = 2048, D_SI = 128;

struct __align__(8) Shared {
    uint64_t bar_R, bar_B, bar_S[4], bar_A[4];
    uint32_t tmem_base;
    float xmax[2][TILE];      // per-patch partial max |y| of the two column halves
    float xsum[2][TILE];      // per-patch partial sum of valid row norms (in-kernel 4||H||_F^2)
    uint32_t xrow[TILE];      // validity bits of window column 0 (pixels 0..7)
    float rn[64];             // ||D[i,:]||^2
};

__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_h2(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }

// x (already scaled) -> two fp16 pieces, two values per 32-bit word
__device__ __forceinline__ void split_pair(float x0, float x1, uint32_t& p1, uint32_t& p2) {
    __half2 h = __floats2half2_rn(x0, x1);
    float2 f = __half22float2(h);
    __half2 l = __floats2half2_rn(x0 - f.x, x1 - f.y);
    p1 = pack_h2(h);
    p2 = pack_h2(l);
}

__global__ void __launch_bounds__(NTHREADS, 1) sparse_fused_tc_kernel(FusedParams prm) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* Dsm = smem;
    Shared& sh = *reinterpret_cast<Shared*>(smem + D_SMEM_BYTES);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t total = prm.p_end - prm.p_begin;
    const int64_t ntiles = (total + TILE - 1) / TILE;
    const int Nit = prm.Nit;

    // ---- one-time setup: D -> fp16 pieces, barriers, TMEM -------------------------------------------
    for (int e = tid; e < 64 * KATOMS; e += NTHREADS) {
        int i = e / KATOMS, k = e % KATOMS;
        float v = prm.D[e] * S_D;
        __half h1 = __float2half_rn(v);
        __half h2 = __float2half_rn(v - __half2float(h1));
        uint32_t off = (uint32_t)((k % 8) * 2 + (i % 8) * 16) + (uint32_t)(i / 8) * D_SI + (uint32_t)(k / 8) * D_SK;
        *reinterpret_cast<__half*>(Dsm + off) = h1;
        *reinterpret_cast<__half*>(Dsm + off + 8 * D_SI) = h2;
    }
    if (tid < 64) {
        float s = 0.f;
        for (int k = 0; k < KATOMS; ++k) {
            float d = prm.D[tid * KATOMS + k];
            s = fmaf(d, d, s);
        }
        sh.rn[tid] = s;
    }
    if (tid == 0) {
        mbar_init(&sh.bar_R, 8);
        mbar_init(&sh.bar_B, 1);
        for (int j = 0; j < 4; ++j) {
            mbar_init(&sh.bar_S[j], 8);
            mbar_init(&sh.bar_A[j], 1);
        }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(&sh.tmem_base, 512);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = sh.tmem_base;

    if (warp == 0) {
        // ================================ MMA issuer ================================
        const uint32_t leader = elect_one();
        const uint32_t dbase = smem_u32(Dsm);
        const uint32_t idescB = make_idesc_f16(128, 256, /*b_mn_major=*/true);
        const uint32_t idescA128 = make_idesc_f16(128, 128, false);
        const uint32_t idescA64 = make_idesc_f16(128, 64, false);
        // GEMM-B: B = D as (N = atoms, K = pixels), MN-major: 16-byte atom chunks SBO = D_SK apart, 8-pixel groups
        //         LBO = D_SI apart; k-step ks covers pixel groups 2ks, 2ks+1 of piece p.
        const uint64_t descB0 = make_smem_desc(dbase, /*lbo=*/D_SI, /*sbo=*/D_SK);
        // GEMM-A: B = D as (N = pixels [D1;D2], K = atoms), K-major: 16-byte atom chunks LBO = D_SK apart, 8-pixel
        //         groups SBO = D_SI apart; k-step covers atom groups 2g, 2g+1.
        const uint64_t descA0 = make_smem_desc(dbase, /*lbo=*/D_SK, /*sbo=*/D_SI);
        uint32_t gi = 0;
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            for (int it = 0; it < Nit; ++it, ++gi) {
                const uint32_t par = gi & 1;
                // ---- GEMM-B: alpha += r D ----
                mbar_wait(&sh.bar_R, par);
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint64_t d1 = descB0 + (uint64_t)(((0 * 8 + 2 * ks) * D_SI) >> 4);
                    const uint64_t d2 = descB0 + (uint64_t)(((1 * 8 + 2 * ks) * D_SI) >> 4);
                    const uint32_t r1 = tbase + COL_STG0 + 8 * ks, r2 = tbase + COL_STG0 + 32 + 8 * ks;
                    if (leader) {
                        mma_f16_ts(tbase + COL_ALPHA, r1, d1, idescB, !(it == 0 && ks == 0));
                        mma_f16_ts(tbase + COL_ALPHA, r2, d1, idescB, true);
                        mma_f16_ts(tbase + COL_ALPHA, r1, d2, idescB, true);
                    }
                }
                if (leader) mma_commit(&sh.bar_B);
                __syncwarp();
                // ---- GEMM-A: Da = alpha D^T, chunk by chunk as the soft-threshold epilogue releases them ----
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    mbar_wait(&sh.bar_S[j], par);
                    tc_fence_after();
                    const uint32_t stg = tbase + ((j & 1) ? COL_STG1 : COL_STG0);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint64_t d = descA0 + (uint64_t)(((8 * j + 2 * ks) * D_SK) >> 4);
                        if (leader) {
                            mma_f16_ts(tbase + COL_ACC, stg + 8 * ks, d, idescA128, !(j == 0 && ks == 0));  // a1 [D1;D2]
                            mma_f16_ts(tbase + COL_ACC, stg + 32 + 8 * ks, d, idescA64, true);               // a2 D1
                        }
                    }
                    if (j != 2 && leader) mma_commit(&sh.bar_A[j]);
                    __syncwarp();
                }
            }
        }
    } else if (warp >= 4) {
        // ================================ epilogue warps ================================
        const int q = warp & 3;               // TMEM lane quarter
        const int h = (warp - 4) >> 2;        // column half
        const int m = q * 32 + lane;          // patch within the tile = TMEM lane
        const uint32_t lane_addr = tbase + ((uint32_t)(q * 32) << 16);
        const int64_t nR = prm.g.row.n, C = prm.g.C;
        uint32_t gi = 0;
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            // ---- tile prologue: gather my 32 pixels (window columns 4h..4h+3), mask, step constant, scale ----
            const int64_t pi = tile * TILE + m;
            const bool valid = pi < total;
            const int64_t p = prm.p_begin + (valid ? pi : total - 1);
            const int64_t ci = p / nR, ri = p - ci * nR;
            const int64_t rs = prm.g.row.start(ri), cs = prm.g.col.start(ci);
            float ysc[32];
            uint32_t mbits = 0;
            float amax = 0.f, nsum = 0.f;
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                const int64_t src = (rs + (c & 7)) * C + cs + 4 * h + (c >> 3);
                float v = __ldg(prm.X + src);
                if (prm.L) v = __fadd_rn(v, __fdiv_rn(__ldg(prm.L + src), prm.mu1));
                const bool ok = __ldg(prm.Yobs + src) != 0.0f;
                ysc[c] = v;
                if (ok) {
                    mbits |= 1u << c;
                    amax = fmaxf(amax, fabsf(v));
                    nsum += sh.rn[32 * h + c];
                }
            }
            sh.xmax[h][m] = amax;
            sh.xsum[h][m] = nsum;
            if (h == 0) sh.xrow[m] = mbits & 0xFFu;
            epi_barrier();
            amax = fmaxf(sh.xmax[0][m], sh.xmax[1][m]);
            float a;
            if (prm.a_patch) a = __ldg(prm.a_patch + p);
            else if (prm.a_table) a = __ldg(prm.a_table + sh.xrow[m]);
            else a = 4.0f * (sh.xsum[0][m] + sh.xsum[1][m]);
            epi_barrier();  // exchange buffers are free for the next tile
            const bool ok_a = a > 0.0f;
            const float inv_a = ok_a ? __fdiv_rn(1.0f, a) : 0.0f;
            int ex = 0;
            if (amax > 0.0f) (void)frexpf(amax, &ex);          // amax = f * 2^ex, f in [0.5, 1)
            const float dn = ldexpf(1.0f, -ex), up = ldexpf(1.0f, ex);
            const float Tn = ok_a ? 0.5f * prm.lambda * dn : 0.0f;  // a * T = lambda / 2 (ista.m:17), normalised
            const float c1 = S_R * dn;                            // y -> scaled residual units
            const float c2 = inv_a * S_R / (S_ALPHA * S_D);       // acc (= S_ALPHA S_D a D alpha') -> scaled residual units
#pragma unroll
            for (int c = 0; c < 32; ++c) ysc[c] *= c1;

            for (int it = 0; it < Nit; ++it, ++gi) {
                const uint32_t par = gi & 1;
                // ---- residual: r = m .* (y - D alpha) / a  -> fp16 pieces in staging buffer 0 ----
                {
                    uint32_t p1[16], p2[16];
                    if (it > 0) {
                        uint32_t a0[32], a1[32];
                        mbar_wait(&sh.bar_A[3], par ^ 1);
                        tc_fence_after();
                        tmem_ld32(lane_addr + COL_ACC + 32 * h, a0);
                        tmem_ld32(lane_addr + COL_ACC + 64 + 32 * h, a1);
                        tmem_wait_ld();
#pragma unroll
                        for (int c = 0; c < 16; ++c) {
                            float s0 = __uint_as_float(a0[2 * c]) + __uint_as_float(a1[2 * c]);
                            float s1 = __uint_as_float(a0[2 * c + 1]) + __uint_as_float(a1[2 * c + 1]);
                            float r0 = ((mbits >> (2 * c)) & 1u) ? fmaf(-c2, s0, ysc[2 * c]) : 0.f;
                            float r1 = ((mbits >> (2 * c + 1)) & 1u) ? fmaf(-c2, s1, ysc[2 * c + 1]) : 0.f;
                            split_pair(r0, r1, p1[c], p2[c]);
                        }
                    } else {
#pragma unroll
                        for (int c = 0; c < 16; ++c) {
                            float r0 = ((mbits >> (2 * c)) & 1u) ? ysc[2 * c] : 0.f;
                            float r1 = ((mbits >> (2 * c + 1)) & 1u) ? ysc[2 * c + 1] : 0.f;
                            split_pair(r0, r1, p1[c], p2[c]);
                        }
                    }
                    tmem_st16(lane_addr + COL_STG0 + 16 * h, p1);
                    tmem_st16(lane_addr + COL_STG0 + 32 + 16 * h, p2);
                    tmem_wait_st();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sh.bar_R);
                }
                // ---- soft threshold, 4 chunks of 64 atoms; my 32 columns of each ----
                mbar_wait(&sh.bar_B, par);
                tc_fence_after();
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint32_t g[32], p1[16], p2[16];
                    const uint32_t col = COL_ALPHA + 64 * j + 32 * h;
                    tmem_ld32(lane_addr + col, g);
                    tmem_wait_ld();
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        float x0 = soft_thr(__uint_as_float(g[2 * c]), Tn);
                        float x1 = soft_thr(__uint_as_float(g[2 * c + 1]), Tn);
                        g[2 * c] = __float_as_uint(x0);
                        g[2 * c + 1] = __float_as_uint(x1);
                        split_pair(x0 * S_ALPHA, x1 * S_ALPHA, p1[c], p2[c]);
                    }
                    if (j >= 2) {  // staging buffer (j&1) is free once GEMM-A of chunk j-2 has completed
                        mbar_wait(&sh.bar_A[j - 2], par);
                        tc_fence_after();
                    }
                    const uint32_t stg = (j & 1) ? COL_STG1 : COL_STG0;
                    tmem_st32(lane_addr + col, g);
                    tmem_st16(lane_addr + stg + 16 * h, p1);
                    tmem_st16(lane_addr + stg + 32 + 16 * h, p2);
                    tmem_wait_st();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sh.bar_S[j]);
                }
            }
            // ---- Phi_z = D alpha_final (main_LRS_PnP.py:294): my 32 pixels of my patch ----
            {
                uint32_t a0[32], a1[32];
                mbar_wait(&sh.bar_A[3], (gi - 1) & 1);
                tc_fence_after();
                tmem_ld32(lane_addr + COL_ACC + 32 * h, a0);
                tmem_ld32(lane_addr + COL_ACC + 64 + 32 * h, a1);
                tmem_wait_ld();
                tc_fence_before();   // order these loads before the next tile's MMAs (via bar_R)
                const float sc = up * inv_a / (S_ALPHA * S_D);
                if (valid) {
#pragma unroll
                    for (int c = 0; c < 32; ++c)
                        prm.phi[(int64_t)(32 * h + c) * total + pi] = (__uint_as_float(a0[c]) + __uint_as_float(a1[c])) * sc;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tbase, 512);
}

}  // namespace

bool sparse_fused_tc_supported(const FusedParams& prm, int K) {
    if (K != KATOMS || prm.g.bb != 8 || prm.Nit < 1) return false;
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return false;
    return major == 10;
}

int sparse_fused_tc_launch(const FusedParams& prm, int K, cudaStream_t st) {
    const char* fn = "lrs_sparse_step_fused_f32";
    if (!sparse_fused_tc_supported(prm, K)) return fail_arg(fn, "tcgen05 engine needs K = 256, Nit >= 1 and an sm_100 device");
    const size_t smem = D_SMEM_BYTES + sizeof(Shared);
    int rc = check_cuda(fn, cudaFuncSetAttribute(sparse_fused_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (rc != LRS_OK) return rc;
    int sms = device_sm_count();
    if (sms <= 0) return check_cuda(fn, cudaErrorNoDevice);
    int64_t ntiles = (prm.p_end - prm.p_begin + TILE - 1) / TILE;
    unsigned grid = (unsigned)(ntiles < sms ? ntiles : sms);
    sparse_fused_tc_kernel<<<grid, NTHREADS, smem, st>>>(prm);
    note_launch();
    return check_cuda(fn, cudaGetLastError());
}

}  // namespace lrs
