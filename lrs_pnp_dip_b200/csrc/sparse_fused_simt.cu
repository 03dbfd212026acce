// Fused sparse-coding step on the implicit patch set, fp32 FFMA engine (bb = 8, n = 64).
//
// Replaces, for every selected 8x8 window of the unfolded matrix, the chain
//   get_image_block(X + lambda_1/mu_1)            main_LRS_PnP.py:259
//   mask = (blocks_copy == 0)                     :276-280
//   Coefs = ista(valid_pixel, pruned_D, ...)      :288-292 / ista.m:13-24 (soft threshold)
//   Phi_z[:, p] = D @ Coefs                       :294
// without materialising the patch matrix: the window is gathered straight from X/lambda_1/Yobs,
// the coefficient vector never leaves registers during the Nit iterations, and only Phi_z is
// written.  Exact fp32 arithmetic (no tensor cores): this engine is the parity anchor for the
// tcgen05 engine and serves the K != 256 shapes.
//
// Work decomposition: 256 threads = 64 patches x 4 lanes; lane q of a patch owns atoms
// [q*K/4, (q+1)*K/4) of alpha and pixels [16q, 16q+16) of the residual.  D sits in shared memory
// interleaved so that the 4 lanes of a patch read one contiguous 64-byte line per step.
#include "common.cuh"

namespace lrs {


template <int K>
__global__ void __launch_bounds__(256, 1) sparse_fused_simt_kernel(FusedParams prm) {
    constexpr int KT = K / 4;   // atoms per lane
    constexpr int Q4 = KT / 4;  // float4 per lane per pixel
    extern __shared__ float4 smem4[];
    float4* Dsm = smem4;                             // [64][Q4][4 lanes]
    float* rn = reinterpret_cast<float*>(Dsm + 64 * K / 4);  // ||D[i,:]||^2

    const int tid = threadIdx.x;
    for (int e = tid; e < 64 * (K / 4); e += 256) {
        int i = e / (K / 4), w = e % (K / 4), t4 = w / 4, q = w % 4;
        Dsm[e] = *reinterpret_cast<const float4*>(prm.D + (int64_t)i * K + q * KT + 4 * t4);
    }
    __syncthreads();
    if (tid < 64) {
        float s = 0.f;
        for (int w = 0; w < K / 4; ++w) {
            float4 d = Dsm[tid * (K / 4) + w];
            s = fmaf(d.x, d.x, s);
            s = fmaf(d.y, d.y, s);
            s = fmaf(d.z, d.z, s);
            s = fmaf(d.w, d.w, s);
        }
        rn[tid] = s;
    }
    __syncthreads();

    const unsigned FULL = 0xffffffffu;
    const int lane = tid & 31, warp = tid >> 5;
    const int q = lane & 3, pl = warp * 8 + (lane >> 2), lane_base = lane & ~3;
    const int64_t total = prm.p_end - prm.p_begin;
    const int64_t nR = prm.g.row.n, C = prm.g.C;

    for (int64_t grp = blockIdx.x; grp * 64 < total; grp += gridDim.x) {
        int64_t pi = grp * 64 + pl;
        bool valid = pi < total;
        int64_t p = prm.p_begin + (valid ? pi : total - 1);
        int64_t ci = p / nR, ri = p - ci * nR;
        int64_t rs = prm.g.row.start(ri), cs = prm.g.col.start(ci);

        float y[16];
        unsigned mbits = 0;
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            int64_t src = (rs + (t & 7)) * C + cs + 2 * q + (t >> 3);
            float v = __ldg(prm.X + src);
            if (prm.L) v = __fadd_rn(v, __fdiv_rn(__ldg(prm.L + src), prm.mu1));
            y[t] = v;
            if (__ldg(prm.Yobs + src) != 0.0f) mbits |= 1u << t;
        }
        float a;
        if (prm.a_patch) {
            a = __ldg(prm.a_patch + p);
        } else if (prm.a_table) {
            unsigned rowbits = __shfl_sync(FULL, mbits, lane_base) & 0xFFu;  // lane q=0 holds window column 0
            a = __ldg(prm.a_table + rowbits);
        } else {
            float s = 0.f;
#pragma unroll
            for (int t = 0; t < 16; ++t)
                if ((mbits >> t) & 1u) s += rn[16 * q + t];
            s += __shfl_xor_sync(FULL, s, 1);
            s += __shfl_xor_sync(FULL, s, 2);
            a = 4.0f * s;  // 2*(tr(H^T H) + tr(H^T H)), main_LRS_PnP_DIP_pro.py:190
        }
        const bool ok = a > 0.0f;
        const float inv_a = ok ? __fdiv_rn(1.0f, a) : 0.0f;
        const float T = ok ? __fdiv_rn(prm.lambda, __fmul_rn(2.0f, a)) : 0.0f;

        float al[KT];
#pragma unroll
        for (int k = 0; k < KT; ++k) al[k] = 0.f;
        float rr[16];

        for (int it = 0;; ++it) {
            // ---- D * alpha for the 64 pixels; lane q keeps pixels [16q, 16q+16) ----------------
            if (it == 0) {
#pragma unroll
                for (int t = 0; t < 16; ++t) rr[t] = 0.f;
            } else {
#pragma unroll 1
                for (int jj = 0; jj < 4; ++jj) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        float part[8];
#pragma unroll
                        for (int ii = 0; ii < 8; ++ii) {
                            const float4* drow = Dsm + ((16 * jj + 8 * h + ii) * Q4) * 4 + q;
                            float acc = 0.f;
#pragma unroll
                            for (int t4 = 0; t4 < Q4; ++t4) {
                                float4 d = drow[t4 * 4];
                                acc = fmaf(al[4 * t4 + 0], d.x, acc);
                                acc = fmaf(al[4 * t4 + 1], d.y, acc);
                                acc = fmaf(al[4 * t4 + 2], d.z, acc);
                                acc = fmaf(al[4 * t4 + 3], d.w, acc);
                            }
                            part[ii] = acc;
                        }
#pragma unroll
                        for (int ii = 0; ii < 8; ++ii) {
                            part[ii] += __shfl_xor_sync(FULL, part[ii], 1);
                            part[ii] += __shfl_xor_sync(FULL, part[ii], 2);
                            if (q == jj) rr[8 * h + ii] = part[ii];
                        }
                    }
                }
            }
            if (it == prm.Nit) break;  // rr = D alpha_final = Phi_z column (:294)
            // ---- masked, scaled residual ------------------------------------------------------
#pragma unroll
            for (int t = 0; t < 16; ++t) rr[t] = ((mbits >> t) & 1u) ? (y[t] - rr[t]) * inv_a : 0.f;
            // ---- alpha += D^T r ; soft ---------------------------------------------------------
#pragma unroll 1
            for (int jj = 0; jj < 4; ++jj) {
                float v[16];
#pragma unroll
                for (int t = 0; t < 16; ++t) v[t] = __shfl_sync(FULL, rr[t], lane_base + jj);
#pragma unroll
                for (int t = 0; t < 16; ++t) {
                    const float4* drow = Dsm + ((16 * jj + t) * Q4) * 4 + q;
#pragma unroll
                    for (int t4 = 0; t4 < Q4; ++t4) {
                        float4 d = drow[t4 * 4];
                        al[4 * t4 + 0] = fmaf(v[t], d.x, al[4 * t4 + 0]);
                        al[4 * t4 + 1] = fmaf(v[t], d.y, al[4 * t4 + 1]);
                        al[4 * t4 + 2] = fmaf(v[t], d.z, al[4 * t4 + 2]);
                        al[4 * t4 + 3] = fmaf(v[t], d.w, al[4 * t4 + 3]);
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < KT; ++k) al[k] = soft_thr(al[k], T);
        }
        if (valid) {
#pragma unroll
            for (int t = 0; t < 16; ++t) prm.phi[(int64_t)(16 * q + t) * total + pi] = rr[t];
        }
    }
}

template <int K>
static int launch_simt(const FusedParams& prm, cudaStream_t st) {
    const char* fn = "lrs_sparse_step_fused_f32";
    size_t smem = (size_t)64 * K * 4 + 64 * 4;
    int rc = check_cuda(fn, cudaFuncSetAttribute(sparse_fused_simt_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)smem));
    if (rc != LRS_OK) return rc;
    int sms = device_sm_count();
    if (sms <= 0) return check_cuda(fn, cudaErrorNoDevice);
    int64_t groups = (prm.p_end - prm.p_begin + 63) / 64;
    unsigned grid = (unsigned)(groups < sms ? groups : sms);
    sparse_fused_simt_kernel<K><<<grid, 256, smem, st>>>(prm);
    note_launch();
    return check_cuda(fn, cudaGetLastError());
}


}  // namespace lrs

using namespace lrs;

extern "C" int lrs_sparse_step_fused_f32(const float* X_dev, const float* L_dev, float mu_1, const float* Yobs_dev,
                                         const float* D_dev, int K, const float* a_patch_dev, const float* a_table_dev,
                                         float lambda_ista, int Nit, int64_t R, int64_t C, int bb, int s,
                                         int64_t p_begin, int64_t p_end, float* phi_z_dev, int engine,
                                         lrs_stream_t stream) {
    const char* fn = "lrs_sparse_step_fused_f32";
    FusedParams prm;
    if (bb != 8) return fail_arg(fn, "the fused engines cover bb = 8 (n = 64); use lrs_ista_soft_f32 for other sizes");
    if (!make_geom(R, C, bb, s, prm.g)) return fail_arg(fn, "need 0 < bb <= min(R,C) and s > 0");
    if (!X_dev || !Yobs_dev || !D_dev || !phi_z_dev) return fail_arg(fn, "null pointer");
    if (L_dev && mu_1 == 0.0f) return fail_arg(fn, "mu_1 must be non-zero");
    if (p_begin < 0 || p_end > prm.g.P || p_begin > p_end) return fail_arg(fn, "patch range outside [0, P]");
    if (Nit < 0) return fail_arg(fn, "Nit < 0");
    if ((uintptr_t)D_dev % 16 != 0) return fail_arg(fn, "D must be 16-byte aligned");
    if (p_begin == p_end) return LRS_OK;
    prm.X = X_dev;
    prm.L = L_dev;
    prm.Yobs = Yobs_dev;
    prm.D = D_dev;
    prm.a_patch = a_patch_dev;
    prm.a_table = a_table_dev;
    prm.mu1 = mu_1;
    prm.lambda = lambda_ista;
    prm.Nit = Nit;
    prm.p_begin = p_begin;
    prm.p_end = p_end;
    prm.phi = phi_z_dev;
    cudaStream_t st = (cudaStream_t)stream;
    const bool dynamic_tiles = (engine & LRS_ENGINE_DYNAMIC_TILES) != 0;     // tcgen05 engine only; ignored by the FFMA kernel
    engine &= ~LRS_ENGINE_DYNAMIC_TILES;
    if (engine == LRS_ENGINE_TC || (engine == LRS_ENGINE_AUTO && sparse_fused_tc_supported(prm, K))) {
        if (!sparse_fused_tc_supported(prm, K))
            return fail_arg(fn, "tcgen05 engine needs K in {64,128,192,256}, Nit >= 1 and an sm_100 device");
        return sparse_fused_tc_launch(prm, K, dynamic_tiles, st);
    }
    if (engine != LRS_ENGINE_SIMT && engine != LRS_ENGINE_AUTO) return fail_arg(fn, "unknown engine");
    switch (K) {
        case 64: return launch_simt<64>(prm, st);
        case 128: return launch_simt<128>(prm, st);
        case 192: return launch_simt<192>(prm, st);
        case 256: return launch_simt<256>(prm, st);
        default: return fail_arg(fn, "fused engine supports K in {64,128,192,256}; use lrs_ista_soft_f32 otherwise");
    }
}
