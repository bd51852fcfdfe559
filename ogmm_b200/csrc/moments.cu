// K3: GMM M-step as segmented reductions (sm_100a).
//
//   gmm_moments_small_kernel   lib/utils.py:130-149  pi, mu (+ isotropic sigma) for D <= 4 (xyz)
//   gmm_moments_feat_kernel    lib/utils.py:130-140  mu for wide features, streamed once from HBM in
//                              their native (B,D,N) layout (models/gmmreg.py:26-27 passes a view)
//   softmax_moments_kernel     baseline/deepgmr.py:71-74 softmax over J fused with the M-step + sigma
//
// Quirks kept from the reference: npi = pi*N + 1e-5 (:138); sigma = sum_n gamma |x-mu|^2 / npi, NOT
// divided by D (:146-147); pi is the plain mean of whatever gamma holds (rows need not sum to 1).
#include <stdlib.h>

#include "common.cuh"

namespace ogmm {

// =====================================================================================================
// small-D kernel: one CTA per cloud, lanes over j, warps (and lane groups when J | 32) over n.
// =====================================================================================================
constexpr int kSmallThreads = 256;
constexpr int kSmallD = 4;          // D <= 4
constexpr int kSmallSlots = 4;      // J <= 128 per pass

template <bool kSigmaPass>
__device__ __forceinline__ void small_accumulate(const float* __restrict__ g, int64_t g_sn, int64_t g_sj,
                                                 const float* __restrict__ x, int64_t p_sn, int64_t p_sd, int N,
                                                 int J, int D, int j_base, const float* __restrict__ s_mu,
                                                 float (&acc)[kSmallSlots][kSmallD + 1]) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = kSmallThreads / 32;
    // lane layout: when J <= 32 divides 32, 32/J rows share a warp; otherwise one row per warp step
    const int rows_per_warp = (J <= 32 && (32 % J) == 0) ? 32 / J : 1;
    const int jl = rows_per_warp > 1 ? lane % J : lane;
    const int rsub = rows_per_warp > 1 ? lane / J : 0;
    for (int n = warp * rows_per_warp + rsub; n < N; n += NW * rows_per_warp) {
        float xv[kSmallD];
#pragma unroll
        for (int d = 0; d < kSmallD; ++d) xv[d] = d < D ? x[(int64_t)n * p_sn + d * p_sd] : 0.f;
#pragma unroll
        for (int s = 0; s < kSmallSlots; ++s) {
            const int j = j_base + jl + 32 * s;
            if ((rows_per_warp > 1 && s > 0) || j >= J) continue;
            const float gv = g[(int64_t)n * g_sn + (int64_t)j * g_sj];
            if constexpr (!kSigmaPass) {
                acc[s][0] += gv;
#pragma unroll
                for (int d = 0; d < kSmallD; ++d) acc[s][d + 1] = fmaf(gv, xv[d], acc[s][d + 1]);
            } else {
                float sq = 0.f;
#pragma unroll
                for (int d = 0; d < kSmallD; ++d) {
                    const float df = d < D ? xv[d] - s_mu[j * kSmallD + d] : 0.f;
                    sq = fmaf(df, df, sq);
                }
                acc[s][0] = fmaf(sq, gv, acc[s][0]);
            }
        }
    }
}

__global__ void __launch_bounds__(kSmallThreads)
gmm_moments_small_kernel(const float* __restrict__ gamma, int64_t g_sb, int64_t g_sn, int64_t g_sj,
                         const float* __restrict__ pts, int64_t p_sb, int64_t p_sn, int64_t p_sd,
                         int N, int J, int D, float* __restrict__ pi_out, float* __restrict__ mu_out,
                         float* __restrict__ sigma_out) {
    extern __shared__ __align__(16) float sm[];
    constexpr int NW = kSmallThreads / 32;
    // s_part [NW][128][5] | s_mu [J][4] | s_npi [J]
    float* s_part = sm;
    float* s_mu = s_part + NW * 128 * (kSmallD + 1);
    float* s_npi = s_mu + (size_t)J * kSmallD;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* g = gamma + (int64_t)b * g_sb;
    const float* x = pts + (int64_t)b * p_sb;
    const int rows_per_warp = (J <= 32 && (32 % J) == 0) ? 32 / J : 1;

    for (int pass = 0; pass < (sigma_out ? 2 : 1); ++pass) {
        for (int j_base = 0; j_base < J; j_base += 128) {
            float acc[kSmallSlots][kSmallD + 1];
#pragma unroll
            for (int s = 0; s < kSmallSlots; ++s)
#pragma unroll
                for (int d = 0; d <= kSmallD; ++d) acc[s][d] = 0.f;
            if (pass == 0) small_accumulate<false>(g, g_sn, g_sj, x, p_sn, p_sd, N, J, D, j_base, s_mu, acc);
            else small_accumulate<true>(g, g_sn, g_sj, x, p_sn, p_sd, N, J, D, j_base, s_mu, acc);
            // combine lane groups that share a column
            if (rows_per_warp > 1) {
                for (int off = 16; off >= J; off >>= 1)
#pragma unroll
                    for (int d = 0; d <= kSmallD; ++d) acc[0][d] += __shfl_xor_sync(kFull, acc[0][d], off);
            }
            __syncthreads();
#pragma unroll
            for (int s = 0; s < kSmallSlots; ++s) {
                const int jj = (rows_per_warp > 1 ? lane % J : lane) + 32 * s;
                if (rows_per_warp > 1 && (s > 0 || lane >= J)) continue;
                if (j_base + jj >= J) continue;
#pragma unroll
                for (int d = 0; d <= kSmallD; ++d) s_part[(warp * 128 + jj) * (kSmallD + 1) + d] = acc[s][d];
            }
            __syncthreads();
            for (int jj = tid; jj < 128 && j_base + jj < J; jj += kSmallThreads) {
                const int j = j_base + jj;
                float t[kSmallD + 1];
#pragma unroll
                for (int d = 0; d <= kSmallD; ++d) t[d] = 0.f;
                for (int w = 0; w < NW; ++w)
#pragma unroll
                    for (int d = 0; d <= kSmallD; ++d) t[d] += s_part[(w * 128 + jj) * (kSmallD + 1) + d];
                if (pass == 0) {
                    const float pi = __fdiv_rn(t[0], (float)N);
                    const float npi = __fadd_rn(__fmul_rn(pi, (float)N), 1e-5f);
                    pi_out[(int64_t)b * J + j] = pi;
                    s_npi[j] = npi;
#pragma unroll
                    for (int d = 0; d < kSmallD; ++d) {
                        const float m = __fdiv_rn(t[d + 1], npi);
                        s_mu[j * kSmallD + d] = m;
                        if (d < D) mu_out[((int64_t)b * J + j) * D + d] = m;
                    }
                } else {
                    const float sg = __fdiv_rn(t[0], s_npi[j]);
                    float* so = sigma_out + ((int64_t)b * J + j) * D * D;
                    for (int r = 0; r < D; ++r)
                        for (int c = 0; c < D; ++c) so[r * D + c] = r == c ? sg : 0.f;
                }
            }
            __syncthreads();
        }
    }
}

// =====================================================================================================
// wide-feature kernel.  out[j][d] = sum_n gamma[n][j] * f[n][d] / npi[j]  -- the one HBM-bound kernel.
//
// CTA = (tile of TD = 32*DT feature rows, one cloud), 256 threads.  Lane l of every warp owns rows
// d0 + l + 32 i (i < DT); the 8 warps split the points.  For one point n a thread reads DT feature
// values (conflict-free: the tile is stored TRANSPOSED, [n][d]) and the 16 responsibilities of that
// point as four broadcast 128-bit loads, and issues DT*JP/2 packed FFMA2 (fma.rn.f32x2; the feature
// value is the scalar-broadcast operand) -- ~0.19 shared-memory wavefronts per FMA instruction, so the
// loop is FP32-pipe bound, not LDS bound.  Features stream HBM -> registers (coalesced LDG.128 along n,
// L1 no-allocate) -> transposed STS with a row pitch == 2 (mod 8) words (conflict-free stores AND loads),
// double-buffered: the loads of stage s+1 are in flight while stage s is consumed.  Every feature value
// is read from HBM exactly once; gamma (64 KB per cloud, just written by the clustering kernel) is
// re-read from L2 once per feature-row tile.  Partial sums of the 8 warps are folded through shared
// memory at the end and written coalesced along d.
// =====================================================================================================
constexpr int kFeatThreads = 256;
constexpr int kTN = 32;                              // points per stage

__device__ __forceinline__ void ffma2(float2& d, const float2 a, const float b) {
    const float2 bb = make_float2(b, b);
    asm("fma.rn.f32x2 %0, %1, %2, %0;"
        : "+l"(*reinterpret_cast<unsigned long long*>(&d))
        : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&bb)));
}

template <int JP, int DT>
struct FeatCfg {
    static constexpr int TD = 32 * DT;
    static constexpr int PITCH = TD + 2;            // == 2 (mod 8)
    static constexpr int F_STAGE = kTN * PITCH;     // floats
    static constexpr int G_STAGE = kTN * JP;
    static constexpr int RED = (kFeatThreads / 64) * TD * JP;      // fold buffer: half of the warps at a time
    static constexpr int STAGES_FLOATS = 2 * (F_STAGE + G_STAGE);
    static constexpr size_t SMEM = sizeof(float) * ((STAGES_FLOATS > RED ? STAGES_FLOATS : RED) + 8 * JP + JP);
};

template <int JP, int DT>
__global__ void __launch_bounds__(kFeatThreads, (DT * JP <= 32) ? 3 : ((DT * JP <= 64) ? 2 : 1))
gmm_moments_feat_kernel(const float* __restrict__ gamma, int64_t g_sb, int64_t g_sn, int64_t g_sj,
                        const float* __restrict__ feats, int64_t f_sb, int64_t f_sn, int64_t f_sd,
                        int N, int J, int D, float* __restrict__ pi_out, float* __restrict__ mu_out,
                        int n_per_split, float* __restrict__ part_sum, float* __restrict__ part_gs) {
    using C = FeatCfg<JP, DT>;
    constexpr int TD = C::TD, PITCH = C::PITCH;
    extern __shared__ __align__(16) float sm[];
    float* s_f = sm;                                 // [2][kTN][PITCH]
    float* s_g = sm + 2 * C::F_STAGE;                // [2][kTN][JP]
    float* s_red = sm;                               // fold buffer aliases the stages (used after the loop)
    float* s_gs = sm + (C::STAGES_FLOATS > C::RED ? C::STAGES_FLOATS : C::RED);   // [8][JP] gamma column sums per warp
    float* s_npi = s_gs + 8 * JP;                    // [JP]

    const int b = blockIdx.y, d0 = blockIdx.x * TD;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // split mode (part_sum != nullptr): blockIdx.z owns points [n_lo, n_hi) and writes RAW sums; a fold kernel adds the
    // splits in order and normalises (few clouds with very many points: cloud x row-block items alone leave SMs idle)
    const int n_lo = part_sum ? (int)blockIdx.z * n_per_split : 0;
    const int n_hi = part_sum ? min(N, n_lo + n_per_split) : N;
    const float* g = gamma + (int64_t)b * g_sb;
    const float* f = feats + (int64_t)b * f_sb;

    const bool f_fast = (f_sn == 1) && ((f_sd & 3) == 0) && ((f_sb & 3) == 0) && ((reinterpret_cast<uintptr_t>(feats) & 15) == 0) && (d0 + TD <= D);
    const bool g_fast = (g_sj == 1) && (g_sn == J) && (J == JP) && ((g_sb & 3) == 0) && ((reinterpret_cast<uintptr_t>(gamma) & 15) == 0);

    float2 acc[DT][JP / 2];
#pragma unroll
    for (int i = 0; i < DT; ++i)
#pragma unroll
        for (int j = 0; j < JP / 2; ++j) acc[i][j] = make_float2(0.f, 0.f);
    constexpr int GC = (JP + 31) / 32;               // gamma columns per lane: j = lane + 32 c
    float gsum[GC];
#pragma unroll
    for (int c = 0; c < GC; ++c) gsum[c] = 0.f;

    // ---- stage loaders -------------------------------------------------------------------------------
    // fast path: unit u = tid + 256 k (k < DT) -> 4 consecutive points of one row: c4 = u%4, rsub = (u/4)%8,
    // block = u/32 -> (row group = block % (TD/8), chunk half = block / (TD/8)).  A warp covers 8 rows x 64 B.
    float4 fr[DT];
    constexpr int GQ = (kTN * JP / 4 + kFeatThreads - 1) / kFeatThreads;      // gamma float4s per thread per stage
    float4 gr[GQ];
    auto ldg_stage = [&](int n0) {
        if (f_fast && n0 + kTN <= n_hi) {
#pragma unroll
            for (int k = 0; k < DT; ++k) {
                const int u = tid + kFeatThreads * k;
                const int c4 = u & 3, rsub = (u >> 2) & 7, blk = u >> 5;
                const int row = (blk % (TD / 8)) * 8 + rsub, chunk = c4 + 4 * (blk / (TD / 8));
                fr[k] = ldg_stream4(f + (int64_t)(d0 + row) * f_sd + n0 + 4 * chunk);
            }
        }
        if (g_fast && n0 + kTN <= n_hi) {
#pragma unroll
            for (int q = 0; q < GQ; ++q)
                if (tid + kFeatThreads * q < kTN * JP / 4)
                    gr[q] = *reinterpret_cast<const float4*>(g + (int64_t)n0 * JP + 4 * (tid + kFeatThreads * q));
        }
    };
    auto sts_stage = [&](int n0, int slot) {
        float* fs = s_f + slot * C::F_STAGE;
        float* gs = s_g + slot * C::G_STAGE;
        if (f_fast && n0 + kTN <= n_hi) {
#pragma unroll
            for (int k = 0; k < DT; ++k) {
                const int u = tid + kFeatThreads * k;
                const int c4 = u & 3, rsub = (u >> 2) & 7, blk = u >> 5;
                const int row = (blk % (TD / 8)) * 8 + rsub, chunk = c4 + 4 * (blk / (TD / 8));
                float* dst = fs + (4 * chunk) * PITCH + row;
                dst[0] = fr[k].x; dst[PITCH] = fr[k].y; dst[2 * PITCH] = fr[k].z; dst[3 * PITCH] = fr[k].w;
            }
        } else {
            // generic: any strides, ragged tiles, zero fill
            const bool d_fastest = (f_sd == 1);
            for (int e = tid; e < kTN * TD; e += kFeatThreads) {
                int n, r;
                if (d_fastest) { n = e / TD; r = e - n * TD; } else { r = e / kTN; n = e - r * kTN; }
                float v = 0.f;
                if (n0 + n < n_hi && d0 + r < D) v = ldg_stream(f + (int64_t)(n0 + n) * f_sn + (int64_t)(d0 + r) * f_sd);
                fs[n * PITCH + r] = v;
            }
        }
        if (g_fast && n0 + kTN <= n_hi) {
#pragma unroll
            for (int q = 0; q < GQ; ++q)
                if (tid + kFeatThreads * q < kTN * JP / 4) *reinterpret_cast<float4*>(gs + 4 * (tid + kFeatThreads * q)) = gr[q];
        } else {
            for (int e = tid; e < kTN * JP; e += kFeatThreads) {
                const int n = e / JP, j = e - n * JP;
                float v = 0.f;
                if (n0 + n < n_hi && j < J) v = g[(int64_t)(n0 + n) * g_sn + (int64_t)j * g_sj];
                gs[e] = v;
            }
        }
    };

    // L2 prefetch kPF stages ahead: one 128-B line per feature row per stage, no registers held.  The demand
    // loads of the next stage then hit L2 (~300 cycles) instead of HBM (~800+), which one stage of FMAs covers.
    constexpr int kPF = 6;
    auto prefetch_stage = [&](int n0) {
        if (f_fast && n0 + kTN <= n_hi) {
            for (int r = tid; r < TD; r += kFeatThreads)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(f + (int64_t)(d0 + r) * f_sd + n0));
        }
    };
    const int n_stages = (max(n_hi - n_lo, 0) + kTN - 1) / kTN;
#pragma unroll 1
    for (int s = 1; s <= kPF && s < n_stages; ++s) prefetch_stage(n_lo + s * kTN);
    if (n_stages > 0) {
        ldg_stage(n_lo);
        sts_stage(n_lo, 0);
    }
    __syncthreads();
    for (int s = 0; s < n_stages; ++s) {
        const int slot = s & 1;
        if (s + 1 < n_stages) ldg_stage(n_lo + (s + 1) * kTN);
        if (s + 1 + kPF < n_stages) prefetch_stage(n_lo + (s + 1 + kPF) * kTN);
        const float* fs = s_f + slot * C::F_STAGE;
        const float* gs = s_g + slot * C::G_STAGE;
#pragma unroll
        for (int t = 0; t < kTN / 8; ++t) {
            const int n = warp + 8 * t;
            float fv[DT];
#pragma unroll
            for (int i = 0; i < DT; ++i) fv[i] = fs[n * PITCH + lane + 32 * i];
#pragma unroll
            for (int c = 0; c < GC; ++c)
                if (lane + 32 * c < JP) gsum[c] += gs[n * JP + lane + 32 * c];
            const float4* grow = reinterpret_cast<const float4*>(gs + n * JP);
#pragma unroll
            for (int q = 0; q < JP / 4; ++q) {
                const float4 gv = grow[q];
#pragma unroll
                for (int i = 0; i < DT; ++i) {
                    ffma2(acc[i][2 * q], make_float2(gv.x, gv.y), fv[i]);
                    ffma2(acc[i][2 * q + 1], make_float2(gv.z, gv.w), fv[i]);
                }
            }
        }
        if (s + 1 < n_stages) sts_stage(n_lo + (s + 1) * kTN, slot ^ 1);
        __syncthreads();
    }

    // ---- gamma column sums -> npi ------------------------------------------------------------------------
#pragma unroll
    for (int c = 0; c < GC; ++c)
        if (lane + 32 * c < JP) s_gs[warp * JP + lane + 32 * c] = gsum[c];
    __syncthreads();
    if (tid < JP) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += s_gs[w * JP + tid];
        const float pi = __fdiv_rn(t, (float)N);
        s_npi[tid] = __fadd_rn(__fmul_rn(pi, (float)N), 1e-5f);
        if (part_sum) {
            if (blockIdx.x == 0 && tid < J) part_gs[((int64_t)b * gridDim.z + blockIdx.z) * J + tid] = t;
        } else if (blockIdx.x == 0 && tid < J && pi_out) pi_out[(int64_t)b * J + tid] = pi;
    }
    // ---- fold the 8 warps' partial sums: 4 -> 0..3, then 2,3 -> 0,1, then 1 -> 0 (buffer aliases the stages) --
    // layout of one warp's block: [j][TD] with d = lane + 32 i contiguous across lanes
    for (int half = 4; half >= 1; half >>= 1) {
        __syncthreads();
        if (warp >= half && warp < 2 * half) {
            float* dst = s_red + (warp - half) * TD * JP;
#pragma unroll
            for (int i = 0; i < DT; ++i)
#pragma unroll
                for (int j = 0; j < JP / 2; ++j) {
                    dst[(2 * j) * TD + lane + 32 * i] = acc[i][j].x;
                    dst[(2 * j + 1) * TD + lane + 32 * i] = acc[i][j].y;
                }
        }
        __syncthreads();
        if (warp < half) {
            const float* src = s_red + warp * TD * JP;
#pragma unroll
            for (int i = 0; i < DT; ++i)
#pragma unroll
                for (int j = 0; j < JP / 2; ++j) {
                    acc[i][j].x += src[(2 * j) * TD + lane + 32 * i];
                    acc[i][j].y += src[(2 * j + 1) * TD + lane + 32 * i];
                }
        }
    }
    if (warp == 0) {
#pragma unroll
        for (int j = 0; j < JP / 2; ++j) {
#pragma unroll
            for (int i = 0; i < DT; ++i) {
                const int d = d0 + lane + 32 * i;
                if (d < D) {
                    if (part_sum) {
                        float* ps = part_sum + ((int64_t)b * gridDim.z + blockIdx.z) * J * D;
                        if (2 * j < J) ps[(int64_t)(2 * j) * D + d] = acc[i][j].x;
                        if (2 * j + 1 < J) ps[(int64_t)(2 * j + 1) * D + d] = acc[i][j].y;
                    } else {
                        if (2 * j < J) mu_out[((int64_t)b * J + 2 * j) * D + d] = __fdiv_rn(acc[i][j].x, s_npi[2 * j]);
                        if (2 * j + 1 < J) mu_out[((int64_t)b * J + 2 * j + 1) * D + d] = __fdiv_rn(acc[i][j].y, s_npi[2 * j + 1]);
                    }
                }
            }
        }
    }
}

// =====================================================================================================
// Fold of the split mode: adds the splits' raw sums in split order (deterministic) and normalises like the one-pass kernel.
__global__ void __launch_bounds__(256)
gmm_moments_feat_fold_kernel(const float* __restrict__ part_sum, const float* __restrict__ part_gs, int S, int N, int J, int D,
                             float* __restrict__ pi_out, float* __restrict__ mu_out) {
    const int b = blockIdx.y;
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // (j, d)
    if (e >= (int64_t)J * D) return;
    const int j = (int)(e / D);
    float t = 0.f, a = 0.f;
    for (int s = 0; s < S; ++s) {
        t += part_gs[((int64_t)b * S + s) * J + j];
        a += part_sum[((int64_t)b * S + s) * J * D + e];
    }
    const float pi = __fdiv_rn(t, (float)N);
    const float npi = __fadd_rn(__fmul_rn(pi, (float)N), 1e-5f);
    mu_out[(int64_t)b * J * D + e] = __fdiv_rn(a, npi);
    if (e - (int64_t)j * D == 0 && pi_out) pi_out[(int64_t)b * J + j] = pi;
}

// =====================================================================================================
// DeepGMR: gamma = softmax_j(logits[b,:,n]); pi, mu, sigma as above.  One CTA per cloud.  Thread per
// point; the softmax row is recomputed in the sigma pass rather than stored (logits stay L2-resident).
// =====================================================================================================
constexpr int kSoftThreads = 256;

__global__ void __launch_bounds__(kSoftThreads)
softmax_moments_kernel(const float* __restrict__ logits, const float* __restrict__ pts, int64_t p_sb, int64_t p_sn,
                       int64_t p_sd, int N, int J, float* __restrict__ gamma_out, float* __restrict__ pi_out,
                       float* __restrict__ mu_out, float* __restrict__ sigma_out) {
    extern __shared__ __align__(16) float sm[];
    constexpr int NW = kSoftThreads / 32;
    float* s_part = sm;                              // [NW][J][4]
    float* s_mu = s_part + (size_t)NW * J * 4;       // [J][4]: mu xyz, npi
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* lg = logits + (int64_t)b * J * N;
    const float* x = pts + (int64_t)b * p_sb;

    for (int pass = 0; pass < 2; ++pass) {
        for (int e = tid; e < NW * J * 4; e += kSoftThreads) s_part[e] = 0.f;
        __syncthreads();
        for (int n0 = 0; n0 < N; n0 += kSoftThreads) {
            const int n = n0 + tid;
            const bool ok = n < N;
            float mx = -INFINITY, den = 0.f, px = 0.f, py = 0.f, pz = 0.f;
            if (ok) {
                px = x[(int64_t)n * p_sn]; py = x[(int64_t)n * p_sn + p_sd]; pz = x[(int64_t)n * p_sn + 2 * p_sd];
                for (int j = 0; j < J; ++j) mx = fmaxf(mx, lg[(int64_t)j * N + n]);
                for (int j = 0; j < J; ++j) den += expf(lg[(int64_t)j * N + n] - mx);
            }
            for (int j = 0; j < J; ++j) {
                float gv = 0.f;
                if (ok) {
                    gv = __fdiv_rn(expf(lg[(int64_t)j * N + n] - mx), den);
                    if (pass == 0 && gamma_out) gamma_out[((int64_t)b * J + j) * N + n] = gv;
                }
                float a0, a1 = 0.f, a2 = 0.f, a3 = 0.f;
                if (pass == 0) { a0 = gv; a1 = gv * px; a2 = gv * py; a3 = gv * pz; }
                else {
                    const float dx = px - s_mu[4 * j], dy = py - s_mu[4 * j + 1], dz = pz - s_mu[4 * j + 2];
                    a0 = fmaf(dz, dz, fmaf(dy, dy, dx * dx)) * gv;
                }
                a0 = warp_sum(a0);
                if (pass == 0) { a1 = warp_sum(a1); a2 = warp_sum(a2); a3 = warp_sum(a3); }
                if (lane == 0) {
                    float* p = s_part + ((size_t)warp * J + j) * 4;
                    p[0] += a0; p[1] += a1; p[2] += a2; p[3] += a3;
                }
            }
        }
        __syncthreads();
        for (int j = tid; j < J; j += kSoftThreads) {
            float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
            for (int w = 0; w < NW; ++w) {
                const float* p = s_part + ((size_t)w * J + j) * 4;
                t0 += p[0]; t1 += p[1]; t2 += p[2]; t3 += p[3];
            }
            if (pass == 0) {
                const float pi = __fdiv_rn(t0, (float)N);
                const float npi = __fadd_rn(__fmul_rn(pi, (float)N), 1e-5f);
                pi_out[(int64_t)b * J + j] = pi;
                const float m0 = __fdiv_rn(t1, npi), m1 = __fdiv_rn(t2, npi), m2 = __fdiv_rn(t3, npi);
                s_mu[4 * j] = m0; s_mu[4 * j + 1] = m1; s_mu[4 * j + 2] = m2; s_mu[4 * j + 3] = npi;
                float* m = mu_out + ((int64_t)b * J + j) * 3;
                m[0] = m0; m[1] = m1; m[2] = m2;
            } else if (sigma_out) {
                const float sg = __fdiv_rn(t0, s_mu[4 * j + 3]);
                float* so = sigma_out + ((int64_t)b * J + j) * 9;
                so[0] = sg; so[1] = 0.f; so[2] = 0.f; so[3] = 0.f; so[4] = sg; so[5] = 0.f; so[6] = 0.f; so[7] = 0.f; so[8] = sg;
            }
        }
        __syncthreads();
        if (!sigma_out) break;
    }
}

// J <= 16 variant (DeepGMR's J = 16): thread per point with all J softmax terms in registers, the four moment sums
// of a cluster accumulated per thread over its points and folded ONCE per pass with the 16-column butterfly (64 + 16
// shuffles per thread instead of 80 per point and cluster), one reciprocal per point, MUFU exponentials.  The generic
// kernel above stays for J > 16.
__global__ void __launch_bounds__(kSoftThreads, 2)
softmax_moments16_kernel(const float* __restrict__ logits, const float* __restrict__ pts, int64_t p_sb, int64_t p_sn,
                         int64_t p_sd, int N, int J, float* __restrict__ gamma_out, float* __restrict__ pi_out,
                         float* __restrict__ mu_out, float* __restrict__ sigma_out) {
    constexpr int NW = kSoftThreads / 32;
    __shared__ float s_w[NW][kJC][4];
    __shared__ float s_mu[kJC][4];                   // mu xyz, npi
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* lg = logits + (int64_t)b * J * N;
    const float* x = pts + (int64_t)b * p_sb;
    const int col = (lane >> 1) & 15;

    // softmax row of point n in registers: e[j] = exp(l_j - max) / sum, zero for j >= J
    auto softmax_row = [&](int n, float (&e)[kJC]) {
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < kJC; ++j) { e[j] = j < J ? __ldg(lg + (int64_t)j * N + n) : -INFINITY; mx = fmaxf(mx, e[j]); }
        float den = 0.f;
#pragma unroll
        for (int j = 0; j < kJC; ++j) { e[j] = __expf(e[j] - mx); den += e[j]; }      // exp(-inf) = 0 for the padding
        const float inv = __fdiv_rn(1.0f, den);
#pragma unroll
        for (int j = 0; j < kJC; ++j) e[j] *= inv;
    };

    float a0[kJC], ax[kJC], ay[kJC], az[kJC];
#pragma unroll
    for (int j = 0; j < kJC; ++j) { a0[j] = 0.f; ax[j] = 0.f; ay[j] = 0.f; az[j] = 0.f; }
    for (int n = tid; n < N; n += kSoftThreads) {
        const float px = x[(int64_t)n * p_sn], py = x[(int64_t)n * p_sn + p_sd], pz = x[(int64_t)n * p_sn + 2 * p_sd];
        float e[kJC];
        softmax_row(n, e);
#pragma unroll
        for (int j = 0; j < kJC; ++j) {
            if (gamma_out && j < J) gamma_out[((int64_t)b * J + j) * N + n] = e[j];
            a0[j] += e[j];
            ax[j] = fmaf(e[j], px, ax[j]); ay[j] = fmaf(e[j], py, ay[j]); az[j] = fmaf(e[j], pz, az[j]);
        }
    }
    {
        const float t0 = butterfly16(a0, lane), tx = butterfly16(ax, lane), ty = butterfly16(ay, lane), tz = butterfly16(az, lane);
        if ((lane & 1) == 0) { s_w[warp][col][0] = t0; s_w[warp][col][1] = tx; s_w[warp][col][2] = ty; s_w[warp][col][3] = tz; }
    }
    __syncthreads();
    if (tid < J) {
        float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) { t0 += s_w[w][tid][0]; t1 += s_w[w][tid][1]; t2 += s_w[w][tid][2]; t3 += s_w[w][tid][3]; }
        const float pi = __fdiv_rn(t0, (float)N);
        const float npi = __fadd_rn(__fmul_rn(pi, (float)N), 1e-5f);
        pi_out[(int64_t)b * J + tid] = pi;
        const float m0 = __fdiv_rn(t1, npi), m1 = __fdiv_rn(t2, npi), m2 = __fdiv_rn(t3, npi);
        s_mu[tid][0] = m0; s_mu[tid][1] = m1; s_mu[tid][2] = m2; s_mu[tid][3] = npi;
        float* m = mu_out + ((int64_t)b * J + tid) * 3;
        m[0] = m0; m[1] = m1; m[2] = m2;
    }
    if (!sigma_out) return;
    __syncthreads();
    // ---- sigma_j = sum_n gamma_nj |x_n - mu_j|^2 / npi_j: the softmax rows are recomputed (logits are L2-resident)
    float s2[kJC];
#pragma unroll
    for (int j = 0; j < kJC; ++j) s2[j] = 0.f;
    for (int n = tid; n < N; n += kSoftThreads) {
        const float px = x[(int64_t)n * p_sn], py = x[(int64_t)n * p_sn + p_sd], pz = x[(int64_t)n * p_sn + 2 * p_sd];
        float e[kJC];
        softmax_row(n, e);
#pragma unroll
        for (int j = 0; j < kJC; ++j) {
            if (j < J) {
                const float dx = px - s_mu[j][0], dy = py - s_mu[j][1], dz = pz - s_mu[j][2];
                s2[j] = fmaf(fmaf(dz, dz, fmaf(dy, dy, dx * dx)), e[j], s2[j]);
            }
        }
    }
    {
        const float t = butterfly16(s2, lane);
        if ((lane & 1) == 0) s_w[warp][col][0] = t;
    }
    __syncthreads();
    if (tid < J) {
        float t0 = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) t0 += s_w[w][tid][0];
        const float sg = __fdiv_rn(t0, s_mu[tid][3]);
        float* so = sigma_out + ((int64_t)b * J + tid) * 9;
        so[0] = sg; so[1] = 0.f; so[2] = 0.f; so[3] = 0.f; so[4] = sg; so[5] = 0.f; so[6] = 0.f; so[7] = 0.f; so[8] = sg;
    }
}

}  // namespace ogmm

using namespace ogmm;

int ogmm_launch_moments_feat_tma(const float* gamma, int64_t g_sb, int64_t g_sn, int64_t g_sj,
                                 const float* feats, int64_t f_sb, int64_t f_sn, int64_t f_sd,
                                 int64_t B, int64_t N, int64_t J, int64_t D, float* pi_out, float* mu_out, cudaStream_t s);
int ogmm_launch_moments_feat_tc(const float* gamma, int64_t g_sb, int64_t g_sn, int64_t g_sj,
                                const float* feats, int64_t f_sb, int64_t f_sn, int64_t f_sd,
                                int64_t B, int64_t N, int64_t J, int64_t D, float* pi_out, float* mu_out, cudaStream_t s);

// FP32 FFMA2 kernel for every shape; splits > 1: split mode + fold (see feat_splits below).
static int launch_feat_generic(const float* gamma, int64_t g_sb, int64_t g_sn, int64_t g_sj,
                               const float* feats, int64_t f_sb, int64_t f_sn, int64_t f_sd,
                               int64_t B, int64_t N, int64_t J, int64_t D, float* pi_out, float* mu_out,
                               int splits, float* part_sum, float* part_gs, cudaStream_t s) {
    // points per split: a multiple of the stage size, so only the last split of a cloud sees a ragged tile
    const int n_per = splits > 1 ? (int)(((N + splits - 1) / splits + kTN - 1) / kTN * kTN) : (int)N;
#define LAUNCH(JP, DT)                                                                                              \
    do {                                                                                                            \
        using C = FeatCfg<JP, DT>;                                                                                  \
        if (C::SMEM > 48 * 1024) {                                                                                  \
            int st = cuda_status(cudaFuncSetAttribute(gmm_moments_feat_kernel<JP, DT>,                              \
                                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM),   \
                                 "cudaFuncSetAttribute(gmm_moments_feat_kernel)");                                  \
            if (st != OGMM_OK) return st;                                                                           \
        }                                                                                                           \
        dim3 grid((unsigned)((D + C::TD - 1) / C::TD), (unsigned)B, (unsigned)splits);                              \
        gmm_moments_feat_kernel<JP, DT><<<grid, kFeatThreads, C::SMEM, s>>>(gamma, g_sb, g_sn, g_sj, feats, f_sb,   \
                                                                            f_sn, f_sd, (int)N, (int)J, (int)D,     \
                                                                            pi_out, mu_out, n_per,                  \
                                                                            splits > 1 ? part_sum : nullptr,        \
                                                                            part_gs);                               \
    } while (0)
    if (J <= 16) LAUNCH(16, 4);
    else if (J <= 32) LAUNCH(32, 2);
    else if (J <= 64) LAUNCH(64, 1);
    else LAUNCH(128, 1);
#undef LAUNCH
    OGMM_LAUNCH_CHECK("gmm_moments_feat_kernel");
    if (splits > 1) {
        dim3 grid((unsigned)((J * D + 255) / 256), (unsigned)B);
        gmm_moments_feat_fold_kernel<<<grid, 256, 0, s>>>(part_sum, part_gs, splits, (int)N, (int)J, (int)D, pi_out, mu_out);
        OGMM_LAUNCH_CHECK("gmm_moments_feat_fold_kernel");
    }
    return OGMM_OK;
}

extern "C" __attribute__((visibility("default"))) int ogmm_gmm_moments_feat(const float* gamma, int64_t g_sb, int64_t g_sn, int64_t g_sj,
                                     const float* feats, int64_t f_sb, int64_t f_sn, int64_t f_sd,
                                     int64_t B, int64_t N, int64_t J, int64_t D,
                                     float* pi_out, float* mu_out, ogmm_stream_t stream) {
    OGMM_REQUIRE(B >= 0 && N >= 1 && J >= 1 && D >= 1 && B < 65536 && N < (1ll << 31), OGMM_EINVAL,
                 "ogmm_gmm_moments_feat: bad sizes B=%lld N=%lld J=%lld D=%lld", (long long)B, (long long)N,
                 (long long)J, (long long)D);
    OGMM_REQUIRE(J <= 128, OGMM_EUNSUPPORTED, "ogmm_gmm_moments_feat: J=%lld > 128", (long long)J);
    if (B == 0) return OGMM_OK;
    OGMM_REQUIRE(gamma && feats && mu_out, OGMM_EINVAL, "ogmm_gmm_moments_feat: null pointer");
    cudaStream_t s = as_stream(stream);
    // J == 16, native (B,D,N) layout: TMA -> shared -> mma.sync 3xTF32 pipeline (moments_tma.cu); OGMM_FEAT_NO_TMA=1
    // drops to the FP32 FFMA2 kernel below, which also takes every other shape.
    // OGMM_FEAT_TENSOR=1 (checked first when set) opts into the tcgen05 3xTF32 kernel (moments_tc.cu; J == 16, native (B,D,N) layout).  It is
    // correct to FP32 accuracy but measured slower than the FP32 kernels on B200, so it is not
    // the default.
    {
        const char* off = getenv("OGMM_FEAT_NO_TMA");
        if (!(off && off[0] == '1')) {
            const int st = ogmm_launch_moments_feat_tma(gamma, g_sb, g_sn, g_sj, feats, f_sb, f_sn, f_sd, B, N, J, D, pi_out,
                                                        mu_out, s);
            if (st != OGMM_EUNSUPPORTED) return st;
        }
    }
    {
        const char* opt = getenv("OGMM_FEAT_TENSOR");
        if (opt && opt[0] == '1') {
            const int st = ogmm_launch_moments_feat_tc(gamma, g_sb, g_sn, g_sj, feats, f_sb, f_sn, f_sd, B, N, J, D, pi_out,
                                                       mu_out, s);
            if (st != OGMM_EUNSUPPORTED) return st;
        }
    }
    return launch_feat_generic(gamma, g_sb, g_sn, g_sj, feats, f_sb, f_sn, f_sd, B, N, J, D, pi_out, mu_out, 1, nullptr, nullptr, s);
}

// Split mode for few clouds with very many points (BASELINE.json configs[3]: 4 clouds x 16384 points give 64 CTAs of
// cloud x row-block items on 148 SMs): the points of a cloud are divided over `splits` CTAs per row block, raw sums go
// to the workspace and a fold kernel adds them in split order.  ogmm_gmm_moments_feat_workspace returns 0 when the
// one-pass kernels already fill the GPU.
static int64_t feat_splits(int64_t B, int64_t N, int64_t J, int64_t D) {
    const int64_t td = J <= 16 ? 128 : (J <= 32 ? 64 : 32);
    const int64_t ctas = B * ((D + td - 1) / td);
    if (J <= 16 || N < 4096 || ctas >= 2 * 148) return 1;
    int64_t sp = (2 * 148 + ctas - 1) / ctas;
    if (sp > N / 2048) sp = N / 2048;
    return sp < 1 ? 1 : sp;
}

extern "C" __attribute__((visibility("default"))) int64_t ogmm_gmm_moments_feat_workspace(int64_t B, int64_t N, int64_t J, int64_t D) {
    if (B < 1 || N < 1 || J < 1 || D < 1) return 0;
    const int64_t sp = feat_splits(B, N, J, D);
    return sp <= 1 ? 0 : 4 * B * sp * (J * D + J);
}

extern "C" __attribute__((visibility("default"))) int ogmm_gmm_moments_feat_ws(const float* gamma, int64_t g_sb, int64_t g_sn, int64_t g_sj,
                                                                             const float* feats, int64_t f_sb, int64_t f_sn, int64_t f_sd,
                                                                             int64_t B, int64_t N, int64_t J, int64_t D,
                                                                             float* pi_out, float* mu_out, void* workspace,
                                                                             int64_t workspace_bytes, ogmm_stream_t stream) {
    const int64_t need = ogmm_gmm_moments_feat_workspace(B, N, J, D);
    if (need == 0 || workspace == nullptr)
        return ogmm_gmm_moments_feat(gamma, g_sb, g_sn, g_sj, feats, f_sb, f_sn, f_sd, B, N, J, D, pi_out, mu_out, stream);
    OGMM_REQUIRE(B < 65536 && J <= 128, OGMM_EUNSUPPORTED, "ogmm_gmm_moments_feat_ws: B=%lld J=%lld", (long long)B, (long long)J);
    OGMM_REQUIRE(workspace_bytes >= need, OGMM_EWORKSPACE, "ogmm_gmm_moments_feat_ws: workspace of %lld B given, %lld B needed",
                 (long long)workspace_bytes, (long long)need);
    OGMM_REQUIRE(gamma && feats && mu_out, OGMM_EINVAL, "ogmm_gmm_moments_feat_ws: null pointer");
    const int64_t sp = feat_splits(B, N, J, D);
    float* part_sum = static_cast<float*>(workspace);
    float* part_gs = part_sum + B * sp * J * D;
    return launch_feat_generic(gamma, g_sb, g_sn, g_sj, feats, f_sb, f_sn, f_sd, B, N, J, D, pi_out, mu_out, (int)sp, part_sum, part_gs,
                               as_stream(stream));
}

extern "C" __attribute__((visibility("default"))) int ogmm_gmm_moments(const float* gamma, int64_t g_sb, int64_t g_sn, int64_t g_sj,
                                const float* pts, int64_t p_sb, int64_t p_sn, int64_t p_sd,
                                int64_t B, int64_t N, int64_t J, int64_t D,
                                float* pi_out, float* mu_out, float* sigma_out, ogmm_stream_t stream) {
    OGMM_REQUIRE(B >= 0 && N >= 1 && J >= 1 && D >= 1 && B < (1ll << 31) && N < (1ll << 31), OGMM_EINVAL,
                 "ogmm_gmm_moments: bad sizes B=%lld N=%lld J=%lld D=%lld", (long long)B, (long long)N, (long long)J,
                 (long long)D);
    if (B == 0) return OGMM_OK;
    OGMM_REQUIRE(gamma && pts && pi_out && mu_out, OGMM_EINVAL, "ogmm_gmm_moments: null pointer");
    if (D > kSmallD) {
        OGMM_REQUIRE(sigma_out == nullptr, OGMM_EUNSUPPORTED, "ogmm_gmm_moments: sigma is built for D <= %d, got D=%lld",
                     kSmallD, (long long)D);
        return ogmm_gmm_moments_feat(gamma, g_sb, g_sn, g_sj, pts, p_sb, p_sn, p_sd, B, N, J, D, pi_out, mu_out, stream);
    }
    OGMM_REQUIRE(J <= 8192, OGMM_EUNSUPPORTED, "ogmm_gmm_moments: J=%lld > 8192", (long long)J);
    const size_t smem = sizeof(float) * ((size_t)(kSmallThreads / 32) * 128 * (kSmallD + 1) + (size_t)J * kSmallD + J);
    if (smem > 48 * 1024) {
        int st = cuda_status(cudaFuncSetAttribute(gmm_moments_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                  (int)smem), "cudaFuncSetAttribute(gmm_moments_small_kernel)");
        if (st != OGMM_OK) return st;
    }
    gmm_moments_small_kernel<<<(unsigned)B, kSmallThreads, smem, as_stream(stream)>>>(
        gamma, g_sb, g_sn, g_sj, pts, p_sb, p_sn, p_sd, (int)N, (int)J, (int)D, pi_out, mu_out, sigma_out);
    OGMM_LAUNCH_CHECK("gmm_moments_small_kernel");
    return OGMM_OK;
}

extern "C" __attribute__((visibility("default"))) int ogmm_softmax_moments(const float* logits, const float* pts, int64_t p_sb, int64_t p_sn, int64_t p_sd,
                                    int64_t B, int64_t N, int64_t J, float* gamma_out, float* pi_out, float* mu_out,
                                    float* sigma_out, ogmm_stream_t stream) {
    OGMM_REQUIRE(B >= 0 && N >= 1 && J >= 1 && B < (1ll << 31) && N < (1ll << 31), OGMM_EINVAL,
                 "ogmm_softmax_moments: bad sizes");
    OGMM_REQUIRE(J <= 1024, OGMM_EUNSUPPORTED, "ogmm_softmax_moments: J=%lld > 1024", (long long)J);
    if (B == 0) return OGMM_OK;
    OGMM_REQUIRE(logits && pts && pi_out && mu_out, OGMM_EINVAL, "ogmm_softmax_moments: null pointer");
    if (J <= kJC) {
        softmax_moments16_kernel<<<(unsigned)B, kSoftThreads, 0, as_stream(stream)>>>(
            logits, pts, p_sb, p_sn, p_sd, (int)N, (int)J, gamma_out, pi_out, mu_out, sigma_out);
        OGMM_LAUNCH_CHECK("softmax_moments16_kernel");
        return OGMM_OK;
    }
    const size_t smem = sizeof(float) * ((size_t)(kSoftThreads / 32) * J * 4 + (size_t)J * 4);
    if (smem > 48 * 1024) {
        int st = cuda_status(cudaFuncSetAttribute(softmax_moments_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                  (int)smem), "cudaFuncSetAttribute(softmax_moments_kernel)");
        if (st != OGMM_OK) return st;
    }
    softmax_moments_kernel<<<(unsigned)B, kSoftThreads, smem, as_stream(stream)>>>(
        logits, pts, p_sb, p_sn, p_sd, (int)N, (int)J, gamma_out, pi_out, mu_out, sigma_out);
    OGMM_LAUNCH_CHECK("softmax_moments_kernel");
    return OGMM_OK;
}
