// K2: farthest-point init + overlap-guided Sinkhorn k-means, and the stand-alone Sinkhorn (sm_100a).
//
// One CTA per cloud; every point lives in the registers of one thread for the whole 10 x 10
// iteration loop (lib/utils.py:269-288), so the only HBM traffic is the algorithmic one: xyz +
// overlap scores in, gamma / pi / mu out.
//
// Reference arithmetic reproduced here
//   FPS init             lib/utils.py:170-198 (is_center branch :183-188)
//   cost                 lib/utils.py:280  torch.cdist -> ATen _euclidean_dist (matmul form, K=5),
//                        clamp_min(0).sqrt(), .clip(0) / tau
//   Sinkhorn             lib/utils.py:69-108; row normalisation over j, column normalisation over i
//   early exit           lib/utils.py:99-102 BATCH mean of sum|du|+sum|dv| < thresh (see below)
//   post                 lib/utils.py:282 nan_to_num, :287 gamma / clip(sum_j gamma, 1e-3)
//   M-step (xyz)         lib/utils.py:130-140
//
// Two arithmetic paths for the inner Sinkhorn iterations:
//
//  * scaled domain (kFast; J <= 16, N <= 1024).  For one Sinkhorn call the thread keeps
//    G_ij = exp(-(c_ij - m_i)/eps), m_i = min_j c_ij, for its points in REGISTERS and iterates on the
//    scalings  r_i = sum_j G_ij b_j, a_i = (p_i+1e-8)/r_i, s_j = sum_i G_ij a_i, b_j = (q_j+1e-8)/s_j
//    with FMAs only: one exp per (i,j) per call instead of two per iteration.  In exact arithmetic this
//    IS the reference's log-domain update (u_i = m_i + eps log a_i, v_j = eps log b_j); in fp32 it agrees
//    with the fp32 reference to within the reference's own fp32-vs-fp64 spread
//    (tools/sinkhorn_scaled_emulation.py).  A monitor (every s_j in [1e-30, 1e30], max b / min b <= 1e24)
//    guards the dynamic range; if it ever trips, that Sinkhorn call is redone in the log domain.
//  * log domain (generic J, N <= 8192; also the rescue path).  exp2/log2 with a max shift on rows;
//    column sums need no shift after a row update (every K_ij <= log(p_i+1e-8) < 0) and a column whose
//    sum underflows is redone with an exact shift.  Costs are recomputed from registers + shared memory.
//
// Batch-coupled early exit without a host sync.  Iteration counts couple the clouds of one call only
// through the exit test.  The main launch runs every Sinkhorn call for max_iter iterations and records
// each cloud's change per (outer, inner) iteration; the last CTA to finish evaluates the batch means in
// order and, at the first (outer o, inner i) with mean < thresh and i+1 below the count that was run,
// stores n_inner[o] = i+1 and resume = o.  Only then does that CTA queue a REDO round itself: a device-side tail
// launch (CUDA dynamic parallelism, cudaStreamTailLaunch) of the same kernel, which starts after the parent grid has
// drained and before anything else in the stream, re-runs every cloud from the centroids saved at the start of outer
// iteration `resume`, re-evaluates the means in its own last CTA and, if the schedule still moves, queues the next
// round.  A redo round certifies at least one more outer iteration (the re-run of iteration `resume` repeats the
// first n_inner[resume] recorded iterations exactly), so at most `iters` rounds run.  In the common case -- no exit
// fires -- the host's single launch is all there is.  No launch ever waits on another CTA: there is no grid barrier
// and no co-residency requirement, so any number of calls may be in flight on a device (streams, MPS partitions,
// green contexts), and the chain is captured by CUDA graphs like any kernel.  The result is bit-identical to running
// the exit test inline, and deterministic (fixed-order reductions, no float atomics).
#pragma once

#include <stdlib.h>

#include "common.cuh"

namespace ogmm {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

struct ClusterWsLayout {
    int64_t state_off, ninner_off, means_off, diffs_off, hist_off, total;
};
__host__ __device__ inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }
__host__ __device__ inline ClusterWsLayout cluster_ws_layout(int64_t B, int64_t J, int64_t iters, int64_t max_iter) {
    ClusterWsLayout l;
    l.state_off = 0;                                         // int32[16]: [0]=resume [1]=done counter [2]=rescues [5]=device-side launch failed
    l.ninner_off = 64;                                       // int32[iters]
    l.means_off = align_up(l.ninner_off + 4 * iters, 256);   // float[iters][max_iter] batch means
    l.diffs_off = align_up(l.means_off + 4 * iters * max_iter, 256);   // float[iters][max_iter][B]
    l.hist_off = align_up(l.diffs_off + 4 * iters * max_iter * B, 256);   // float[B][iters][J][3]
    l.total = align_up(l.hist_off + 4 * B * iters * J * 3, 256);
    return l;
}

__device__ __forceinline__ float fast_exp2(float x) {         // MUFU.EX2, flush-to-zero: no denormal fix-up code
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_rcp(float x) {          // MUFU.RCP, flush-to-zero
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_lg2(float x) {          // MUFU.LG2, flush-to-zero
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sq3(float x, float y, float z) {
    return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
}
__device__ __forceinline__ unsigned long long far_key(float best, int i) {
    return ((unsigned long long)__float_as_uint(best) << 32) | (0xffffffffu - (unsigned)i);
}

// ---- farthest point sampling on register-resident points ----------------------------------------
// One relaxation against centre (cx,cy,cz) followed by the block-wide arg-max (lowest index on ties).
template <int NT, int PPT>
__device__ __forceinline__ int fps_relax(const float (&px)[PPT], const float (&py)[PPT], const float (&pz)[PPT],
                                         float (&best)[PPT], int N, float cx, float cy, float cz,
                                         unsigned long long* s_key) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = NT / 32;
    unsigned long long key = 0ull;
#pragma unroll
    for (int p = 0; p < PPT; ++p) {
        const int i = tid + p * NT;
        if (i < N) {
            const float d = sq3(px[p] - cx, py[p] - cy, pz[p] - cz);      // sum((xyz - c) ** 2, -1): no FMA
            if (d < best[p]) best[p] = d;
            const unsigned long long k = far_key(best[p], i);
            key = k > key ? k : key;
        }
    }
    key = warp_max_u64(key);
    __syncthreads();                       // previous readers of s_key are done
    if (lane == 0) s_key[warp] = key;
    __syncthreads();
    unsigned long long k2 = lane < NW ? s_key[lane] : 0ull;
    k2 = warp_max_u64(k2);
    return (int)(0xffffffffu - (unsigned)(k2 & 0xffffffffull));
}

// lib/utils.py:191-197.  `far` is the first index; `emit(s, far)` is called for each of the npoint picks.
template <int NT, int PPT, typename Emit>
__device__ __forceinline__ void fps_run(const float (&px)[PPT], const float (&py)[PPT], const float (&pz)[PPT],
                                        float (&best)[PPT], int N, int npoint, int far, float* s_pick,
                                        unsigned long long* s_key, Emit emit) {
    const int tid = threadIdx.x;
    for (int s = 0; s < npoint; ++s) {
        emit(s, far);
        __syncthreads();                   // previous readers of s_pick are done
        if ((far % NT) == tid) {
            const int slot = far / NT;
#pragma unroll
            for (int p = 0; p < PPT; ++p)
                if (p == slot) { s_pick[0] = px[p]; s_pick[1] = py[p]; s_pick[2] = pz[p]; }
        }
        __syncthreads();
        far = fps_relax<NT, PPT>(px, py, pz, best, N, s_pick[0], s_pick[1], s_pick[2], s_key);
    }
}

// First index for is_center=True (:183-188): relax against the centroid, take the farthest point.
template <int NT, int PPT>
__device__ __forceinline__ int fps_center_start(const float (&px)[PPT], const float (&py)[PPT], const float (&pz)[PPT],
                                                float (&best)[PPT], int N, float* s_red, unsigned long long* s_key) {
    float sx = 0.f, sy = 0.f, sz = 0.f;
#pragma unroll
    for (int p = 0; p < PPT; ++p)
        if ((int)threadIdx.x + p * NT < N) { sx += px[p]; sy += py[p]; sz += pz[p]; }
    const float cx = block_sum<NT>(sx, s_red) / (float)N;
    const float cy = block_sum<NT>(sy, s_red) / (float)N;
    const float cz = block_sum<NT>(sz, s_red) / (float)N;
    return fps_relax<NT, PPT>(px, py, pz, best, N, cx, cy, cz, s_key);
}

template <int NT, int PPT>
__global__ void __launch_bounds__(NT)
fps_kernel(const float* __restrict__ xyz, int64_t sb, int64_t sn, int64_t sc, int N, int npoint,
           const int64_t* __restrict__ start, int64_t* __restrict__ ids_out, float* __restrict__ pts_out) {
    __shared__ float s_pick[4];
    __shared__ unsigned long long s_key[32];
    __shared__ float s_red[32];
    const int b = blockIdx.x, tid = threadIdx.x;
    const float* base = xyz + (int64_t)b * sb;
    float px[PPT], py[PPT], pz[PPT], best[PPT];
#pragma unroll
    for (int p = 0; p < PPT; ++p) {
        const int i = tid + p * NT;
        px[p] = py[p] = pz[p] = 0.f;
        best[p] = 1e10f;
        if (i < N) { px[p] = base[(int64_t)i * sn]; py[p] = base[(int64_t)i * sn + sc]; pz[p] = base[(int64_t)i * sn + 2 * sc]; }
    }
    int far;
    if (start == nullptr) far = fps_center_start<NT, PPT>(px, py, pz, best, N, s_red, s_key);
    else {
        far = (int)start[b];
        far = far < 0 ? 0 : (far >= N ? N - 1 : far);
    }
    int64_t* ids = ids_out + (int64_t)b * npoint;
    float* pts = pts_out ? pts_out + (int64_t)b * npoint * 3 : nullptr;
    fps_run<NT, PPT>(px, py, pz, best, N, npoint, far, s_pick, s_key, [&](int s, int f) {
        if (tid == 0) {
            ids[s] = f;
            if (pts) {
                pts[3 * s] = base[(int64_t)f * sn]; pts[3 * s + 1] = base[(int64_t)f * sn + sc];
                pts[3 * s + 2] = base[(int64_t)f * sn + 2 * sc];
            }
        }
    });
}

// ---- cost ------------------------------------------------------------------------------------------------
// c_ij = cdist(x_i, node_j).clip(0) / tau.  ATen _euclidean_dist: [-2x, |x|^2, 1] . [y, 1, |y|^2] in k order.
// cdist alone (tau == 1, how Clustering.forward calls wkeans_plus, models/gmmreg.py:28): callers that build a whole row
// test tau once and pick this version, instead of one compare + select per entry.
__device__ __forceinline__ float node_dist(float m2x, float m2y, float m2z, float pn, const float4 c) {
    float d2 = __fmul_rn(m2x, c.x);
    d2 = fmaf(m2y, c.y, d2);
    d2 = fmaf(m2z, c.z, d2);
    d2 = __fadd_rn(d2, pn);
    d2 = __fadd_rn(d2, c.w);
    float d;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(fmaxf(d2, 0.f)));  // 1 ulp; clamp_min(0).sqrt()
    return d;
}
__device__ __forceinline__ float node_cost(float m2x, float m2y, float m2z, float pn, const float4 c, float tau) {
    const float d = node_dist(m2x, m2y, m2z, pn, c);
    return tau == 1.0f ? d : __fdiv_rn(d, tau);
}

struct SinkhornParams {
    const float* xyz; int64_t sb, sn, sc;      // cluster mode: (B,N,3) view
    const float* o_scores;                     // cluster mode (B,N)
    const float* cost;                         // matrix mode (B,N,J)
    const float* p;                            // matrix mode (B,N) or null
    const float* q;                            // matrix mode (B,J) or null
    int B, N, J, iters, max_iter;
    float tau, eps, thresh;
    float* gamma; float* pi; float* mu; float* loss; int32_t* iters_run;
    int32_t* state; int32_t* n_inner; float* means; float* diffs; float* hist;
};

// Shared memory of one CTA.
struct Smem {
    float4* node;             // [Jp]  centroid (x, y, z, |n|^2)
    float* v;                 // [Jp]  column potential (log domain) / eps log b (fast path)
    float* logq;              // [Jp]
    float* bq;                // [Jp]  fast path: b_j
    float* wtot;              // [4][NW][Jp] per-warp column totals
    float* red;               // [32]  block_sum scratch
    float* du;                // [32]  per-warp sum |du|
    float* tmp;               // [32]
    float* misc;              // [16]  [0..2] FPS pick, [5] last-CTA flag, [6] monitor flag
    unsigned long long* key;  // [32]
    float4* pts;              // fast path: [NT * PPT] staged points (-2x, -2y, -2z, |x|^2), one LDS.128 per reload
};
template <int NT>
__host__ __device__ inline size_t sinkhorn_smem(int J, int staged_points = 0) {
    const int Jp = (J + kJC - 1) / kJC * kJC;
    const size_t small = sizeof(float) * ((size_t)4 * Jp + 3 * (size_t)Jp + (size_t)(NT / 32) * Jp * 4 + 32 + 32 + 32 + 16) +
                         sizeof(unsigned long long) * 32;
    return ((small + 15) & ~(size_t)15) + sizeof(float4) * (size_t)staged_points;
}
template <int NT>
__device__ __forceinline__ Smem carve_smem(unsigned char* raw, int Jp) {
    Smem s;
    s.node = reinterpret_cast<float4*>(raw);
    s.v = reinterpret_cast<float*>(s.node + Jp);
    s.logq = s.v + Jp;
    s.bq = s.logq + Jp;
    s.wtot = s.bq + Jp;
    s.red = s.wtot + (size_t)(NT / 32) * Jp * 4;
    s.du = s.red + 32;
    s.tmp = s.du + 32;
    s.misc = s.tmp + 32;
    s.key = reinterpret_cast<unsigned long long*>(s.misc + 16);
    s.pts = reinterpret_cast<float4*>(raw + ((sinkhorn_smem<NT>(Jp) + 15) & ~(size_t)15));
    return s;
}

// ---- log-domain Sinkhorn iterations for one call ----------------------------------------------------------
// u[] (registers) and S.v (shared) start at 0; runs n_it iterations; records the change of each iteration.
template <int NT, int PPT, typename CostAt>
__device__ __forceinline__ void ld_iterations(const SinkhornParams& P, const Smem& S, int b, int o, int n_it,
                                              const float (&logp)[PPT], float (&u)[PPT], CostAt cost_at) {
    constexpr int NW = NT / 32;
    const int N = P.N, J = P.J, Jp = (J + kJC - 1) / kJC * kJC;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float k2 = (1.0f / P.eps) * kLog2e;
#pragma unroll
    for (int p = 0; p < PPT; ++p) u[p] = 0.f;
    __syncthreads();
    for (int j = tid; j < Jp; j += NT) S.v[j] = 0.f;
    __syncthreads();

    for (int it = 0; it < n_it; ++it) {
        // ---- row update: u_i += eps * (log p_i - LSE_j K_ij)
        float du_abs = 0.f;
#pragma unroll
        for (int p = 0; p < PPT; ++p) {
            if (tid + p * NT < N) {
                float m = -INFINITY, s = 0.f;
                for (int j0 = 0; j0 < J; j0 += kJC) {
                    float x[kJC];
                    float mc = -INFINITY;
#pragma unroll
                    for (int jj = 0; jj < kJC; ++jj) {
                        const int j = j0 + jj;
                        if (j < J) {
                            x[jj] = __fadd_rn(__fadd_rn(-cost_at(p, j), u[p]), S.v[j]) * k2;
                            mc = fmaxf(mc, x[jj]);
                        } else x[jj] = -INFINITY;
                    }
                    const float mn = fmaxf(m, mc);
                    float acc = 0.f;
#pragma unroll
                    for (int jj = 0; jj < kJC; ++jj) acc += exp2f(x[jj] - mn);
                    s = s * exp2f(m - mn) + acc;        // m == -inf on the first chunk -> s * 0
                    m = mn;
                }
                const float lse = (m + log2f(s)) * kLn2;
                const float un = __fadd_rn(__fmul_rn(P.eps, logp[p] - lse), u[p]);
                du_abs += fabsf(un - u[p]);
                u[p] = un;
            }
        }
        // ---- column update: v_j += eps * (log q_j - LSE_i K_ij); no max shift needed after a row update
        for (int j0 = 0; j0 < J; j0 += kJC) {
            float part[kJC];
#pragma unroll
            for (int jj = 0; jj < kJC; ++jj) part[jj] = 0.f;
#pragma unroll
            for (int p = 0; p < PPT; ++p) {
                if (tid + p * NT < N) {
#pragma unroll
                    for (int jj = 0; jj < kJC; ++jj) {
                        const int j = j0 + jj;
                        if (j < J) part[jj] += exp2f(__fadd_rn(__fadd_rn(-cost_at(p, j), u[p]), S.v[j]) * k2);
                    }
                }
            }
            const float tot = butterfly16(part, lane);
            if ((lane & 1) == 0) S.wtot[warp * Jp + j0 + ((lane >> 1) & 15)] = tot;
        }
        du_abs = warp_sum(du_abs);
        if (lane == 0) S.du[warp] = du_abs;
        __syncthreads();
        // final column sums (fixed order over warps) -> staged new v in wtot[j] (warp-0 row, which only this
        // thread reads); NaN marks a column that must be redone with a max shift
        int need_exact = 0;
        for (int j = tid; j < J; j += NT) {
            float sj = 0.f;
#pragma unroll
            for (int w = 0; w < NW; ++w) sj += S.wtot[w * Jp + j];
            if (sj > 1e-30f && sj < INFINITY) {
                S.wtot[j] = __fadd_rn(__fmul_rn(P.eps, S.logq[j] - logf(sj)), S.v[j]);
            } else {
                S.wtot[j] = NAN;
                need_exact = 1;
            }
        }
        if (__syncthreads_or(need_exact)) {
            for (int j = 0; j < J; ++j) {
                if (S.wtot[j] == S.wtot[j]) continue;               // block-uniform (shared value)
                float mx = -INFINITY;
#pragma unroll
                for (int p = 0; p < PPT; ++p)
                    if (tid + p * NT < N) mx = fmaxf(mx, __fadd_rn(__fadd_rn(-cost_at(p, j), u[p]), S.v[j]) * k2);
                mx = warp_max(mx);
                if (lane == 0) S.tmp[warp] = mx;
                __syncthreads();
                float mall = -INFINITY;
                for (int w = 0; w < NW; ++w) mall = fmaxf(mall, S.tmp[w]);
                float sm = 0.f;
#pragma unroll
                for (int p = 0; p < PPT; ++p)
                    if (tid + p * NT < N) sm += exp2f(__fadd_rn(__fadd_rn(-cost_at(p, j), u[p]), S.v[j]) * k2 - mall);
                sm = block_sum<NT>(sm, S.red);
                if (tid == 0) {
                    const float lse = (mall + log2f(sm)) * kLn2;
                    S.wtot[j] = __fadd_rn(__fmul_rn(P.eps, S.logq[j] - lse), S.v[j]);
                }
                __syncthreads();
            }
        }
        // commit v, sum |dv|, record this iteration's change
        float dv_abs = 0.f;
        for (int j = tid; j < J; j += NT) {
            const float vn = S.wtot[j];
            dv_abs += fabsf(vn - S.v[j]);
            S.v[j] = vn;
        }
        dv_abs = warp_sum(dv_abs);
        if (lane == 0) S.tmp[warp] = dv_abs;
        __syncthreads();
        if (tid == 0) {
            float du = 0.f, dv = 0.f;
#pragma unroll
            for (int w = 0; w < NW; ++w) { du += S.du[w]; dv += S.tmp[w]; }
            P.diffs[((int64_t)o * P.max_iter + it) * P.B + b] = du + dv;
        }
        __syncthreads();
    }
}

// ---- one cloud, outer iterations resume..iters-1 -----------------------------------------------------------
// `first`: main launch (every call runs max_iter inner iterations); otherwise counts come from P.n_inner.
template <int NT, int PPT, bool kCluster, bool kFast, bool kExactJ>
__device__ __forceinline__ void process_cloud(const SinkhornParams& P, const Smem& S, int b, int resume, bool first) {
    constexpr int NW = NT / 32;
    constexpr int JF = 16;                              // fast path register tile width
    const int N = P.N, J = kExactJ ? JF : P.J;                  // kExactJ: J == 16 known at compile time (no column masks)
    const int Jp = kFast ? JF : (J + kJC - 1) / kJC * kJC;      // compile-time in the fast kernel: fixed smem offsets
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int iters = P.iters, max_iter = P.max_iter;
    const float k2 = (1.0f / P.eps) * kLog2e;

    // ---- this cloud's points -------------------------------------------------------------------------------
    float m2x[PPT], m2y[PPT], m2z[PPT], pn[PPT], logp[PPT], u[PPT];
    float px[PPT], py[PPT], pz[PPT];
    __syncthreads();
    if constexpr (kCluster) {
        const float* base = P.xyz + (int64_t)b * P.sb;
        float osum = 0.f;
#pragma unroll
        for (int p = 0; p < PPT; ++p) {
            const int i = tid + p * NT;
            px[p] = py[p] = pz[p] = 0.f; logp[p] = 0.f;
            if (i < N) {
                px[p] = base[(int64_t)i * P.sn]; py[p] = base[(int64_t)i * P.sn + P.sc];
                pz[p] = base[(int64_t)i * P.sn + 2 * P.sc];
                logp[p] = P.o_scores[(int64_t)b * N + i];
                osum += logp[p];
            }
            m2x[p] = -2.f * px[p]; m2y[p] = -2.f * py[p]; m2z[p] = -2.f * pz[p];
            pn[p] = sq3(px[p], py[p], pz[p]);
            if constexpr (kFast) S.pts[i] = make_float4(m2x[p], m2y[p], m2z[p], pn[p]);      // read back by this thread only
        }
        // lib/utils.py:276  o / clip(sum o, 1e-4); logp holds p_i + 1e-8 (fast) or its log (log domain, :92)
        osum = fmaxf(block_sum<NT>(osum, S.red), 1e-4f);
#pragma unroll
        for (int p = 0; p < PPT; ++p) {
            const float pe = __fdiv_rn(logp[p], osum) + 1e-8f;
            logp[p] = kFast ? pe : logf(pe);
        }
        for (int j = tid; j < Jp; j += NT) S.logq[j] = kFast ? (1.0f / (float)J + 1e-8f) : logf(1.0f / (float)J + 1e-8f);
    } else {
#pragma unroll
        for (int p = 0; p < PPT; ++p) {
            const int i = tid + p * NT;
            const float pv = P.p ? (i < N ? P.p[(int64_t)b * N + i] : 1.f) : 1.0f / (float)N;
            logp[p] = logf(pv + 1e-8f);
            px[p] = py[p] = pz[p] = m2x[p] = m2y[p] = m2z[p] = pn[p] = 0.f;
        }
        for (int j = tid; j < Jp; j += NT) {
            const float qv = P.q ? (j < J ? P.q[(int64_t)b * J + j] : 1.f) : 1.0f / (float)J;
            S.logq[j] = logf(qv + 1e-8f);
        }
    }
    __syncthreads();

    // Fast path: the coordinates are not held in registers across the inner Sinkhorn loop, which needs every register
    // for the G tile; they come back from the thread's own shared-memory slot (one 16-byte load per point; x = -0.5 (-2x)
    // is exact).  Re-reading them from global memory cost 5.6 % of the kernel's instructions in 64-bit address arithmetic.
    auto reload_xyz = [&]() {
        if constexpr (kCluster && kFast) {
#pragma unroll
            for (int p = 0; p < PPT; ++p) {
                const float4 q = S.pts[tid + p * NT];
                m2x[p] = q.x; m2y[p] = q.y; m2z[p] = q.z; pn[p] = q.w;
                px[p] = -0.5f * q.x; py[p] = -0.5f * q.y; pz[p] = -0.5f * q.z;
            }
        }
    };
    const float* cost_b = kCluster ? nullptr : P.cost + (int64_t)b * N * J;
    auto cost_at = [&](int p, int j) -> float {
        if constexpr (kCluster) return node_cost(m2x[p], m2y[p], m2z[p], pn[p], S.node[j], P.tau);
        else return cost_b[(int64_t)(tid + p * NT) * J + j];
    };

    // ---- initial centroids ---------------------------------------------------------------------------------
    if constexpr (kCluster) {
        if (resume == 0) {
            float best[PPT];
#pragma unroll
            for (int p = 0; p < PPT; ++p) best[p] = 1e10f;
            const int far = fps_center_start<NT, PPT>(px, py, pz, best, N, S.red, S.key);
            fps_run<NT, PPT>(px, py, pz, best, N, J, far, S.misc, S.key, [&](int s, int f) {
                if ((f % NT) == tid) {                  // the owner of point f publishes it as centroid s
                    const int slot = f / NT;
#pragma unroll
                    for (int p = 0; p < PPT; ++p)
                        if (p == slot) S.node[s] = make_float4(px[p], py[p], pz[p], pn[p]);
                }
            });
        } else {
            const float* h = P.hist + ((int64_t)b * iters + resume) * J * 3;
            for (int j = tid; j < J; j += NT) {
                const float x = h[3 * j], y = h[3 * j + 1], z = h[3 * j + 2];
                S.node[j] = make_float4(x, y, z, sq3(x, y, z));
            }
        }
        for (int j = J + tid; j < Jp; j += NT) S.node[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncthreads();
    }

    // ---- outer iterations ----------------------------------------------------------------------------------
    for (int o = resume; o < iters; ++o) {
        if constexpr (kCluster) {
            float* h = P.hist + ((int64_t)b * iters + o) * J * 3;
            for (int j = tid; j < J; j += NT) { const float4 c = S.node[j]; h[3 * j] = c.x; h[3 * j + 1] = c.y; h[3 * j + 2] = c.z; }
        }
        const int n_it = first ? max_iter : __ldcg(P.n_inner + o);
        const bool last = (o == iters - 1);

        // K_ij of the thread's points after the Sinkhorn call: fast path a_i * G_ij * b_j, else recomputed
        float G[kFast ? PPT : 1][JF];
        float a[PPT];
        if constexpr (kFast) {
            // ---- scaled-domain iterations ------------------------------------------------------------------
            float wold[PPT];        // u_old - m_i, where u_i = m_i + eps log a_i
            reload_xyz();
            const bool unit_tau = P.tau == 1.0f;                 // block-uniform: one branch per row instead of one select per entry
#pragma unroll
            for (int p = 0; p < PPT; ++p) {
                float mn = INFINITY;
                if (unit_tau) {
#pragma unroll
                    for (int j = 0; j < JF; ++j) {
                        G[p][j] = (kExactJ || j < J) ? node_dist(m2x[p], m2y[p], m2z[p], pn[p], S.node[j]) : INFINITY;
                        mn = fminf(mn, G[p][j]);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < JF; ++j) {
                        G[p][j] = (kExactJ || j < J) ? __fdiv_rn(node_dist(m2x[p], m2y[p], m2z[p], pn[p], S.node[j]), P.tau) : INFINITY;
                        mn = fminf(mn, G[p][j]);
                    }
                }
                // a point past N gets -inf as its row minimum: every entry becomes exp2(-inf) = 0 without a select per entry
                // (masked columns hold +inf and come out as 0 the same way)
                const float mnx = (tid + p * NT < N) ? mn : -INFINITY;
#pragma unroll
                for (int j = 0; j < JF; ++j) G[p][j] = fast_exp2((mnx - G[p][j]) * k2);
                wold[p] = -mn;          // u_old = 0
                a[p] = 0.f;
            }
            if (tid < JF) { S.bq[tid] = tid < J ? 1.f : 0.f; S.v[tid] = 0.f; }
            if (tid == 0) S.misc[6] = 0.f;
            __syncthreads();
            bool tripped = false;
            const float eps_ln2 = P.eps * 0.6931471805599453f;
            for (int it = 0; it < n_it; ++it) {
                // row sums r_i = sum_j G_ij b_j and, below, column sums sum_i G_ij a_i as packed FFMA2 on column pairs
                float r[PPT];
                {
                    float2 r2[PPT];
#pragma unroll
                    for (int p = 0; p < PPT; ++p) r2[p] = make_float2(0.f, 0.f);
#pragma unroll
                    for (int j4 = 0; j4 < JF / 4; ++j4) {
                        const float4 bq = *reinterpret_cast<const float4*>(S.bq + 4 * j4);
#pragma unroll
                        for (int p = 0; p < PPT; ++p) {
                            ffma2_pair(r2[p], make_float2(G[p][4 * j4 + 0], G[p][4 * j4 + 1]), make_float2(bq.x, bq.y));
                            ffma2_pair(r2[p], make_float2(G[p][4 * j4 + 2], G[p][4 * j4 + 3]), make_float2(bq.z, bq.w));
                        }
                    }
#pragma unroll
                    for (int p = 0; p < PPT; ++p) r[p] = r2[p].x + r2[p].y;
                }
                float du_abs = 0.f;
#pragma unroll
                for (int p = 0; p < PPT; ++p) {
                    if (tid + p * NT < N) {
                        // a_i = p_i / r_i and eps ln a_i through MUFU (rcp / lg2, ~1 ulp): the scaled iteration is already
                        // only tolerance-equal to the log-domain reference, and IEEE division cost 7 % of the kernel
                        a[p] = logp[p] * fast_rcp(r[p]);
                        const float wn = eps_ln2 * fast_lg2(a[p]);
                        du_abs += fabsf(wn - wold[p]);
                        wold[p] = wn;
                    }
                }
                float part[JF];
#pragma unroll
                for (int j2 = 0; j2 < JF / 2; ++j2) {
                    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
                    for (int p = 0; p < PPT; ++p) ffma2_pair(acc, make_float2(G[p][2 * j2], G[p][2 * j2 + 1]), make_float2(a[p], a[p]));
                    part[2 * j2] = acc.x; part[2 * j2 + 1] = acc.y;
                }
                const float tot = butterfly16(part, lane);
                if ((lane & 1) == 0) S.wtot[warp * JF + ((lane >> 1) & 15)] = tot;
                du_abs = warp_sum(du_abs);
                if (lane == 0) S.du[warp] = du_abs;
                __syncthreads();
                if (warp == 0) {
                    const int j = lane & 15;
                    float sj = 0.f;
#pragma unroll
                    for (int w = 0; w < NW; ++w) sj += S.wtot[w * JF + j];
                    const bool live = j < J;
                    const float bn = live ? S.logq[j] * fast_rcp(sj) : 0.f;
                    const float vn = live ? eps_ln2 * fast_lg2(bn) : 0.f;
                    float dv = (live && lane < 16) ? fabsf(vn - S.v[j]) : 0.f;
                    bool ok = !live || (sj > 1e-30f && sj < 1e30f);
                    float bmax = live ? bn : 0.f, bmin = live ? bn : INFINITY;
#pragma unroll
                    for (int off = 8; off > 0; off >>= 1) {
                        dv += __shfl_xor_sync(kFull, dv, off);
                        bmax = fmaxf(bmax, __shfl_xor_sync(kFull, bmax, off));
                        bmin = fminf(bmin, __shfl_xor_sync(kFull, bmin, off));
                    }
                    ok = __all_sync(kFull, ok) && (bmax <= bmin * 1e24f);
                    float du = 0.f;
#pragma unroll
                    for (int w = 0; w < NW; ++w) du += S.du[w];
                    if (lane < 16) { S.bq[j] = bn; S.v[j] = vn; }
                    if (lane == 0) {
                        P.diffs[((int64_t)o * max_iter + it) * P.B + b] = du + dv;
                        if (!ok) S.misc[6] = 1.f;
                    }
                }
                __syncthreads();
                if (S.misc[6] != 0.f) { tripped = true; break; }
            }
            if (!tripped) {
                // ---- fast post-processing: gamma_ij = K_ij / clip(sum_j K_ij, 1e-3) with K = a G b (finite by the
                // monitor, so nan_to_num is the identity); M-step sums over t_ij = scale_i G_ij, b_j applied after
                // the reduction (it is constant along i).
                reload_xyz();
                float scale[PPT];
                {
                    float rp[PPT];
#pragma unroll
                    for (int p = 0; p < PPT; ++p) rp[p] = 0.f;
#pragma unroll
                    for (int j4 = 0; j4 < JF / 4; ++j4) {
                        const float4 bq = *reinterpret_cast<const float4*>(S.bq + 4 * j4);
#pragma unroll
                        for (int p = 0; p < PPT; ++p) {
                            rp[p] = fmaf(G[p][4 * j4 + 0], bq.x, rp[p]);
                            rp[p] = fmaf(G[p][4 * j4 + 1], bq.y, rp[p]);
                            rp[p] = fmaf(G[p][4 * j4 + 2], bq.z, rp[p]);
                            rp[p] = fmaf(G[p][4 * j4 + 3], bq.w, rp[p]);
                        }
                    }
#pragma unroll
                    for (int p = 0; p < PPT; ++p) scale[p] = a[p] * __frcp_rn(fmaxf(a[p] * rp[p], 1e-3f));
                }
                float a0[JF], ax[JF], ay[JF], az[JF];
#pragma unroll
                for (int j = 0; j < JF; ++j) {
                    float s0 = 0.f, sx = 0.f, sy = 0.f, sz = 0.f;
#pragma unroll
                    for (int p = 0; p < PPT; ++p) {
                        const float t = scale[p] * G[p][j];
                        G[p][j] = t;
                        s0 += t;
                        sx = fmaf(t, px[p], sx); sy = fmaf(t, py[p], sy); sz = fmaf(t, pz[p], sz);
                    }
                    a0[j] = s0; ax[j] = sx; ay[j] = sy; az[j] = sz;
                }
                if (last) {
#pragma unroll
                    for (int p = 0; p < PPT; ++p) {
                        const int i = tid + p * NT;
                        if (i < N) {
                            float* grow = P.gamma + ((int64_t)b * N + i) * J;
                            if (kExactJ || (J & 3) == 0) {
#pragma unroll
                                for (int j4 = 0; j4 < JF / 4; ++j4) {
                                    if (4 * j4 < J) {
                                        const float4 bq = *reinterpret_cast<const float4*>(S.bq + 4 * j4);
                                        *reinterpret_cast<float4*>(grow + 4 * j4) = make_float4(
                                            G[p][4 * j4] * bq.x, G[p][4 * j4 + 1] * bq.y, G[p][4 * j4 + 2] * bq.z, G[p][4 * j4 + 3] * bq.w);
                                    }
                                }
                            } else {
#pragma unroll
                                for (int j = 0; j < JF; ++j)
                                    if (j < J) grow[j] = G[p][j] * S.bq[j];
                            }
                        }
                    }
                }
                const float t0 = butterfly16(a0, lane);
                const float tx = butterfly16(ax, lane);
                const float ty = butterfly16(ay, lane);
                const float tz = butterfly16(az, lane);
                if ((lane & 1) == 0) {
                    const int col = (lane >> 1) & 15;
                    S.wtot[(0 * NW + warp) * JF + col] = t0;
                    S.wtot[(1 * NW + warp) * JF + col] = tx;
                    S.wtot[(2 * NW + warp) * JF + col] = ty;
                    S.wtot[(3 * NW + warp) * JF + col] = tz;
                }
                __syncthreads();
                if (tid < J) {
                    const int j = tid;
                    float s0 = 0.f, sx = 0.f, sy = 0.f, sz = 0.f;
#pragma unroll
                    for (int w = 0; w < NW; ++w) {
                        s0 += S.wtot[(0 * NW + w) * JF + j]; sx += S.wtot[(1 * NW + w) * JF + j];
                        sy += S.wtot[(2 * NW + w) * JF + j]; sz += S.wtot[(3 * NW + w) * JF + j];
                    }
                    const float bq = S.bq[j];
                    s0 *= bq; sx *= bq; sy *= bq; sz *= bq;
                    const float pi = __fdiv_rn(s0, (float)N);
                    const float npi = __fadd_rn(__fmul_rn(pi, (float)N), 1e-5f);
                    const float mx = __fdiv_rn(sx, npi), my = __fdiv_rn(sy, npi), mz = __fdiv_rn(sz, npi);
                    S.node[j] = make_float4(mx, my, mz, sq3(mx, my, mz));
                    if (last) {
                        P.pi[(int64_t)b * J + j] = pi;
                        float* m = P.mu + ((int64_t)b * J + j) * 3;
                        m[0] = mx; m[1] = my; m[2] = mz;
                    }
                }
                __syncthreads();
                continue;           // next outer iteration
            } else {
                if (tid == 0) atomicAdd(&P.state[2], 1);
                // rescue: redo this call in the log domain (logp holds p+1e-8 here; the log-domain code wants logs)
                float lp[PPT];
                reload_xyz();
#pragma unroll
                for (int p = 0; p < PPT; ++p) lp[p] = logf(logp[p]);
                __syncthreads();
                for (int j = tid; j < Jp; j += NT) S.logq[j] = logf(S.logq[j]);
                ld_iterations<NT, PPT>(P, S, b, o, n_it, lp, u, cost_at);
                for (int j = tid; j < Jp; j += NT) S.logq[j] = 1.0f / (float)J + 1e-8f;
                __syncthreads();
            }
        } else {
            ld_iterations<NT, PPT>(P, S, b, o, n_it, logp, u, cost_at);
        }

        auto k_at = [&](int p, int jj, int j0) -> float {
            const float c = cost_at(p, j0 + jj);
            return exp2f(__fadd_rn(__fadd_rn(-c, u[p]), S.v[j0 + jj]) * k2);
        };

        // ---- gamma = exp(K); nan_to_num; row normalise; M-step on xyz ----------------------------------------
        if constexpr (kCluster) {
            float rinv[PPT];
#pragma unroll
            for (int p = 0; p < PPT; ++p) {
                float rs = 0.f;
                if (tid + p * NT < N) {
                    for (int j0 = 0; j0 < J; j0 += kJC)
#pragma unroll
                        for (int jj = 0; jj < kJC; ++jj)
                            if (j0 + jj < J) rs += nan_to_num(k_at(p, jj, j0), 0.f);
                }
                rinv[p] = __frcp_rn(fmaxf(rs, 1e-3f));
            }
            for (int j0 = 0; j0 < J; j0 += kJC) {
                float a0[kJC], ax[kJC], ay[kJC], az[kJC];
#pragma unroll
                for (int jj = 0; jj < kJC; ++jj) a0[jj] = ax[jj] = ay[jj] = az[jj] = 0.f;
#pragma unroll
                for (int p = 0; p < PPT; ++p) {
                    const int i = tid + p * NT;
                    if (i < N) {
                        float g[kJC];
#pragma unroll
                        for (int jj = 0; jj < kJC; ++jj) {
                            g[jj] = 0.f;
                            if (j0 + jj < J) {
                                g[jj] = nan_to_num(k_at(p, jj, j0), 0.f) * rinv[p];
                                a0[jj] += g[jj];
                                ax[jj] = fmaf(g[jj], px[p], ax[jj]);
                                ay[jj] = fmaf(g[jj], py[p], ay[jj]);
                                az[jj] = fmaf(g[jj], pz[p], az[jj]);
                            }
                        }
                        if (last) {
                            float* grow = P.gamma + ((int64_t)b * N + i) * J + j0;
                            if ((J & 3) == 0) {
#pragma unroll
                                for (int jj = 0; jj < kJC; jj += 4)
                                    if (j0 + jj < J)
                                        *reinterpret_cast<float4*>(grow + jj) = make_float4(g[jj], g[jj + 1], g[jj + 2], g[jj + 3]);
                            } else {
#pragma unroll
                                for (int jj = 0; jj < kJC; ++jj)
                                    if (j0 + jj < J) grow[jj] = g[jj];
                            }
                        }
                    }
                }
                const float t0 = butterfly16(a0, lane);
                const float tx = butterfly16(ax, lane);
                const float ty = butterfly16(ay, lane);
                const float tz = butterfly16(az, lane);
                if ((lane & 1) == 0) {
                    const int col = j0 + ((lane >> 1) & 15);
                    S.wtot[(0 * NW + warp) * Jp + col] = t0;
                    S.wtot[(1 * NW + warp) * Jp + col] = tx;
                    S.wtot[(2 * NW + warp) * Jp + col] = ty;
                    S.wtot[(3 * NW + warp) * Jp + col] = tz;
                }
            }
            __syncthreads();
            for (int j = tid; j < J; j += NT) {
                float s0 = 0.f, sx = 0.f, sy = 0.f, sz = 0.f;
#pragma unroll
                for (int w = 0; w < NW; ++w) {
                    s0 += S.wtot[(0 * NW + w) * Jp + j]; sx += S.wtot[(1 * NW + w) * Jp + j];
                    sy += S.wtot[(2 * NW + w) * Jp + j]; sz += S.wtot[(3 * NW + w) * Jp + j];
                }
                // lib/utils.py:137-140: pi = mean; npi = pi*N + 1e-5; mu = sum / npi
                const float pi = __fdiv_rn(s0, (float)N);
                const float npi = __fadd_rn(__fmul_rn(pi, (float)N), 1e-5f);
                const float mx = __fdiv_rn(sx, npi), my = __fdiv_rn(sy, npi), mz = __fdiv_rn(sz, npi);
                S.node[j] = make_float4(mx, my, mz, sq3(mx, my, mz));
                if (last) {
                    P.pi[(int64_t)b * J + j] = pi;
                    float* m = P.mu + ((int64_t)b * J + j) * 3;
                    m[0] = mx; m[1] = my; m[2] = mz;
                }
            }
            __syncthreads();
        } else {
            // matrix mode: gamma = exp(K) as is, loss_b = sum gamma * cost
            float lsum = 0.f;
#pragma unroll
            for (int p = 0; p < PPT; ++p) {
                const int i = tid + p * NT;
                if (i < N) {
                    float* grow = P.gamma + ((int64_t)b * N + i) * J;
                    for (int j = 0; j < J; ++j) {
                        const float c = cost_at(p, j);
                        const float e = exp2f(__fadd_rn(__fadd_rn(-c, u[p]), S.v[j]) * k2);
                        grow[j] = e;
                        lsum = fmaf(e, c, lsum);
                    }
                }
            }
            lsum = block_sum<NT>(lsum, S.red);
            if (tid == 0 && P.loss) P.loss[b] = lsum;
            __syncthreads();
        }
    }
}

// ---- batch-mean exit test: means for (resume.., all inner), then the schedule update (one CTA) ---------------
template <int NT>
__device__ __forceinline__ void verify_schedule(const SinkhornParams& P, int resume, bool first) {
    constexpr int NW = NT / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int iters = P.iters, max_iter = P.max_iter;
    const int total = (iters - resume) * max_iter;
    for (int e = warp; e < total; e += NW) {
        const int o = resume + e / max_iter, it = e % max_iter;
        const int n_it = first ? max_iter : P.n_inner[o];
        if (it >= n_it) continue;
        const float* d = P.diffs + ((int64_t)o * max_iter + it) * P.B;
        float acc = 0.f;
        for (int bb = lane; bb < P.B; bb += 32) acc += __ldcg(d + bb);
        acc = warp_sum(acc);
        if (lane == 0) P.means[o * max_iter + it] = acc / (float)P.B;
    }
    __syncthreads();
    if (tid == 0) {
        int new_resume = iters;
        for (int o = resume; o < iters; ++o) {
            const int n_it = first ? max_iter : P.n_inner[o];
            int run = n_it;
            for (int it = 0; it < n_it; ++it)
                if (P.means[o * max_iter + it] < P.thresh) { run = it + 1; break; }
            P.n_inner[o] = run;
            if (run < n_it) { new_resume = o; break; }
        }
        for (int o = new_resume + 1; o < iters; ++o) P.n_inner[o] = max_iter;   // restart later calls in full
        if (P.iters_run)
            for (int o = 0; o < iters; ++o) P.iters_run[o] = P.n_inner[o];
        P.state[0] = new_resume;
        P.state[1] = 0;
        __threadfence();
    }
    __syncthreads();
}

// mode 0: main launch (from the host), one CTA per cloud, every Sinkhorn call runs max_iter iterations.
// mode 1: redo round (tail-launched by the previous round's last CTA): re-runs from outer iteration state[0] with the
//         inner counts of n_inner[].
// In both modes the last CTA to finish evaluates the batch-mean exit test, publishes the next `resume` and, if the
// schedule moved, tail-launches the next round.  (One kernel for both so the cloud body is instantiated once.)
template <int NT, int PPT, bool kCluster, bool kFast, bool kExactJ>
__global__ void __launch_bounds__(NT, (kFast && NT == 256) ? 2 : 1)
sinkhorn_kernel(SinkhornParams P, int mode) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int resume = 0;
    if (mode != 0) {
        resume = __ldcg(P.state);
        if (resume >= P.iters) return;
    }
    const int Jp = kFast ? 16 : (P.J + kJC - 1) / kJC * kJC;
    const Smem S = carve_smem<NT>(smem_raw, Jp);
    for (int b = blockIdx.x; b < P.B; b += gridDim.x) process_cloud<NT, PPT, kCluster, kFast, kExactJ>(P, S, b, resume, mode == 0);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int prev = atomicAdd(&P.state[1], 1);
        S.misc[5] = (prev == (int)gridDim.x - 1) ? 1.f : 0.f;
    }
    __syncthreads();
    if (S.misc[5] != 0.f) {
        __threadfence();
        verify_schedule<NT>(P, resume, mode == 0);       // also resets the done counter for the next round
        if (threadIdx.x == 0 && *reinterpret_cast<volatile int32_t*>(P.state) < P.iters) {
            unsigned dyn_smem;
            asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn_smem));
            sinkhorn_kernel<NT, PPT, kCluster, kFast, kExactJ><<<gridDim.x, NT, dyn_smem, cudaStreamTailLaunch>>>(P, 1);
            if (cudaGetLastError() != cudaSuccess) atomicExch(&P.state[5], 1);      // surfaced by ogmm_sinkhorn_status (tests)
        }
    }
}

}  // namespace ogmm

using namespace ogmm;

// ---- dispatch helpers --------------------------------------------------------------------------------------
// (threads per CTA, points per thread) by cloud size.  N > 8192: the 16-CTA cluster variant (cluster_dsmem.cu).
#define OGMM_DISPATCH_POINTS(N, CALL)                                   \
    do {                                                                \
        if ((N) <= 256)        { CALL(256, 1); }                        \
        else if ((N) <= 512)   { CALL(256, 2); }                        \
        else if ((N) <= 1024)  { CALL(256, 4); }                        \
        else if ((N) <= 2048)  { CALL(256, 8); }                        \
        else if ((N) <= 4096)  { CALL(1024, 4); }                       \
        else                   { CALL(1024, 8); }                       \
    } while (0)

constexpr int64_t kMaxPoints = 8192;
constexpr int64_t kMaxClusterPoints = 16384;      // clustering only: 16-CTA cluster variant (cluster_dsmem.cu)

// defined in cluster_dsmem.cu: clouds of 8193..16384 points, one cloud per 16-CTA thread-block cluster
int ogmm_launch_cluster_dsmem(ogmm::SinkhornParams P, cudaStream_t s);

template <int NT, int PPT, bool kCluster, bool kFast, bool kExactJ = false>
static inline int launch_sinkhorn_variant(SinkhornParams P, cudaStream_t s) {
    const size_t smem = sinkhorn_smem<NT>(P.J, (kCluster && kFast) ? NT * PPT : 0);
    OGMM_REQUIRE(smem <= 200 * 1024, OGMM_EUNSUPPORTED, "sinkhorn: J=%d needs %zu B of shared memory (> 200 KiB)", P.J, smem);
    auto kern = sinkhorn_kernel<NT, PPT, kCluster, kFast, kExactJ>;
    int st;
    if (smem > 48 * 1024) {
        st = cuda_status(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute(sinkhorn_kernel)");
        if (st != OGMM_OK) return st;
    }
    // one launch from the host; redo rounds, when the early exit fires, are tail-launched from the device (header comment)
    kern<<<(unsigned)P.B, NT, smem, s>>>(P, 0);
    return cuda_status(cudaGetLastError(), "sinkhorn_kernel");
}

template <bool kCluster>
static inline int launch_sinkhorn(SinkhornParams P, void* workspace, int64_t workspace_bytes, ogmm_stream_t stream) {
    const ClusterWsLayout l = cluster_ws_layout(P.B, kCluster ? P.J : 1, P.iters, P.max_iter);
    OGMM_REQUIRE(workspace != nullptr && workspace_bytes >= l.total, OGMM_EWORKSPACE,
                 "sinkhorn: workspace of %lld B given, %lld B needed", (long long)workspace_bytes, (long long)l.total);
    OGMM_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, OGMM_EINVAL, "sinkhorn: workspace must be 256-B aligned");
    char* ws = static_cast<char*>(workspace);
    P.state = reinterpret_cast<int32_t*>(ws + l.state_off);
    P.n_inner = reinterpret_cast<int32_t*>(ws + l.ninner_off);
    P.means = reinterpret_cast<float*>(ws + l.means_off);
    P.diffs = reinterpret_cast<float*>(ws + l.diffs_off);
    P.hist = reinterpret_cast<float*>(ws + l.hist_off);
    cudaStream_t s = as_stream(stream);
    int st = cuda_status(cudaMemsetAsync(ws, 0, 64, s), "cudaMemsetAsync(workspace header)");
    if (st != OGMM_OK) return st;
    if constexpr (kCluster) {
        if (P.N > kMaxPoints) return ogmm_launch_cluster_dsmem(P, s);
        if (P.J <= 16 && P.N <= 1024) {
            if (P.N <= 256) return launch_sinkhorn_variant<256, 1, true, true>(P, s);
            if (P.N <= 512) return launch_sinkhorn_variant<256, 2, true, true>(P, s);
            if (P.J == 16) {
                // OGMM_CLUSTER_NT=128 selects 4 warps x 8 points per thread (half the per-warp overheads per cloud, but 255
                // registers and 8 warps per SM): measured slower on B200, 0.534 vs 0.462 ms per 2 x 256 clouds
                const char* nt = getenv("OGMM_CLUSTER_NT");
                if (nt && nt[0] == '1') return launch_sinkhorn_variant<128, 8, true, true, true>(P, s);
                return launch_sinkhorn_variant<256, 4, true, true, true>(P, s);
            }
            return launch_sinkhorn_variant<256, 4, true, true>(P, s);
        }
    }
#define CALL(NT, PPT) return launch_sinkhorn_variant<NT, PPT, kCluster, false>(P, s)
    OGMM_DISPATCH_POINTS(P.N, CALL);
#undef CALL
    return OGMM_OK;
}
