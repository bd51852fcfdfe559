// K2: farthest-point init + overlap-guided Sinkhorn k-means, and the stand-alone Sinkhorn (sm_100a).
//
// One CTA per cloud; every point lives in the registers of one thread for the whole
// 10 x 10 iteration loop (lib/utils.py:269-288), so the only HBM traffic is the algorithmic
// one: xyz + overlap scores in, gamma / pi / mu out.  The cost matrix is never stored: a cost
// c_ij is recomputed from the point (registers) and the centroid (shared memory) where needed.
//
// Reference arithmetic reproduced here
//   FPS init             lib/utils.py:170-198 (is_center branch :183-188)
//   cost                 lib/utils.py:280  torch.cdist -> ATen _euclidean_dist (matmul form, K=5),
//                        clamp_min(0).sqrt(), .clip(0) / tau
//   Sinkhorn             lib/utils.py:69-108 log domain; row LSE over j, column LSE over i
//   early exit           lib/utils.py:99-102 BATCH mean of sum|du|+sum|dv| < thresh (see below)
//   post                 lib/utils.py:282 nan_to_num, :287 gamma / clip(sum_j gamma, 1e-3)
//   M-step (xyz)         lib/utils.py:130-140
//
// Batch-coupled early exit without a host sync.  Iteration counts couple the clouds of one call
// only through the exit test.  Launch 0 runs every Sinkhorn call for max_iter iterations and each
// CTA records its cloud's change per (outer, inner) iteration; the last CTA to finish evaluates the
// batch means in order and, at the first (outer o, inner i) with mean < thresh and i+1 below the
// count that was run, stores n_inner[o] = i+1 and resume = o.  Follow-up launches (queued
// unconditionally; they return at once when resume == iters) restart from the centroids saved at
// the start of outer iteration `resume`.  Each follow-up certifies at least one more outer
// iteration, so `iters` follow-ups always suffice.  Results are bit-identical to running the exit
// test inline, and deterministic (fixed-order reductions, no float atomics).
#pragma once
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
    l.state_off = 0;                                         // int32[16]: [0]=resume  [1]=done counter
    l.ninner_off = 64;                                       // int32[iters]
    l.means_off = align_up(l.ninner_off + 4 * iters, 256);   // float[iters][max_iter] batch means
    l.diffs_off = align_up(l.means_off + 4 * iters * max_iter, 256);   // float[iters][max_iter][B]
    l.hist_off = align_up(l.diffs_off + 4 * iters * max_iter * B, 256);   // float[B][iters][J][3]
    l.total = align_up(l.hist_off + 4 * B * iters * J * 3, 256);
    return l;
}

// ---- farthest point sampling on register-resident points ----------------------------------------
// lib/utils.py:191-197.  `far` is the first index; writes `npoint` indices through `emit`.
template <int NT, int PPT, typename Emit>
__device__ __forceinline__ void fps_run(const float (&px)[PPT], const float (&py)[PPT], const float (&pz)[PPT],
                                        float (&best)[PPT], int N, int npoint, int far, float* s_pick,
                                        unsigned long long* s_key, Emit emit) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = NT / 32;
    for (int s = 0; s < npoint; ++s) {
        emit(s, far);
        // broadcast the coordinates of point `far`
        if ((far % NT) == tid) {
            const int slot = far / NT;
#pragma unroll
            for (int p = 0; p < PPT; ++p)
                if (p == slot) { s_pick[0] = px[p]; s_pick[1] = py[p]; s_pick[2] = pz[p]; }
        }
        __syncthreads();
        const float cx = s_pick[0], cy = s_pick[1], cz = s_pick[2];
        unsigned long long key = 0ull;
#pragma unroll
        for (int p = 0; p < PPT; ++p) {
            const int i = tid + p * NT;
            if (i < N) {
                float dx = px[p] - cx, dy = py[p] - cy, dz = pz[p] - cz;
                float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                if (d < best[p]) best[p] = d;
                unsigned long long k = ((unsigned long long)__float_as_uint(best[p]) << 32) | (0xffffffffu - (unsigned)i);
                key = k > key ? k : key;
            }
        }
        key = warp_max_u64(key);
        if (lane == 0) s_key[warp] = key;
        __syncthreads();
        unsigned long long k2 = lane < NW ? s_key[lane] : 0ull;
        k2 = warp_max_u64(k2);
        far = (int)(0xffffffffu - (unsigned)(k2 & 0xffffffffull));
        // s_key / s_pick are rewritten only after the next __syncthreads pair
    }
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
    float sx = 0.f, sy = 0.f, sz = 0.f;
#pragma unroll
    for (int p = 0; p < PPT; ++p) {
        const int i = tid + p * NT;
        px[p] = py[p] = pz[p] = 0.f;
        best[p] = 1e10f;
        if (i < N) {
            px[p] = base[(int64_t)i * sn]; py[p] = base[(int64_t)i * sn + sc]; pz[p] = base[(int64_t)i * sn + 2 * sc];
            sx += px[p]; sy += py[p]; sz += pz[p];
        }
    }
    int far;
    if (start == nullptr) {
        // is_center: relax against the centroid first, start from the farthest point (:183-188)
        const float cx = block_sum<NT>(sx, s_red) / (float)N;
        const float cy = block_sum<NT>(sy, s_red) / (float)N;
        const float cz = block_sum<NT>(sz, s_red) / (float)N;
        unsigned long long key = 0ull;
#pragma unroll
        for (int p = 0; p < PPT; ++p) {
            const int i = tid + p * NT;
            if (i < N) {
                float dx = px[p] - cx, dy = py[p] - cy, dz = pz[p] - cz;
                float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                if (d < best[p]) best[p] = d;
                unsigned long long k = ((unsigned long long)__float_as_uint(best[p]) << 32) | (0xffffffffu - (unsigned)i);
                key = k > key ? k : key;
            }
        }
        key = warp_max_u64(key);
        __syncthreads();
        if ((tid & 31) == 0) s_key[tid >> 5] = key;
        __syncthreads();
        unsigned long long k2 = (tid & 31) < NT / 32 ? s_key[tid & 31] : 0ull;
        k2 = warp_max_u64(k2);
        far = (int)(0xffffffffu - (unsigned)(k2 & 0xffffffffull));
        __syncthreads();
    } else {
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

// ---- cost sources -----------------------------------------------------------------------------------
// Cluster mode: c_ij = cdist(x_i, node_j) / tau from registers + shared memory.
struct NodeCost {
    const float4* node;     // shared: (x, y, z, |n|^2) per centroid
    float inv_tau_is_one;   // tau == 1 -> skip the division (x / 1 == x)
    float tau;
    __device__ __forceinline__ float operator()(float m2x, float m2y, float m2z, float pn, int j) const {
        const float4 c = node[j];
        // ATen _euclidean_dist: [-2x, |x|^2, 1] . [y, 1, |y|^2], accumulated in k order
        float d2 = __fmul_rn(m2x, c.x);
        d2 = fmaf(m2y, c.y, d2);
        d2 = fmaf(m2z, c.z, d2);
        d2 = __fadd_rn(d2, pn);
        d2 = __fadd_rn(d2, c.w);
        float d = sqrtf(fmaxf(d2, 0.f));
        d = fmaxf(d, 0.f);
        return inv_tau_is_one != 0.f ? d : __fdiv_rn(d, tau);
    }
};

struct SinkhornParams {
    // geometry / marginals
    const float* xyz; int64_t sb, sn, sc;      // cluster mode
    const float* o_scores;                     // cluster mode (B,N)
    const float* cost;                         // matrix mode (B,N,J)
    const float* p;                            // matrix mode (B,N) or null
    const float* q;                            // matrix mode (B,J) or null
    int B, N, J, iters, max_iter;
    float tau, eps, thresh;
    // outputs
    float* gamma; float* pi; float* mu; float* loss; int32_t* iters_run;
    // workspace
    int32_t* state; int32_t* n_inner; float* means; float* diffs; float* hist;
    int launch;
};

// Shared memory carve-up (floats): node float4[Jp] | v[Jp] | logq[Jp] | wtot[4][NW][Jp] (one plane in
// the Sinkhorn loop, four in the M-step) | red[32] | du[32] | tmp[32] | misc[16] | key u64[32]
template <int NT>
__host__ __device__ inline size_t sinkhorn_smem(int J) {
    const int Jp = (J + kJC - 1) / kJC * kJC;
    return sizeof(float) * ((size_t)4 * Jp + Jp + Jp + (size_t)(NT / 32) * Jp * 4 + 32 + 32 + 32 + 16) +
           sizeof(unsigned long long) * 32;
}

template <int NT, int PPT, bool kCluster>
__global__ void __launch_bounds__(NT)
sinkhorn_kernel(SinkhornParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NW = NT / 32;
    const int N = P.N, J = P.J;
    const int Jp = (J + kJC - 1) / kJC * kJC;
    float4* s_node = reinterpret_cast<float4*>(smem_raw);
    float* s_v = reinterpret_cast<float*>(s_node + Jp);
    float* s_logq = s_v + Jp;
    float* s_wtot = s_logq + Jp;                         // [NW][Jp] x 4 planes
    float* s_red = s_wtot + (size_t)NW * Jp * 4;         // block_sum scratch
    float* s_du = s_red + 32;                            // per-warp sum |du|
    float* s_tmp = s_du + 32;                            // per-warp scratch (sum |dv|, column max)
    float* s_misc = s_tmp + 32;                          // [0..3] FPS pick, [5] last-CTA flag
    unsigned long long* s_key = reinterpret_cast<unsigned long long*>(s_misc + 16);

    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int iters = P.iters, max_iter = P.max_iter;

    int resume = 0;
    if (P.launch > 0) {
        resume = P.state[0];
        if (resume >= iters) return;
    }

    // ---- load this cloud's points into registers ---------------------------------------------------
    float m2x[PPT], m2y[PPT], m2z[PPT], pn[PPT], logp[PPT], u[PPT];
    float px[PPT], py[PPT], pz[PPT];
    const float k2 = (1.0f / P.eps) * kLog2e;      // exponent scale for exp2
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
            pn[p] = __fadd_rn(__fadd_rn(__fmul_rn(px[p], px[p]), __fmul_rn(py[p], py[p])), __fmul_rn(pz[p], pz[p]));
        }
        // lib/utils.py:276  o / clip(sum o, 1e-4); then log(p + 1e-8) (:92)
        osum = fmaxf(block_sum<NT>(osum, s_red), 1e-4f);
#pragma unroll
        for (int p = 0; p < PPT; ++p) logp[p] = logf(__fdiv_rn(logp[p], osum) + 1e-8f);
        for (int j = tid; j < Jp; j += NT) s_logq[j] = logf(1.0f / (float)J + 1e-8f);
    } else {
#pragma unroll
        for (int p = 0; p < PPT; ++p) {
            const int i = tid + p * NT;
            float pv = P.p ? (i < N ? P.p[(int64_t)b * N + i] : 1.f) : 1.0f / (float)N;
            logp[p] = logf(pv + 1e-8f);
            px[p] = py[p] = pz[p] = m2x[p] = m2y[p] = m2z[p] = pn[p] = 0.f;
        }
        for (int j = tid; j < Jp; j += NT) {
            float qv = P.q ? (j < J ? P.q[(int64_t)b * J + j] : 1.f) : 1.0f / (float)J;
            s_logq[j] = logf(qv + 1e-8f);
        }
    }
    __syncthreads();

    NodeCost ncost{s_node, P.tau == 1.0f ? 1.f : 0.f, P.tau};
    const float* cost_b = kCluster ? nullptr : P.cost + (int64_t)b * N * J;
    auto cost_at = [&](int p, int j) -> float {
        if constexpr (kCluster) return ncost(m2x[p], m2y[p], m2z[p], pn[p], j);
        else return cost_b[(int64_t)(tid + p * NT) * J + j];
    };

    // ---- initial centroids ------------------------------------------------------------------------------
    if constexpr (kCluster) {
        float* hist_b = P.hist + (int64_t)b * iters * J * 3;
        if (resume == 0) {
            float best[PPT];
            float sx = 0.f, sy = 0.f, sz = 0.f;
#pragma unroll
            for (int p = 0; p < PPT; ++p) { best[p] = 1e10f; if (tid + p * NT < N) { sx += px[p]; sy += py[p]; sz += pz[p]; } }
            const float cx = block_sum<NT>(sx, s_red) / (float)N;
            const float cy = block_sum<NT>(sy, s_red) / (float)N;
            const float cz = block_sum<NT>(sz, s_red) / (float)N;
            unsigned long long key = 0ull;
#pragma unroll
            for (int p = 0; p < PPT; ++p) {
                const int i = tid + p * NT;
                if (i < N) {
                    float dx = px[p] - cx, dy = py[p] - cy, dz = pz[p] - cz;
                    float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                    if (d < best[p]) best[p] = d;
                    unsigned long long k = ((unsigned long long)__float_as_uint(best[p]) << 32) | (0xffffffffu - (unsigned)i);
                    key = k > key ? k : key;
                }
            }
            key = warp_max_u64(key);
            __syncthreads();
            if (lane == 0) s_key[warp] = key;
            __syncthreads();
            unsigned long long kk = lane < NW ? s_key[lane] : 0ull;
            kk = warp_max_u64(kk);
            int far = (int)(0xffffffffu - (unsigned)(kk & 0xffffffffull));
            __syncthreads();
            fps_run<NT, PPT>(px, py, pz, best, N, J, far, s_misc, s_key, [&](int s, int f) {
                // the owner of point f publishes it as centroid s
                if ((f % NT) == tid) {
                    const int slot = f / NT;
#pragma unroll
                    for (int p = 0; p < PPT; ++p)
                        if (p == slot) s_node[s] = make_float4(px[p], py[p], pz[p], pn[p]);
                }
            });
            __syncthreads();
        } else {
            for (int j = tid; j < J; j += NT) {
                const float* h = hist_b + ((int64_t)resume * J + j) * 3;
                float x = h[0], y = h[1], z = h[2];
                s_node[j] = make_float4(x, y, z, __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
            }
        }
        for (int j = J + tid; j < Jp; j += NT) s_node[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncthreads();
    }

    // ---- outer iterations ------------------------------------------------------------------------------------
    for (int o = resume; o < iters; ++o) {
        if constexpr (kCluster) {
            float* h = P.hist + ((int64_t)b * iters + o) * J * 3;
            for (int j = tid; j < J; j += NT) { float4 c = s_node[j]; h[3 * j] = c.x; h[3 * j + 1] = c.y; h[3 * j + 2] = c.z; }
        }
        const int n_it = P.launch > 0 ? P.n_inner[o] : max_iter;
#pragma unroll
        for (int p = 0; p < PPT; ++p) u[p] = 0.f;
        for (int j = tid; j < Jp; j += NT) s_v[j] = 0.f;
        __syncthreads();

        for (int it = 0; it < n_it; ++it) {
            // ---- row update: u_i += eps * (log p_i - LSE_j K_ij) ------------------------------------------
            float du_abs = 0.f;
#pragma unroll
            for (int p = 0; p < PPT; ++p) {
                const int i = tid + p * NT;
                if (i < N) {
                    float m = -INFINITY, s = 0.f;
                    for (int j0 = 0; j0 < J; j0 += kJC) {
                        float x[kJC];
                        float mc = -INFINITY;
#pragma unroll
                        for (int jj = 0; jj < kJC; ++jj) {
                            const int j = j0 + jj;
                            if (j < J) {
                                x[jj] = __fadd_rn(__fadd_rn(-cost_at(p, j), u[p]), s_v[j]) * k2;
                                mc = fmaxf(mc, x[jj]);
                            } else x[jj] = -INFINITY;
                        }
                        const float mn = fmaxf(m, mc);
                        float acc = 0.f;
#pragma unroll
                        for (int jj = 0; jj < kJC; ++jj) acc += exp2f(x[jj] - mn);
                        s = s * exp2f(m - mn) + acc;        // m == -inf on the first chunk -> s*0
                        m = mn;
                    }
                    const float lse = (m + log2f(s)) * kLn2;
                    const float un = __fadd_rn(__fmul_rn(P.eps, logp[p] - lse), u[p]);
                    du_abs += fabsf(un - u[p]);
                    u[p] = un;
                }
            }
            // ---- column update: v_j += eps * (log q_j - LSE_i K_ij) -------------------------------------
            // After the row update every K_ij <= log(p_i + 1e-8) < 0, so the column sums need no
            // max shift (sum_i exp K_ij <= 1); a column whose sum underflows is redone exactly below.
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
                            if (j < J) part[jj] += exp2f(__fadd_rn(__fadd_rn(-cost_at(p, j), u[p]), s_v[j]) * k2);
                        }
                    }
                }
                const float tot = butterfly16(part, lane);
                if ((lane & 1) == 0) s_wtot[warp * Jp + j0 + ((lane >> 1) & 15)] = tot;
            }
            du_abs = warp_sum(du_abs);
            if (lane == 0) s_du[warp] = du_abs;
            __syncthreads();
            // final column sums (fixed order over warps) -> staged new v in s_wtot[j] (plane 0, warp-0 row,
            // which only this thread reads); NaN marks a column that must be redone with a max shift
            int need_exact = 0;
            for (int j = tid; j < J; j += NT) {
                float sj = 0.f;
#pragma unroll
                for (int w = 0; w < NW; ++w) sj += s_wtot[w * Jp + j];
                if (sj > 1e-30f && sj < INFINITY) {
                    s_wtot[j] = __fadd_rn(__fmul_rn(P.eps, s_logq[j] - logf(sj)), s_v[j]);
                } else {
                    s_wtot[j] = NAN;
                    need_exact = 1;
                }
            }
            if (__syncthreads_or(need_exact)) {
                // exact max-shifted LSE for the flagged columns (rare: a centroid nobody is near)
                for (int j = 0; j < J; ++j) {
                    if (s_wtot[j] == s_wtot[j]) continue;               // block-uniform (shared value)
                    float mx = -INFINITY;
#pragma unroll
                    for (int p = 0; p < PPT; ++p)
                        if (tid + p * NT < N) mx = fmaxf(mx, __fadd_rn(__fadd_rn(-cost_at(p, j), u[p]), s_v[j]) * k2);
                    mx = warp_max(mx);
                    if (lane == 0) s_tmp[warp] = mx;
                    __syncthreads();
                    float mall = -INFINITY;
                    for (int w = 0; w < NW; ++w) mall = fmaxf(mall, s_tmp[w]);
                    float sm = 0.f;
#pragma unroll
                    for (int p = 0; p < PPT; ++p)
                        if (tid + p * NT < N) sm += exp2f(__fadd_rn(__fadd_rn(-cost_at(p, j), u[p]), s_v[j]) * k2 - mall);
                    sm = block_sum<NT>(sm, s_red);
                    if (tid == 0) {
                        const float lse = (mall + log2f(sm)) * kLn2;
                        s_wtot[j] = __fadd_rn(__fmul_rn(P.eps, s_logq[j] - lse), s_v[j]);
                    }
                    __syncthreads();
                }
            }
            // commit v, accumulate sum |dv|, record this iteration's change
            float dv_abs = 0.f;
            for (int j = tid; j < J; j += NT) {
                const float vn = s_wtot[j];
                dv_abs += fabsf(vn - s_v[j]);
                s_v[j] = vn;
            }
            dv_abs = warp_sum(dv_abs);
            if (lane == 0) s_tmp[warp] = dv_abs;
            __syncthreads();
            if (tid == 0) {
                float du = 0.f, dv = 0.f;
#pragma unroll
                for (int w = 0; w < NW; ++w) { du += s_du[w]; dv += s_tmp[w]; }
                P.diffs[((int64_t)o * max_iter + it) * P.B + b] = du + dv;
            }
            __syncthreads();
        }

        // ---- gamma = exp(K); nan_to_num; row normalise; M-step on xyz -----------------------------------
        const bool last = (o == iters - 1);
        if constexpr (kCluster) {
            float rinv[PPT];
#pragma unroll
            for (int p = 0; p < PPT; ++p) {
                float rs = 0.f;
                if (tid + p * NT < N)
                    for (int j = 0; j < J; ++j)
                        rs += nan_to_num(exp2f(__fadd_rn(__fadd_rn(-cost_at(p, j), u[p]), s_v[j]) * k2), 0.f);
                rinv[p] = fmaxf(rs, 1e-3f);
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
                            const int j = j0 + jj;
                            g[jj] = 0.f;
                            if (j < J) {
                                float e = nan_to_num(exp2f(__fadd_rn(__fadd_rn(-cost_at(p, j), u[p]), s_v[j]) * k2), 0.f);
                                g[jj] = __fdiv_rn(e, rinv[p]);
                                a0[jj] += g[jj];
                                ax[jj] = fmaf(g[jj], px[p], ax[jj]);
                                ay[jj] = fmaf(g[jj], py[p], ay[jj]);
                                az[jj] = fmaf(g[jj], pz[p], az[jj]);
                            }
                        }
                        if (last) {
                            float* grow = P.gamma + ((int64_t)b * N + i) * J + j0;
                            if (((J & 3) == 0)) {
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
                    s_wtot[(0 * NW + warp) * Jp + col] = t0;
                    s_wtot[(1 * NW + warp) * Jp + col] = tx;
                    s_wtot[(2 * NW + warp) * Jp + col] = ty;
                    s_wtot[(3 * NW + warp) * Jp + col] = tz;
                }
            }
            __syncthreads();
            for (int j = tid; j < J; j += NT) {
                float s0 = 0.f, sx = 0.f, sy = 0.f, sz = 0.f;
#pragma unroll
                for (int w = 0; w < NW; ++w) {
                    s0 += s_wtot[(0 * NW + w) * Jp + j]; sx += s_wtot[(1 * NW + w) * Jp + j];
                    sy += s_wtot[(2 * NW + w) * Jp + j]; sz += s_wtot[(3 * NW + w) * Jp + j];
                }
                // lib/utils.py:137-140: pi = mean; npi = pi*N + 1e-5; mu = sum / npi
                const float pi = __fdiv_rn(s0, (float)N);
                const float npi = __fadd_rn(__fmul_rn(pi, (float)N), 1e-5f);
                const float mx = __fdiv_rn(sx, npi), my = __fdiv_rn(sy, npi), mz = __fdiv_rn(sz, npi);
                s_node[j] = make_float4(mx, my, mz, __fadd_rn(__fadd_rn(__fmul_rn(mx, mx), __fmul_rn(my, my)), __fmul_rn(mz, mz)));
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
                        const float e = exp2f(__fadd_rn(__fadd_rn(-c, u[p]), s_v[j]) * k2);
                        grow[j] = e;
                        lsum = fmaf(e, c, lsum);
                    }
                }
            }
            lsum = block_sum<NT>(lsum, s_red);
            if (tid == 0 && P.loss) P.loss[b] = lsum;
            __syncthreads();
        }
    }

    // ---- last CTA: evaluate the batch-mean exit test and publish the schedule ---------------------------
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const int prev = atomicAdd(&P.state[1], 1);
        s_misc[5] = (prev == P.B - 1) ? 1.f : 0.f;
    }
    __syncthreads();
    if (s_misc[5] == 0.f) return;
    __threadfence();
    const int total = (iters - resume) * max_iter;
    for (int e = warp; e < total; e += NW) {
        const int o = resume + e / max_iter, it = e % max_iter;
        const int n_it = P.launch > 0 ? P.n_inner[o] : max_iter;
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
            const int n_it = P.launch > 0 ? P.n_inner[o] : max_iter;
            int run = n_it;
            for (int it = 0; it < n_it; ++it)
                if (P.means[o * max_iter + it] < P.thresh) { run = it + 1; break; }
            P.n_inner[o] = run;
            if (run < n_it) { new_resume = o; break; }
        }
        // outer iterations after a shortened one restart with the full count
        for (int o = new_resume + 1; o < iters; ++o) P.n_inner[o] = max_iter;
        if (P.iters_run)
            for (int o = 0; o < iters; ++o) P.iters_run[o] = P.n_inner[o];
        P.state[0] = new_resume;
        P.state[1] = 0;
    }
}

}  // namespace ogmm

using namespace ogmm;

// ---- dispatch helpers --------------------------------------------------------------------------------------
// (threads per CTA, points per thread) by cloud size.  N > 8192 needs the multi-CTA variant (not built yet).
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
    for (int launch = 0; launch <= P.iters; ++launch) {
        P.launch = launch;
#define CALL(NT, PPT)                                                                                         \
    do {                                                                                                      \
        const size_t smem = sinkhorn_smem<NT>(P.J);                                                           \
        if (smem > 48 * 1024) {                                                                               \
            st = cuda_status(cudaFuncSetAttribute(sinkhorn_kernel<NT, PPT, kCluster>,                         \
                                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),    \
                             "cudaFuncSetAttribute(sinkhorn_kernel)");                                        \
            if (st != OGMM_OK) return st;                                                                     \
        }                                                                                                     \
        sinkhorn_kernel<NT, PPT, kCluster><<<(unsigned)P.B, NT, smem, s>>>(P);                                \
    } while (0)
        OGMM_DISPATCH_POINTS(P.N, CALL);
#undef CALL
        OGMM_LAUNCH_CHECK("sinkhorn_kernel");
    }
    return OGMM_OK;
}

