// K2 for large clouds (8192 < N <= 16384, any J <= 64): one CLOUD per THREAD-BLOCK CLUSTER of 16 CTAs (sm_100a).
//
// Same algorithm and arithmetic as the single-CTA log-domain path of sinkhorn_impl.cuh (FPS init lib/utils.py:170-198,
// cost = cdist :280, log-domain Sinkhorn :69-108, nan_to_num / row normalisation :282-287, xyz M-step :130-140, batch-
// coupled exit test :99-102 by record-and-verify with device-side redo rounds) -- but a cloud of 16384 points times 64
// components no longer fits one SM: 200 normalisation passes over 1 M entries per cloud took 15 ms in the
// register-starved 1024-thread variant this file replaces.  Here CTA r of the cluster owns points [1024 r, 1024 r + 1024)
// in the registers of its 512 threads, everything per point (row log-sum-exp, row normalisation, gamma) stays local,
// and everything per COLUMN -- the column sums of each Sinkhorn iteration, the M-step moments, the FPS arg-max, the
// change of the potentials -- is an all-reduce over the cluster through distributed shared memory:
//
//     every CTA stores its partial vector into slot [its rank] of EVERY CTA's exchange buffer (st.shared::cluster),
//     one barrier.cluster (arrive.release / wait.acquire), then each CTA adds the 16 slots in rank order.
//
// All CTAs therefore hold bit-identical totals (same values, same order), so the replicated per-column work (potential
// update, centroid update, exit bookkeeping) needs no broadcast and every data-dependent branch is cluster-uniform.
// The exchange buffer is double buffered by call parity: a CTA can run at most one all-reduce ahead of the slowest CTA,
// so one cluster barrier per all-reduce is enough.  Cost per cloud: ~100 x (one sweep over 1024 x 64 entries per CTA,
// MUFU bound: sqrt + exp2 per entry; the column sums reuse the row pass's exponentials) + ~270 cluster barriers of ~0.2 us.
#include <cooperative_groups.h>

#include "sinkhorn_impl.cuh"

namespace cg = cooperative_groups;

namespace ogmm {

constexpr int kCS = 16;                       // CTAs per cluster (non-portable size; one GPC holds 16..20 SMs)
constexpr int kDNT = 512;                     // threads per CTA
constexpr int kDPPT = 2;                      // points per thread
constexpr int kDChunk = kDNT * kDPPT;         // points per CTA
constexpr int kDMaxJ = 64;
constexpr int kXW = 4 * kDMaxJ + 8;           // floats per exchange slot (M-step: 4 moments x J; iteration: J + 1)

struct DSmem {
    float4* node;      // [64] centroid (x, y, z, |n|^2), replicated in every CTA
    float* v;          // [64] column potential
    float* logq;       // [64]
    float* wtot;       // [4][NW][64] per-warp partials
    float* loc;        // [kXW] this CTA's partial vector
    float* tot;        // [kXW] cluster totals (identical in every CTA)
    float* slots;      // [2][kCS][kXW] exchange buffer
    float* red;        // [32]
    float* misc;       // [16]
    unsigned long long* key;   // [32]
    float4* cost;      // [kDCacheJ / 4][kDChunk] costs of the first kDCacheJ columns for this CTA's points (per outer iteration)
};
constexpr int kDCacheJ = 40;                  // columns whose cost c_ij is cached in shared memory (160 KB at 1024 points; 54 KB are taken by the exchange buffers)
constexpr size_t kDSmemSmall = sizeof(float4) * kDMaxJ + sizeof(float) * (2 * kDMaxJ + 4 * (kDNT / 32) * kDMaxJ + 2 * kXW +
                                                                           2 * kCS * kXW + 32 + 16) + sizeof(unsigned long long) * 32;
constexpr size_t kDSmemBytes = ((kDSmemSmall + 15) & ~(size_t)15) + sizeof(float4) * (kDCacheJ / 4) * kDChunk;

__device__ __forceinline__ DSmem carve_dsmem(unsigned char* raw) {
    DSmem s;
    s.node = reinterpret_cast<float4*>(raw);
    s.v = reinterpret_cast<float*>(s.node + kDMaxJ);
    s.logq = s.v + kDMaxJ;
    s.wtot = s.logq + kDMaxJ;
    s.loc = s.wtot + 4 * (kDNT / 32) * kDMaxJ;
    s.tot = s.loc + kXW;
    s.slots = s.tot + kXW;
    s.red = s.slots + 2 * kCS * kXW;
    s.misc = s.red + 32;
    s.key = reinterpret_cast<unsigned long long*>(s.misc + 16);
    s.cost = reinterpret_cast<float4*>(raw + ((kDSmemSmall + 15) & ~(size_t)15));
    return s;
}

// Sum (or max) of the CTAs' partial vectors S.loc[0..n) -> S.tot[0..n), identical in every CTA.  Collective over the
// whole cluster; `parity` alternates per call.
template <bool kMax>
__device__ __forceinline__ void cluster_allreduce(cg::cluster_group& cl, const DSmem& S, int& parity, int n) {
    const unsigned rank = cl.block_rank();
    __syncthreads();                                           // S.loc is complete
    float* mine = S.slots + ((size_t)parity * kCS + rank) * kXW;
    for (int e = threadIdx.x; e < n * kCS; e += kDNT) {
        const int r = e / n, t = e - r * n;                     // consecutive threads -> consecutive floats of one peer
        cl.map_shared_rank(mine, r)[t] = S.loc[t];
    }
    cl.sync();
    for (int t = threadIdx.x; t < n; t += kDNT) {
        const float* base = S.slots + (size_t)parity * kCS * kXW + t;
        float acc = base[0];
#pragma unroll
        for (int r = 1; r < kCS; ++r) acc = kMax ? fmaxf(acc, base[(size_t)r * kXW]) : acc + base[(size_t)r * kXW];
        S.tot[t] = acc;
    }
    parity ^= 1;
    __syncthreads();
}

// Block-wide fold of per-warp column partials wtot[which][warp][j] -> S.loc[off + j] (fixed order over warps).
__device__ __forceinline__ void fold_warps(const DSmem& S, int which, int J, int off) {
    constexpr int NW = kDNT / 32;
    __syncthreads();
    for (int j = threadIdx.x; j < J; j += kDNT) {
        float acc = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) acc += S.wtot[((size_t)which * NW + w) * kDMaxJ + j];
        S.loc[off + j] = acc;
    }
}

__global__ void __cluster_dims__(kCS, 1, 1) __launch_bounds__(kDNT, 1)
sinkhorn_cluster_dsmem_kernel(SinkhornParams P, int mode) {
    extern __shared__ __align__(16) unsigned char dsm_raw[];
    cg::cluster_group cl = cg::this_cluster();
    const DSmem S = carve_dsmem(dsm_raw);
    constexpr int NW = kDNT / 32;
    const int rank = (int)cl.block_rank();
    const int b = blockIdx.x / kCS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = P.N, J = P.J, iters = P.iters, max_iter = P.max_iter;
    const float k2 = (1.0f / P.eps) * kLog2e;
    const bool unit_tau = P.tau == 1.0f;
    const bool first = mode == 0;
    int resume = 0;
    if (!first) {
        resume = __ldcg(P.state);
        if (resume >= iters) return;                           // cluster-uniform
    }
    int parity = 0;

    // ---- this CTA's points ------------------------------------------------------------------------------------------
    const float* base = P.xyz + (int64_t)b * P.sb;
    float px[kDPPT], py[kDPPT], pz[kDPPT], m2x[kDPPT], m2y[kDPPT], m2z[kDPPT], pn[kDPPT], logp[kDPPT], u[kDPPT];
    bool live[kDPPT];
    float osum = 0.f;
#pragma unroll
    for (int p = 0; p < kDPPT; ++p) {
        const int i = rank * kDChunk + tid + p * kDNT;
        live[p] = i < N;
        px[p] = py[p] = pz[p] = 0.f; logp[p] = 0.f;
        if (live[p]) {
            px[p] = base[(int64_t)i * P.sn]; py[p] = base[(int64_t)i * P.sn + P.sc]; pz[p] = base[(int64_t)i * P.sn + 2 * P.sc];
            logp[p] = P.o_scores[(int64_t)b * N + i];
            osum += logp[p];
        }
        m2x[p] = -2.f * px[p]; m2y[p] = -2.f * py[p]; m2z[p] = -2.f * pz[p];
        pn[p] = sq3(px[p], py[p], pz[p]);
    }
    {
        float sx = 0.f, sy = 0.f, sz = 0.f;
#pragma unroll
        for (int p = 0; p < kDPPT; ++p)
            if (live[p]) { sx += px[p]; sy += py[p]; sz += pz[p]; }
        const float a0 = block_sum<kDNT>(osum, S.red), a1 = block_sum<kDNT>(sx, S.red), a2 = block_sum<kDNT>(sy, S.red),
                    a3 = block_sum<kDNT>(sz, S.red);
        if (tid == 0) { S.loc[0] = a0; S.loc[1] = a1; S.loc[2] = a2; S.loc[3] = a3; }
        cluster_allreduce<false>(cl, S, parity, 4);
    }
    {
        const float otot = fmaxf(S.tot[0], 1e-4f);              // lib/utils.py:276  o / clip(sum o, 1e-4)
#pragma unroll
        for (int p = 0; p < kDPPT; ++p) logp[p] = logf(__fdiv_rn(logp[p], otot) + 1e-8f);
    }
    for (int j = tid; j < kDMaxJ; j += kDNT) S.logq[j] = logf(1.0f / (float)J + 1e-8f);
    auto cost_at = [&](int p, int j) -> float { return node_cost(m2x[p], m2y[p], m2z[p], pn[p], S.node[j], P.tau); };

    // ---- initial centroids -------------------------------------------------------------------------------------------
    if (resume == 0) {
        float best[kDPPT];
#pragma unroll
        for (int p = 0; p < kDPPT; ++p) best[p] = 1e10f;
        float cx = S.tot[1] / (float)N, cy = S.tot[2] / (float)N, cz = S.tot[3] / (float)N;   // is_center start (:183-188)
        for (int s = -1; s < J - 1; ++s) {
            // relax against the current centre, then the cluster-wide arg-max (lowest index on ties)
            unsigned long long key = 0ull;
            float bx = 0.f, by = 0.f, bz = 0.f;
#pragma unroll
            for (int p = 0; p < kDPPT; ++p) {
                if (live[p]) {
                    const float d = sq3(px[p] - cx, py[p] - cy, pz[p] - cz);
                    if (d < best[p]) best[p] = d;
                    const unsigned long long kk = far_key(best[p], rank * kDChunk + tid + p * kDNT);
                    if (kk > key) { key = kk; bx = px[p]; by = py[p]; bz = pz[p]; }
                }
            }
            // block arg-max carrying the coordinates
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned long long ok = __shfl_xor_sync(kFull, key, o);
                const float ox = __shfl_xor_sync(kFull, bx, o), oy = __shfl_xor_sync(kFull, by, o), oz = __shfl_xor_sync(kFull, bz, o);
                if (ok > key) { key = ok; bx = ox; by = oy; bz = oz; }
            }
            __syncthreads();
            if (lane == 0) { S.key[warp] = key; S.wtot[warp * 4] = bx; S.wtot[warp * 4 + 1] = by; S.wtot[warp * 4 + 2] = bz; }
            __syncthreads();
            if (tid == 0) {
                unsigned long long kb = 0ull;
                int wb = 0;
                for (int w = 0; w < NW; ++w) if (S.key[w] > kb) { kb = S.key[w]; wb = w; }
                // exchange as floats: the two halves of the key are compared lexicographically after the sums below
                S.loc[0] = __uint_as_float((unsigned)(kb >> 32));        // best distance bits (non-negative float)
                S.loc[1] = __uint_as_float((unsigned)(kb & 0xffffffffull));
                S.loc[2] = S.wtot[wb * 4]; S.loc[3] = S.wtot[wb * 4 + 1]; S.loc[4] = S.wtot[wb * 4 + 2];
            }
            // all-gather of the 16 candidates (the "sum" output is not used: every CTA scans the slots itself)
            {
                const unsigned r0 = cl.block_rank();
                __syncthreads();
                float* mine = S.slots + ((size_t)parity * kCS + r0) * kXW;
                if (tid < 5 * kCS) { const int r = tid / 5, t = tid - r * 5; cl.map_shared_rank(mine, r)[t] = S.loc[t]; }
                cl.sync();
                if (tid == 0) {
                    unsigned long long kb = 0ull;
                    int rb = 0;
                    for (int r = 0; r < kCS; ++r) {
                        const float* sl = S.slots + ((size_t)parity * kCS + r) * kXW;
                        const unsigned long long kk = ((unsigned long long)__float_as_uint(sl[0]) << 32) | __float_as_uint(sl[1]);
                        if (kk > kb) { kb = kk; rb = r; }
                    }
                    const float* sl = S.slots + ((size_t)parity * kCS + rb) * kXW;
                    S.misc[0] = sl[2]; S.misc[1] = sl[3]; S.misc[2] = sl[4];
                }
                parity ^= 1;
                __syncthreads();
            }
            cx = S.misc[0]; cy = S.misc[1]; cz = S.misc[2];
            // round s == -1 relaxes against the centroid and its winner is the START point (:183-188); a pick is emitted
            // BEFORE the relaxation against it (:191-197), so the winner of round s is centroid s + 1
            if (tid == 0) S.node[s + 1] = make_float4(cx, cy, cz, sq3(cx, cy, cz));
            __syncthreads();
        }
    } else {
        const float* h = P.hist + ((int64_t)b * iters + resume) * J * 3;
        for (int j = tid; j < J; j += kDNT) {
            const float x = h[3 * j], y = h[3 * j + 1], z = h[3 * j + 2];
            S.node[j] = make_float4(x, y, z, sq3(x, y, z));
        }
    }
    for (int j = J + tid; j < kDMaxJ; j += kDNT) S.node[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();

    // ---- outer iterations ------------------------------------------------------------------------------------------------
    for (int o = resume; o < iters; ++o) {
        if (rank == 0) {
            float* h = P.hist + ((int64_t)b * iters + o) * J * 3;
            for (int j = tid; j < J; j += kDNT) { const float4 c = S.node[j]; h[3 * j] = c.x; h[3 * j + 1] = c.y; h[3 * j + 2] = c.z; }
        }
        const int n_it = first ? max_iter : __ldcg(P.n_inner + o);
        const bool last = (o == iters - 1);
#pragma unroll
        for (int p = 0; p < kDPPT; ++p) u[p] = 0.f;
        __syncthreads();
        for (int j = tid; j < kDMaxJ; j += kDNT) S.v[j] = 0.f;
        // The centroids are fixed for the n_it iterations of this call: the cost of the first kDCacheJ columns is
        // computed once and kept in shared memory as [column quad][point] float4 (consecutive lanes -> consecutive 16 B,
        // conflict free; each thread reads back only what it wrote), the remaining columns are recomputed per iteration.
#pragma unroll
        for (int p = 0; p < kDPPT; ++p) {
#pragma unroll
            for (int q = 0; q < kDCacheJ / 4; ++q) {
                if (4 * q < J) {
                    float4 c4;
                    c4.x = cost_at(p, 4 * q); c4.y = cost_at(p, 4 * q + 1); c4.z = cost_at(p, 4 * q + 2); c4.w = cost_at(p, 4 * q + 3);
                    S.cost[(size_t)q * kDChunk + tid + p * kDNT] = c4;
                }
            }
        }
        __syncthreads();

        for (int it = 0; it < n_it; ++it) {
            // ---- row update u_i += eps (log p_i - LSE_j K_ij) and the column sums of the UPDATED kernel in ONE sweep per
            // point.  With x_ij = (-c_ij + u_i + v_j) k2, m_i = max_j x_ij, e_ij = exp2(x_ij - m_i), s_i = sum_j e_ij the row
            // update is u_i' = u_i + eps (log p_i - (m_i + log2 s_i) ln 2), and the column pass needs
            //     exp2((-c_ij + u_i' + v_j) k2) = e_ij * exp2(m_i + (u_i' - u_i) k2)          (= e_ij p_i / s_i)
            // i.e. the e_ij the row pass already holds times ONE more exp2 per point -- sqrt + exp2 once per entry and
            // iteration instead of twice (the kernel is MUFU bound).  The 64 e_ij of the point stay in registers; the
            // points of a thread are processed one after the other and their butterfly outputs add up.
            float du_abs = 0.f;
            float tot4[kDMaxJ / kJC];
#pragma unroll
            for (int c = 0; c < kDMaxJ / kJC; ++c) tot4[c] = 0.f;
#pragma unroll 1
            for (int p = 0; p < kDPPT; ++p) {
                float e[kDMaxJ];
                const bool lv = p == 0 ? live[0] : live[1];
                const float up = p == 0 ? u[0] : u[1], lp = p == 0 ? logp[0] : logp[1];
                const float ax = p == 0 ? m2x[0] : m2x[1], ay = p == 0 ? m2y[0] : m2y[1], az = p == 0 ? m2z[0] : m2z[1];
                const float an = p == 0 ? pn[0] : pn[1];
                float un = up;
                if (lv) {
                    float m = -INFINITY;
#pragma unroll
                    for (int q = 0; q < kDCacheJ / 4; ++q) {
                        float4 c4 = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (4 * q < J) c4 = S.cost[(size_t)q * kDChunk + tid + p * kDNT];
                        const float cq[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            const int j = 4 * q + t;
                            e[j] = -INFINITY;
                            if (j < J) { e[j] = __fadd_rn(__fadd_rn(-cq[t], up), S.v[j]) * k2; m = fmaxf(m, e[j]); }
                        }
                    }
                    if (unit_tau) {                                     // cluster-uniform: tau tested once, not per entry
#pragma unroll
                        for (int j = kDCacheJ; j < kDMaxJ; ++j) {
                            e[j] = -INFINITY;
                            if (j < J) { e[j] = __fadd_rn(__fadd_rn(-node_dist(ax, ay, az, an, S.node[j]), up), S.v[j]) * k2; m = fmaxf(m, e[j]); }
                        }
                    } else {
#pragma unroll
                        for (int j = kDCacheJ; j < kDMaxJ; ++j) {
                            e[j] = -INFINITY;
                            if (j < J) { e[j] = __fadd_rn(__fadd_rn(-__fdiv_rn(node_dist(ax, ay, az, an, S.node[j]), P.tau), up), S.v[j]) * k2; m = fmaxf(m, e[j]); }
                        }
                    }
                    float sum = 0.f;
#pragma unroll
                    for (int j = 0; j < kDMaxJ; ++j) { e[j] = fast_exp2(e[j] - m); sum += e[j]; }      // MUFU.EX2, ftz: no range fix-up code per entry
                    const float lse = (m + log2f(sum)) * kLn2;
                    un = __fadd_rn(__fmul_rn(P.eps, lp - lse), up);
                    du_abs += fabsf(un - up);
                    const float w = exp2f(fmaf(un - up, k2, m));
#pragma unroll
                    for (int j = 0; j < kDMaxJ; ++j) e[j] *= w;
                } else {
#pragma unroll
                    for (int j = 0; j < kDMaxJ; ++j) e[j] = 0.f;
                }
                if (p == 0) u[0] = un; else u[1] = un;
#pragma unroll
                for (int c = 0; c < kDMaxJ / kJC; ++c) {
                    if (c * kJC < J) {                                    // cluster-uniform
                        float part[kJC];
#pragma unroll
                        for (int jj = 0; jj < kJC; ++jj) part[jj] = e[c * kJC + jj];
                        tot4[c] += butterfly16(part, lane);
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < kDMaxJ / kJC; ++c)
                if (c * kJC < J && (lane & 1) == 0) S.wtot[(size_t)warp * kDMaxJ + c * kJC + ((lane >> 1) & 15)] = tot4[c];
            du_abs = block_sum<kDNT>(du_abs, S.red);              // also orders the wtot writes before the fold
            fold_warps(S, 0, J, 0);
            if (tid == 0) S.loc[J] = du_abs;
            cluster_allreduce<false>(cl, S, parity, J + 1);
            // ---- potential update, replicated in every CTA on identical totals
            int need_exact = 0;
            for (int j = tid; j < J; j += kDNT) {
                const float sj = S.tot[j];
                if (sj > 1e-30f && sj < INFINITY) S.wtot[j] = __fadd_rn(__fmul_rn(P.eps, S.logq[j] - logf(sj)), S.v[j]);
                else { S.wtot[j] = NAN; need_exact = 1; }
            }
            const float du_total = S.tot[J];
            if (__syncthreads_or(need_exact)) {                   // cluster-uniform: every CTA sees the same totals
                for (int j = 0; j < J; ++j) {
                    if (S.wtot[j] == S.wtot[j]) continue;
                    float mx = -INFINITY;
#pragma unroll
                    for (int p = 0; p < kDPPT; ++p)
                        if (live[p]) mx = fmaxf(mx, __fadd_rn(__fadd_rn(-cost_at(p, j), u[p]), S.v[j]) * k2);
                    mx = warp_max(mx);
                    if (lane == 0) S.red[warp] = mx;
                    __syncthreads();
                    if (tid == 0) { float mall = -INFINITY; for (int w = 0; w < NW; ++w) mall = fmaxf(mall, S.red[w]); S.loc[0] = mall; }
                    cluster_allreduce<true>(cl, S, parity, 1);
                    const float mall = S.tot[0];
                    float sm = 0.f;
#pragma unroll
                    for (int p = 0; p < kDPPT; ++p)
                        if (live[p]) sm += exp2f(__fadd_rn(__fadd_rn(-cost_at(p, j), u[p]), S.v[j]) * k2 - mall);
                    sm = block_sum<kDNT>(sm, S.red);
                    if (tid == 0) S.loc[0] = sm;
                    cluster_allreduce<false>(cl, S, parity, 1);
                    if (tid == 0) {
                        const float lse = (mall + log2f(S.tot[0])) * kLn2;
                        S.wtot[j] = __fadd_rn(__fmul_rn(P.eps, S.logq[j] - lse), S.v[j]);
                    }
                    __syncthreads();
                }
            }
            float dv_abs = 0.f;
            for (int j = tid; j < J; j += kDNT) {
                const float vn = S.wtot[j];
                dv_abs += fabsf(vn - S.v[j]);
                S.v[j] = vn;
            }
            dv_abs = block_sum<kDNT>(dv_abs, S.red);
            if (rank == 0 && tid == 0) P.diffs[((int64_t)o * max_iter + it) * P.B + b] = du_total + dv_abs;
            __syncthreads();
        }

        // ---- gamma = exp(K); nan_to_num; row normalise; M-step on xyz (column moments through the cluster) --------------
        float rinv[kDPPT];
#pragma unroll
        for (int p = 0; p < kDPPT; ++p) {
            float rs = 0.f;
            if (live[p])
                for (int j = 0; j < J; ++j) rs += nan_to_num(exp2f(__fadd_rn(__fadd_rn(-cost_at(p, j), u[p]), S.v[j]) * k2), 0.f);
            rinv[p] = __frcp_rn(fmaxf(rs, 1e-3f));
        }
        for (int j0 = 0; j0 < J; j0 += kJC) {
            float a0[kJC], ax[kJC], ay[kJC], az[kJC];
#pragma unroll
            for (int jj = 0; jj < kJC; ++jj) a0[jj] = ax[jj] = ay[jj] = az[jj] = 0.f;
#pragma unroll
            for (int p = 0; p < kDPPT; ++p) {
                if (live[p]) {
                    float g[kJC];
#pragma unroll
                    for (int jj = 0; jj < kJC; ++jj) {
                        g[jj] = 0.f;
                        if (j0 + jj < J) {
                            g[jj] = nan_to_num(exp2f(__fadd_rn(__fadd_rn(-cost_at(p, j0 + jj), u[p]), S.v[j0 + jj]) * k2), 0.f) * rinv[p];
                            a0[jj] += g[jj];
                            ax[jj] = fmaf(g[jj], px[p], ax[jj]); ay[jj] = fmaf(g[jj], py[p], ay[jj]); az[jj] = fmaf(g[jj], pz[p], az[jj]);
                        }
                    }
                    if (last) {
                        float* grow = P.gamma + ((int64_t)b * N + rank * kDChunk + tid + p * kDNT) * J + j0;
                        if ((J & 3) == 0) {
#pragma unroll
                            for (int jj = 0; jj < kJC; jj += 4)
                                if (j0 + jj < J) *reinterpret_cast<float4*>(grow + jj) = make_float4(g[jj], g[jj + 1], g[jj + 2], g[jj + 3]);
                        } else {
#pragma unroll
                            for (int jj = 0; jj < kJC; ++jj)
                                if (j0 + jj < J) grow[jj] = g[jj];
                        }
                    }
                }
            }
            const float t0 = butterfly16(a0, lane), tx = butterfly16(ax, lane), ty = butterfly16(ay, lane), tz = butterfly16(az, lane);
            if ((lane & 1) == 0) {
                const int col = j0 + ((lane >> 1) & 15);
                S.wtot[((size_t)0 * NW + warp) * kDMaxJ + col] = t0; S.wtot[((size_t)1 * NW + warp) * kDMaxJ + col] = tx;
                S.wtot[((size_t)2 * NW + warp) * kDMaxJ + col] = ty; S.wtot[((size_t)3 * NW + warp) * kDMaxJ + col] = tz;
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) fold_warps(S, q, J, q * J);
        cluster_allreduce<false>(cl, S, parity, 4 * J);
        for (int j = tid; j < J; j += kDNT) {
            // lib/utils.py:137-140: pi = mean; npi = pi N + 1e-5; mu = sum / npi
            const float pi = __fdiv_rn(S.tot[j], (float)N);
            const float npi = __fadd_rn(__fmul_rn(pi, (float)N), 1e-5f);
            const float mx = __fdiv_rn(S.tot[J + j], npi), my = __fdiv_rn(S.tot[2 * J + j], npi), mz = __fdiv_rn(S.tot[3 * J + j], npi);
            S.node[j] = make_float4(mx, my, mz, sq3(mx, my, mz));
            if (last && rank == 0) {
                P.pi[(int64_t)b * J + j] = pi;
                float* m = P.mu + ((int64_t)b * J + j) * 3;
                m[0] = mx; m[1] = my; m[2] = mz;
            }
        }
        __syncthreads();
    }

    // ---- batch-mean exit test by the last cluster to finish; redo rounds are tail-launched (sinkhorn_impl.cuh) ----------
    __threadfence();
    cl.sync();                                                     // every CTA of this cloud is done
    if (rank != 0) return;
    if (tid == 0) {
        const int prev = atomicAdd(&P.state[1], 1);
        S.misc[5] = (prev == P.B - 1) ? 1.f : 0.f;
    }
    __syncthreads();
    if (S.misc[5] != 0.f) {
        __threadfence();
        verify_schedule<kDNT>(P, resume, first);
        if (tid == 0 && *reinterpret_cast<volatile int32_t*>(P.state) < iters) {
            sinkhorn_cluster_dsmem_kernel<<<gridDim.x, kDNT, kDSmemBytes, cudaStreamTailLaunch>>>(P, 1);
            if (cudaGetLastError() != cudaSuccess) atomicExch(&P.state[5], 1);
        }
    }
}

}  // namespace ogmm

using namespace ogmm;

// Clustering of clouds with 8193..16384 points and J <= 64 (BASELINE.json configs[3]); called by launch_sinkhorn.
int ogmm_launch_cluster_dsmem(ogmm::SinkhornParams P, cudaStream_t s) {
    OGMM_REQUIRE(P.J <= kDMaxJ, OGMM_EUNSUPPORTED, "sinkhorn_cluster: N=%d > 8192 needs J <= %d (got %d)", P.N, kDMaxJ, P.J);
    OGMM_REQUIRE(P.N <= kCS * kDChunk, OGMM_EUNSUPPORTED, "sinkhorn_cluster: N=%d > %d", P.N, kCS * kDChunk);
    int st = cuda_status(cudaFuncSetAttribute(sinkhorn_cluster_dsmem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDSmemBytes),
                         "cudaFuncSetAttribute(sinkhorn_cluster_dsmem_kernel, smem)");
    if (st != OGMM_OK) return st;
    st = cuda_status(cudaFuncSetAttribute(sinkhorn_cluster_dsmem_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1),
                     "cudaFuncSetAttribute(sinkhorn_cluster_dsmem_kernel, cluster size 16)");
    if (st != OGMM_OK) return st;
    sinkhorn_cluster_dsmem_kernel<<<(unsigned)(P.B * kCS), kDNT, kDSmemBytes, s>>>(P, 0);
    return cuda_status(cudaGetLastError(), "sinkhorn_cluster_dsmem_kernel");
}
