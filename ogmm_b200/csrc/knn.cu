// K1: fused pairwise distance + per-row top-k (+ edge-feature gather) (sm_100a).
//
//   knn3_kernel        lib/utils.py:12-44 (+ :47-66 when edge_out is given) for C == 3: one thread per
//                      query, candidates broadcast from shared memory, FP32 FMA in the reference's
//                      expanded form  ((-2 s.d) + |s|^2) + |d|^2, clamp 1e-12  -- never the (B,N,M) matrix.
//   knn_generic_kernel same contract for any C: a 128 x 64 distance tile per CTA step (shared-memory
//                      tiled FP32 FMA), consumed straight out of shared memory by the same selector.
//   edge_gather_kernel lib/utils.py:55-66 alone, for a caller that already holds idx.
//
// Selection: each query thread keeps its K best (distance, index) pairs sorted in registers.  A
// candidate that beats the current K-th best is first parked in a small per-thread staging column in
// shared memory; when any lane of the warp runs low on staging space the whole warp merges its parked
// candidates into the register lists together (converged, no per-candidate divergence).  Candidates
// are visited in index order and every comparison is strict, so equal distances resolve to the LOWEST
// index -- the documented tie-break (torch.topk's own is unspecified).
#include <stdlib.h>

#include "knn_common.cuh"

namespace ogmm {

constexpr int kKnnThreads = 256;
constexpr int kStage = 12;          // staging slots per query thread
constexpr int kStageTrigger = 8;    // merge when a lane holds more than this many

template <int K>
struct TopK {
    float d[K];
    int i[K];
    float thr;
    int cnt;
    bool live;      // false for padding threads: they take part in the warp votes but never keep anything
    float* sd;      // staging columns: sd[s * stride + col]
    int* si;
    int stride, col;

    __device__ __forceinline__ void init(float* stage_d, int* stage_i, int stride_, int col_, bool live_) {
#pragma unroll
        for (int j = 0; j < K; ++j) { d[j] = INFINITY; i[j] = 0; }
        live = live_;
        thr = live ? INFINITY : -INFINITY; cnt = 0; sd = stage_d; si = stage_i; stride = stride_; col = col_;
    }
    __device__ __forceinline__ void offer(float v, int idx) {
        if (v < thr) {
            sd[cnt * stride + col] = v;
            si[cnt * stride + col] = idx;
            ++cnt;
        }
    }
    // Sorted insert.  Up to the insertion point the comparison is strict (an equal, earlier entry stays in
    // front); from there on every entry shifts down by one, so ties keep their arrival (= index) order.
    __device__ __forceinline__ void insert(float v, int idx) {
        bool moved = false;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            moved = moved || (v < d[j]);
            const float td = d[j];
            const int ti = i[j];
            d[j] = moved ? v : td;
            i[j] = moved ? idx : ti;
            v = moved ? td : v;
            idx = moved ? ti : idx;
        }
    }
    // Warp-converged: every lane of the calling warp must call merge() together.
    __device__ __forceinline__ void merge() {
        const int most = __reduce_max_sync(kFull, cnt);
        for (int s = 0; s < most; ++s) {
            float v = INFINITY;
            int idx = 0;
            if (s < cnt) { v = sd[s * stride + col]; idx = si[s * stride + col]; }
            if (__any_sync(kFull, v < d[K - 1])) insert(v, idx);
        }
        cnt = 0;
        thr = live ? d[K - 1] : -INFINITY;
    }
    __device__ __forceinline__ void maybe_merge() {
        if (__any_sync(kFull, cnt > kStageTrigger)) merge();
    }
};

__device__ __forceinline__ float sq_norm3(float x, float y, float z) {
    return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
}

// ---------------------------------------------------------------------------------------------------
constexpr int kCandTile = 1024;

template <int K>
__global__ void __launch_bounds__(kKnnThreads)
knn3_kernel(const float* __restrict__ src, int64_t s_sb, int64_t s_sn, int64_t s_sc,
            const float* __restrict__ dst, int64_t d_sb, int64_t d_sn, int64_t d_sc,
            int N, int M, int k, int normalize,
            int64_t* __restrict__ idx_out, float* __restrict__ dist_out, float* __restrict__ edge_out) {
    // candidate tile as packed group records (knn_common.cuh: group_distances): 64 bytes per four candidates,
    // (xA,xB,yA,yB) (zA,zB,wA,wB) (xC,xD,yC,yD) (zC,zD,wC,wD), coordinates pre-scaled by -2 -- two candidates per packed
    // FP32x2 instruction, each half the same IEEE operation as the scalar form
    __shared__ __align__(16) float s_rec[4 * kCandTile];
    __shared__ float s_stage_d[kStage * kKnnThreads];
    __shared__ int s_stage_i[kStage * kKnnThreads];

    const int b = blockIdx.y, tid = threadIdx.x;
    const int q = blockIdx.x * kKnnThreads + tid;
    const bool valid = q < N;
    const float* sb = src + (int64_t)b * s_sb;
    const float* db = dst + (int64_t)b * d_sb;

    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (valid) { qx = sb[(int64_t)q * s_sn]; qy = sb[(int64_t)q * s_sn + s_sc]; qz = sb[(int64_t)q * s_sn + 2 * s_sc]; }
    const float qs = normalize ? 2.0f : sq_norm3(qx, qy, qz);
    QueryPack Q;
    Q.x = pack2(qx, qx); Q.y = pack2(qy, qy); Q.z = pack2(qz, qz); Q.s = pack2(qs, qs);
    const unsigned rec_base = (unsigned)__cvta_generic_to_shared(s_rec);

    TopK<K> top;
    top.init(s_stage_d, s_stage_i, kKnnThreads, tid, valid);

    for (int m0 = 0; m0 < M; m0 += kCandTile) {
        const int len = min(kCandTile, M - m0);
        const int len8 = (len + 7) & ~7;
        __syncthreads();
        for (int m = tid; m < len8; m += kKnnThreads) {
            float x = 0.f, y = 0.f, z = 0.f, w = INFINITY;             // padding: distance +inf, never offered
            if (m < len) {
                const float* p = db + (int64_t)(m0 + m) * d_sn;
                const float cx = p[0], cy = p[d_sc], cz = p[2 * d_sc];
                x = -2.f * cx; y = -2.f * cy; z = -2.f * cz; w = normalize ? 0.f : sq_norm3(cx, cy, cz);
            }
            float* r = s_rec + 16 * (m >> 2) + 8 * ((m & 3) >> 1) + (m & 1);
            r[0] = x; r[2] = y; r[4] = z; r[6] = w;
        }
        __syncthreads();
        // eight candidates (two group records) per step; (-2 s.d) accumulated like a K=3 GEMM, then + |s|^2, then + |d|^2
        // (lib/utils.py:28-32).  Tiles are padded to a multiple of eight with +inf records.
        for (int m = 0; m < len8; m += 8) {
            float v[4], w[4];
            group_distances(rec_base + 16u * (unsigned)m, Q, v);
            group_distances(rec_base + 16u * (unsigned)m + 64u, Q, w);
            if (!normalize) {
#pragma unroll
                for (int u = 0; u < 4; ++u) { v[u] = fmaxf(v[u], 1e-12f); w[u] = fmaxf(w[u], 1e-12f); }
            }
            // one test per step: after the first tiles a candidate beats the k-th best about once in a hundred
            // (warp vote: the staged count only changes inside, so the merge tests live there too)
            const float mn = fminf(fminf(fminf(v[0], v[1]), fminf(v[2], v[3])), fminf(fminf(w[0], w[1]), fminf(w[2], w[3])));
            if (__any_sync(kFull, mn < top.thr)) {
#pragma unroll
                for (int u = 0; u < 4; ++u) top.offer(v[u], m0 + m + u);
                top.maybe_merge();
#pragma unroll
                for (int u = 0; u < 4; ++u) top.offer(w[u], m0 + m + 4 + u);
                top.maybe_merge();
            }
        }
    }
    top.merge();

    if (!valid) return;
    int64_t* io = idx_out + ((int64_t)b * N + q) * k;
    float* dout = dist_out ? dist_out + ((int64_t)b * N + q) * k : nullptr;
    float* eo = edge_out ? edge_out + ((int64_t)b * N + q) * (int64_t)k * 6 : nullptr;
#pragma unroll
    for (int j = 0; j < K; ++j) {
        if (j < k) {
            io[j] = top.i[j];                                          // unfilled slots (NaN distances) hold index 0
            if (dout) dout[j] = top.d[j];
            if (eo) {
                const float* p = sb + (int64_t)top.i[j] * s_sn;       // self graph: neighbours live in src
                eo[6 * j + 0] = p[0] - qx; eo[6 * j + 1] = p[s_sc] - qy; eo[6 * j + 2] = p[2 * s_sc] - qz;
                eo[6 * j + 3] = qx; eo[6 * j + 4] = qy; eo[6 * j + 5] = qz;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// generic C: CTA tile = 128 queries x 64 candidates, 256 threads, 8 x 4 register tile per thread.
constexpr int kGQ = 128, kGC = 64, kGK = 16;
constexpr size_t kGenericSmem = sizeof(float) * (kGK * (kGQ + 4) + kGK * (kGC + 4) + kGQ * (kGC + 1) + kGQ + kGC + 2 * kStage * kGQ);

template <int K>
__global__ void __launch_bounds__(kKnnThreads)
knn_generic_kernel(const float* __restrict__ src, int64_t s_sb, int64_t s_sn, int64_t s_sc,
                   const float* __restrict__ dst, int64_t d_sb, int64_t d_sn, int64_t d_sc,
                   int N, int M, int C, int k, int normalize,
                   int64_t* __restrict__ idx_out, float* __restrict__ dist_out, float* __restrict__ edge_out) {
    extern __shared__ __align__(16) float g_sm[];
    float (*s_a)[kGQ + 4] = reinterpret_cast<float (*)[kGQ + 4]>(g_sm);                       // query chunk, transposed
    float (*s_b)[kGC + 4] = reinterpret_cast<float (*)[kGC + 4]>(g_sm + kGK * (kGQ + 4));     // candidate chunk, transposed
    float (*s_dist)[kGC + 1] = reinterpret_cast<float (*)[kGC + 1]>(g_sm + kGK * (kGQ + 4) + kGK * (kGC + 4));
    float* s_qn = g_sm + kGK * (kGQ + 4) + kGK * (kGC + 4) + kGQ * (kGC + 1);
    float* s_cn = s_qn + kGQ;
    float* s_stage_d = s_cn + kGC;
    int* s_stage_i = reinterpret_cast<int*>(s_stage_d + kStage * kGQ);

    const int b = blockIdx.y, tid = threadIdx.x;
    const int q0 = blockIdx.x * kGQ;
    const float* sb = src + (int64_t)b * s_sb;
    const float* db = dst + (int64_t)b * d_sb;
    const int tq = (tid >> 4) * 8;      // 16 thread rows x 8 queries
    const int tc = (tid & 15) * 4;      // 16 thread cols x 4 candidates

    // |q|^2 per query (sequential over c, products rounded separately like sum(src ** 2, -1))
    if (tid < kGQ) {
        float acc = 0.f;
        if (q0 + tid < N && !normalize)
            for (int c = 0; c < C; ++c) { const float v = sb[(int64_t)(q0 + tid) * s_sn + (int64_t)c * s_sc]; acc = __fadd_rn(acc, __fmul_rn(v, v)); }
        s_qn[tid] = normalize ? 2.0f : acc;
    }

    TopK<K> top;
    const bool selector = tid < kGQ;
    const bool valid = selector && (q0 + tid < N);
    top.init(s_stage_d, s_stage_i, kGQ, tid & (kGQ - 1), valid);

    for (int m0 = 0; m0 < M; m0 += kGC) {
        float acc[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        __syncthreads();
        if (tid < kGC) {
            float a = 0.f;
            if (m0 + tid < M && !normalize)
                for (int c = 0; c < C; ++c) { const float v = db[(int64_t)(m0 + tid) * d_sn + (int64_t)c * d_sc]; a = __fadd_rn(a, __fmul_rn(v, v)); }
            s_cn[tid] = (m0 + tid < M) ? (normalize ? 0.f : a) : INFINITY;
        }
        for (int c0 = 0; c0 < C; c0 += kGK) {
            __syncthreads();
            for (int e = tid; e < kGQ * kGK; e += kKnnThreads) {
                int r, c;
                if (s_sc == 1) { r = e / kGK; c = e - r * kGK; } else { c = e / kGQ; r = e - c * kGQ; }
                float v = 0.f;
                if (q0 + r < N && c0 + c < C) v = sb[(int64_t)(q0 + r) * s_sn + (int64_t)(c0 + c) * s_sc];
                s_a[c][r] = v;
            }
            for (int e = tid; e < kGC * kGK; e += kKnnThreads) {
                int r, c;
                if (d_sc == 1) { r = e / kGK; c = e - r * kGK; } else { c = e / kGC; r = e - c * kGC; }
                float v = 0.f;
                if (m0 + r < M && c0 + c < C) v = db[(int64_t)(m0 + r) * d_sn + (int64_t)(c0 + c) * d_sc];
                s_b[c][r] = -2.f * v;
            }
            __syncthreads();
#pragma unroll
            for (int c = 0; c < kGK; ++c) {
                float av[8], bv[4];
#pragma unroll
                for (int i = 0; i < 8; i += 4) *reinterpret_cast<float4*>(av + i) = *reinterpret_cast<const float4*>(&s_a[c][tq + i]);
                *reinterpret_cast<float4*>(bv) = *reinterpret_cast<const float4*>(&s_b[c][tc]);
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float v = __fadd_rn(__fadd_rn(acc[i][j], s_qn[tq + i]), s_cn[tc + j]);
                if (!normalize) v = fmaxf(v, 1e-12f);
                s_dist[tq + i][tc + j] = v;
            }
        __syncthreads();
        if (tid < kGQ) {                      // warps 0..3 are the selectors (warp-uniform branch)
            for (int c = 0; c < kGC; c += 4) {
#pragma unroll
                for (int u = 0; u < 4; ++u) top.offer(s_dist[tid][c + u], m0 + c + u);
                top.maybe_merge();
            }
        }
    }
    if (tid < kGQ) top.merge();
    if (!valid) return;
    const int q = q0 + tid;
    int64_t* io = idx_out + ((int64_t)b * N + q) * k;
    float* dout = dist_out ? dist_out + ((int64_t)b * N + q) * k : nullptr;
#pragma unroll
    for (int j = 0; j < K; ++j)
        if (j < k) { io[j] = top.i[j]; if (dout) dout[j] = top.d[j]; }
    if (edge_out) {
        float* eo = edge_out + ((int64_t)b * N + q) * (int64_t)k * 2 * C;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            if (j < k) {
                for (int c = 0; c < C; ++c) {
                    const float ctr = sb[(int64_t)q * s_sn + (int64_t)c * s_sc];
                    eo[(int64_t)j * 2 * C + c] = sb[(int64_t)top.i[j] * s_sn + (int64_t)c * s_sc] - ctr;
                    eo[(int64_t)j * 2 * C + C + c] = ctr;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// edge gather: out[b][n][kk][0:C] = x[:, idx] - x[:, n]; out[..][C:2C] = x[:, n].
// Two kernels, both one cloud per blockIdx.y with 32-bit index arithmetic inside the cloud:
//   C <= 4 (xyz, the DGCNN first layer): one thread per (n, kk) row -- idx and the centre are read once per row,
//     the 2C floats are written as 8-byte pairs; a warp covers 32 consecutive rows = one contiguous span.
//   any C: one thread per output element, consecutive threads write consecutive floats.
template <int C>
__global__ void __launch_bounds__(256)
edge_gather_rows_kernel(const float* __restrict__ x, int64_t sb, int64_t sc, int64_t sn, const int64_t* __restrict__ idx,
                        int N, int k, float* __restrict__ out) {
    // the CTA's 256 rows are one contiguous span of 256 * 2C floats: staged in shared memory (row pitch 2C + 1 keeps
    // the per-thread writes conflict-free) and written back with consecutive threads on consecutive floats
    __shared__ float s_out[256 * (2 * C + 1)];
    const int rows = N * k;
    const int r0 = blockIdx.x * blockDim.x, r = r0 + threadIdx.x;
    const int b = blockIdx.y;
    if (r < rows) {
        const int n = r / k;
        const float* xb = x + (int64_t)b * sb;
        int64_t j = idx[(int64_t)b * rows + r];
        j = j < 0 ? 0 : (j >= N ? N - 1 : j);
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const float ctr = __ldg(xb + (int64_t)c * sc + (int64_t)n * sn);
            s_out[threadIdx.x * (2 * C + 1) + c] = __ldg(xb + (int64_t)c * sc + j * sn) - ctr;
            s_out[threadIdx.x * (2 * C + 1) + C + c] = ctr;
        }
    }
    __syncthreads();
    const int n_rows = min(256, rows - r0);
    float* o = out + ((int64_t)b * rows + r0) * (2 * C);
    for (int e = threadIdx.x; e < n_rows * 2 * C; e += 256) {
        const int rr = e / (2 * C), cc = e - rr * (2 * C);
        o[e] = s_out[rr * (2 * C + 1) + cc];
    }
}

__global__ void __launch_bounds__(256)
edge_gather_kernel(const float* __restrict__ x, int64_t sb, int64_t sc, int64_t sn, const int64_t* __restrict__ idx,
                   int C, int N, int k, float* __restrict__ out) {
    const unsigned per_cloud = (unsigned)N * (unsigned)k * 2u * (unsigned)C;
    const unsigned e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= per_cloud) return;
    const int b = blockIdx.y;
    const unsigned c2 = e % (2u * C);
    const unsigned row = e / (2u * C);        // (n, kk)
    const unsigned n = row / (unsigned)k;
    const float* xb = x + (int64_t)b * sb;
    const unsigned c = c2 < (unsigned)C ? c2 : c2 - C;
    const float ctr = __ldg(xb + (int64_t)c * sc + (int64_t)n * sn);
    float v = ctr;
    if (c2 < (unsigned)C) {
        int64_t j = idx[(int64_t)b * N * k + row];
        j = j < 0 ? 0 : (j >= N ? N - 1 : j);
        v = __ldg(xb + (int64_t)c * sc + j * sn) - ctr;
    }
    out[(int64_t)b * per_cloud + e] = v;
}

// ---------------------------------------------------------------------------------------------------
// square_distance alone (lib/utils.py:12-34): the dense (B,N,M) matrix for callers that want it (losses, metrics).
// Same arithmetic as the selection kernels -- ((-2 s.d) + |s|^2) + |d|^2, FMA chain over c, clamp 1e-12; or
// 2 + (-2 s.d) with normalize -- so dist_out of ogmm_knn_graph is a gather of this matrix, bit for bit.
// CTA = 32 queries x 128 candidates; a thread owns 4 consecutive candidates of 4 queries; candidate coordinates are
// staged (already scaled by -2) in shared memory, stores are 16-byte and coalesced along M.  Write-bound.
constexpr int kSqQ = 32, kSqC = 128, kSqChunk = 32;

__global__ void __launch_bounds__(256)
square_distance_kernel(const float* __restrict__ src, int64_t s_sb, int64_t s_sn, int64_t s_sc,
                       const float* __restrict__ dst, int64_t d_sb, int64_t d_sn, int64_t d_sc,
                       int N, int M, int C, int normalize, float* __restrict__ out) {
    __shared__ float s_q[kSqChunk][kSqQ + 1];
    __shared__ float s_c[kSqChunk][kSqC + 4];
    __shared__ float s_qn[kSqQ], s_cn[kSqC];
    const int b = blockIdx.z, q0 = blockIdx.y * kSqQ, m0 = blockIdx.x * kSqC, tid = threadIdx.x;
    const float* sb = src + (int64_t)b * s_sb;
    const float* db = dst + (int64_t)b * d_sb;
    const int tq = (tid >> 5) * 4, tc = (tid & 31) * 4;
    if (tid < kSqQ) {
        float acc = 0.f;
        if (q0 + tid < N)
            for (int c = 0; c < C; ++c) { const float v = sb[(int64_t)(q0 + tid) * s_sn + (int64_t)c * s_sc]; acc = __fadd_rn(acc, __fmul_rn(v, v)); }
        s_qn[tid] = acc;
    } else if (tid < kSqQ + kSqC) {
        const int m = tid - kSqQ;
        float acc = 0.f;
        if (m0 + m < M)
            for (int c = 0; c < C; ++c) { const float v = db[(int64_t)(m0 + m) * d_sn + (int64_t)c * d_sc]; acc = __fadd_rn(acc, __fmul_rn(v, v)); }
        s_cn[m] = acc;
    }
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int c0 = 0; c0 < C; c0 += kSqChunk) {
        __syncthreads();
        for (int e = tid; e < kSqQ * kSqChunk; e += 256) {
            const int c = e / kSqQ, r = e - c * kSqQ;
            s_q[c][r] = (q0 + r < N && c0 + c < C) ? sb[(int64_t)(q0 + r) * s_sn + (int64_t)(c0 + c) * s_sc] : 0.f;
        }
        for (int e = tid; e < kSqC * kSqChunk; e += 256) {
            const int c = e / kSqC, r = e - c * kSqC;
            s_c[c][r] = (m0 + r < M && c0 + c < C) ? -2.f * db[(int64_t)(m0 + r) * d_sn + (int64_t)(c0 + c) * d_sc] : 0.f;
        }
        __syncthreads();
        const int cw = min(kSqChunk, C - c0);
        for (int c = 0; c < cw; ++c) {
            const float4 cv = *reinterpret_cast<const float4*>(&s_c[c][tc]);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float qv = s_q[c][tq + i];
                if (c0 + c == 0) {          // the first product is rounded on its own (a GEMM's accumulator starts at 0)
                    acc[i][0] = __fmul_rn(qv, cv.x); acc[i][1] = __fmul_rn(qv, cv.y);
                    acc[i][2] = __fmul_rn(qv, cv.z); acc[i][3] = __fmul_rn(qv, cv.w);
                } else {
                    acc[i][0] = fmaf(qv, cv.x, acc[i][0]); acc[i][1] = fmaf(qv, cv.y, acc[i][1]);
                    acc[i][2] = fmaf(qv, cv.z, acc[i][2]); acc[i][3] = fmaf(qv, cv.w, acc[i][3]);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int q = q0 + tq + i;
        if (q >= N) continue;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (normalize) v[j] = __fadd_rn(acc[i][j], 2.0f);
            else v[j] = fmaxf(__fadd_rn(__fadd_rn(acc[i][j], s_qn[tq + i]), s_cn[tc + j]), 1e-12f);
        }
        float* o = out + ((int64_t)b * N + q) * M + m0 + tc;
        if ((M & 3) == 0 && m0 + tc + 3 < M) *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
        else
#pragma unroll
            for (int j = 0; j < 4; ++j) if (m0 + tc + j < M) o[j] = v[j];
    }
}

}  // namespace ogmm

using namespace ogmm;

int ogmm_launch_knn3_sweep(const float* src, int64_t s_sb, int64_t s_sn, int64_t s_sc,
                           const float* dst, int64_t d_sb, int64_t d_sn, int64_t d_sc,
                           int64_t B, int64_t N, int64_t M, int64_t k,
                           int64_t* idx_out, float* dist_out, float* edge_out, cudaStream_t s);

int ogmm_launch_knn3_select(const float* src, int64_t s_sb, int64_t s_sn, int64_t s_sc,
                            const float* dst, int64_t d_sb, int64_t d_sn, int64_t d_sc,
                            int64_t B, int64_t N, int64_t M, int64_t k,
                            int64_t* idx_out, float* dist_out, float* edge_out, int32_t* stats, cudaStream_t s);

int ogmm_launch_knn_wide(const float* src, int64_t s_sb, int64_t s_sn, int64_t s_sc,
                         const float* dst, int64_t d_sb, int64_t d_sn, int64_t d_sc,
                         int64_t B, int64_t N, int64_t M, int64_t C, int64_t k, int normalize,
                         int64_t* idx_out, float* dist_out, int32_t* stats, cudaStream_t s);

int ogmm_launch_knn3_tiles(const float* src, int64_t s_sb, int64_t s_sn, int64_t s_sc, int64_t B, int64_t M, int64_t k,
                           int64_t* idx_out, float* dist_out, float* edge_out, cudaStream_t s);

extern "C" __attribute__((visibility("default"))) int ogmm_knn_graph(const float* src, int64_t s_sb, int64_t s_sn, int64_t s_sc,
                              const float* dst, int64_t d_sb, int64_t d_sn, int64_t d_sc,
                              int64_t B, int64_t N, int64_t M, int64_t C, int64_t k, int normalize,
                              int64_t* idx_out, float* dist_out, float* edge_out, ogmm_stream_t stream) {
    OGMM_REQUIRE(B >= 0 && N >= 1 && M >= 1 && C >= 1 && k >= 1 && N < (1ll << 31) && M < (1ll << 31) && B < 65536,
                 OGMM_EINVAL, "ogmm_knn_graph: bad sizes B=%lld N=%lld M=%lld C=%lld k=%lld", (long long)B,
                 (long long)N, (long long)M, (long long)C, (long long)k);
    OGMM_REQUIRE(k <= M, OGMM_EINVAL, "ogmm_knn_graph: k=%lld exceeds the number of candidates M=%lld", (long long)k, (long long)M);
    OGMM_REQUIRE(k <= 64, OGMM_EUNSUPPORTED, "ogmm_knn_graph: k=%lld > 64", (long long)k);
    if (B == 0) return OGMM_OK;
    OGMM_REQUIRE(src && dst && idx_out, OGMM_EINVAL, "ogmm_knn_graph: null pointer");
    OGMM_REQUIRE(edge_out == nullptr || (N == M && src == dst && s_sb == d_sb && s_sn == d_sn && s_sc == d_sc), OGMM_EINVAL,
                 "ogmm_knn_graph: edge_out needs a self graph (dst must be the same view as src)");
    cudaStream_t s = as_stream(stream);
    // 3-D clouds of up to 4096 points: sorted sweep with slab pruning (knn_sweep.cu); OGMM_KNN_EXHAUSTIVE=1
    // forces the exhaustive kernel (same results; kept for larger clouds and for A/B timing)
    if (C == 3 && !normalize && N <= 4096 && M <= 4096) {
        const char* force = getenv("OGMM_KNN_EXHAUSTIVE");
        const char* old_sweep = getenv("OGMM_KNN_SWEEP_INSERT");
        // threshold-then-collect selection (knn_select.cu) where its group bound is defined: at least k groups of four
        // candidates with room to spare, k within the revisit budget; OGMM_KNN_SWEEP_INSERT=1 keeps the insert-while-
        // sweeping kernel for A/B timing (bit-identical results)
        if (!(force && force[0] == '1') && !(old_sweep && old_sweep[0] == '1') && M >= 256 && k <= 24)
            return ogmm_launch_knn3_select(src, s_sb, s_sn, s_sc, dst, d_sb, d_sn, d_sc, B, N, M, k, idx_out, dist_out,
                                           edge_out, nullptr, s);
        if (!(force && force[0] == '1'))
            return ogmm_launch_knn3_sweep(src, s_sb, s_sn, s_sc, dst, d_sb, d_sn, d_sc, B, N, M, k, idx_out, dist_out,
                                          edge_out, s);
    }
    // large 3-D self graphs (4096 < N <= 16384): pre-sort + tiled sorted sweep (knn_tiles.cu); the exhaustive kernel below
    // for two-cloud calls, the cosine form and larger clouds (and with OGMM_KNN_EXHAUSTIVE=1, same results)
    if (C == 3 && !normalize && N == M && src == dst && s_sb == d_sb && s_sn == d_sn && s_sc == d_sc && M > 4096) {
        const char* force = getenv("OGMM_KNN_EXHAUSTIVE");
        if (!(force && force[0] == '1')) {
            const int st = ogmm_launch_knn3_tiles(src, s_sb, s_sn, s_sc, B, M, k, idx_out, dist_out, edge_out, s);
            if (st != OGMM_EUNSUPPORTED) return st;
        }
    }
    // feature-space graphs (32 <= C <= 256, k <= 32): Gram tiles on the tensor cores (knn_wide.cu), exact FP32 re-rank;
    // OGMM_KNN_NO_TENSOR=1 forces the FP32 FMA kernel (bit-identical results; for A/B checks)
    if (C >= 32 && C <= 256 && k <= 32) {
        const char* force = getenv("OGMM_KNN_NO_TENSOR");
        if (!(force && force[0] == '1')) {
            int st = ogmm_launch_knn_wide(src, s_sb, s_sn, s_sc, dst, d_sb, d_sn, d_sc, B, N, M, C, k, normalize, idx_out,
                                          dist_out, nullptr, s);
            if (st != OGMM_OK || edge_out == nullptr) return st;
            return ogmm_edge_gather(src, s_sb, s_sc, s_sn, idx_out, B, C, N, k, edge_out, stream);
        }
    }
    const bool three = (C == 3);
    dim3 grid((unsigned)((N + (three ? kKnnThreads : kGQ) - 1) / (three ? kKnnThreads : kGQ)), (unsigned)B);
#define LAUNCH(KK)                                                                                                   \
    do {                                                                                                             \
        if (three)                                                                                                   \
            knn3_kernel<KK><<<grid, kKnnThreads, 0, s>>>(src, s_sb, s_sn, s_sc, dst, d_sb, d_sn, d_sc, (int)N,      \
                                                         (int)M, (int)k, normalize, idx_out, dist_out, edge_out);   \
        else {                                                                                                       \
            int st = cuda_status(cudaFuncSetAttribute(knn_generic_kernel<KK>,                                        \
                                                      cudaFuncAttributeMaxDynamicSharedMemorySize,                   \
                                                      (int)kGenericSmem), "cudaFuncSetAttribute(knn_generic)");     \
            if (st != OGMM_OK) return st;                                                                            \
            knn_generic_kernel<KK><<<grid, kKnnThreads, kGenericSmem, s>>>(src, s_sb, s_sn, s_sc, dst, d_sb, d_sn,  \
                                                                           d_sc, (int)N, (int)M, (int)C, (int)k,    \
                                                                           normalize, idx_out, dist_out, edge_out); \
        }                                                                                                            \
    } while (0)
    if (k <= 4) LAUNCH(4);
    else if (k <= 8) LAUNCH(8);
    else if (k <= 16) LAUNCH(16);
    else if (k <= 20) LAUNCH(20);
    else if (k <= 32) LAUNCH(32);
    else LAUNCH(64);
#undef LAUNCH
    OGMM_LAUNCH_CHECK("knn kernel");
    return OGMM_OK;
}

// Diagnostic twin of the C == 3 route of ogmm_knn_graph: the selection kernel with its counters (see knn_select.cu).
extern "C" __attribute__((visibility("default"))) int ogmm_knn3_select_stats(const float* src, int64_t s_sb, int64_t s_sn, int64_t s_sc,
                                                                           const float* dst, int64_t d_sb, int64_t d_sn, int64_t d_sc,
                                                                           int64_t B, int64_t N, int64_t M, int64_t k,
                                                                           int64_t* idx_out, int32_t* stats, ogmm_stream_t stream) {
    OGMM_REQUIRE(B >= 1 && N >= 1 && M >= 256 && M <= 4096 && N <= 4096 && k >= 1 && k <= 24 && k <= M && B < 65536, OGMM_EUNSUPPORTED,
                 "ogmm_knn3_select_stats: outside the selection kernel's range");
    OGMM_REQUIRE(src && dst && idx_out && stats, OGMM_EINVAL, "ogmm_knn3_select_stats: null pointer");
    return ogmm_launch_knn3_select(src, s_sb, s_sn, s_sc, dst, d_sb, d_sn, d_sc, B, N, M, k, idx_out, nullptr, nullptr, stats,
                                   as_stream(stream));
}

extern "C" __attribute__((visibility("default"))) int ogmm_edge_gather(const float* x, int64_t sb, int64_t sc, int64_t sn, const int64_t* idx,
                                int64_t B, int64_t C, int64_t N, int64_t k, float* edge_out, ogmm_stream_t stream) {
    OGMM_REQUIRE(B >= 0 && C >= 1 && N >= 1 && k >= 1 && N < (1ll << 31) && C < (1ll << 20) && k < (1ll << 20),
                 OGMM_EINVAL, "ogmm_edge_gather: bad sizes");
    if (B == 0) return OGMM_OK;
    OGMM_REQUIRE(x && idx && edge_out, OGMM_EINVAL, "ogmm_edge_gather: null pointer");
    OGMM_REQUIRE(B < 65536 && N * k * 2 * C < (1ll << 31), OGMM_EUNSUPPORTED, "ogmm_edge_gather: a cloud's edge tensor exceeds 2^31 elements");
    cudaStream_t s = as_stream(stream);
    if (C <= 4) {
        dim3 grid((unsigned)((N * k + 255) / 256), (unsigned)B);
        switch (C) {
            case 1: edge_gather_rows_kernel<1><<<grid, 256, 0, s>>>(x, sb, sc, sn, idx, (int)N, (int)k, edge_out); break;
            case 2: edge_gather_rows_kernel<2><<<grid, 256, 0, s>>>(x, sb, sc, sn, idx, (int)N, (int)k, edge_out); break;
            case 3: edge_gather_rows_kernel<3><<<grid, 256, 0, s>>>(x, sb, sc, sn, idx, (int)N, (int)k, edge_out); break;
            default: edge_gather_rows_kernel<4><<<grid, 256, 0, s>>>(x, sb, sc, sn, idx, (int)N, (int)k, edge_out); break;
        }
    } else {
        dim3 grid((unsigned)((N * k * 2 * C + 255) / 256), (unsigned)B);
        edge_gather_kernel<<<grid, 256, 0, s>>>(x, sb, sc, sn, idx, (int)C, (int)N, (int)k, edge_out);
    }
    OGMM_LAUNCH_CHECK("edge_gather_kernel");
    return OGMM_OK;
}

extern "C" __attribute__((visibility("default"))) int ogmm_knn_wide(const float* src, int64_t s_sb, int64_t s_sn, int64_t s_sc,
                                                                  const float* dst, int64_t d_sb, int64_t d_sn, int64_t d_sc,
                                                                  int64_t B, int64_t N, int64_t M, int64_t C, int64_t k, int normalize,
                                                                  int64_t* idx_out, float* dist_out, int32_t* fallback_count,
                                                                  ogmm_stream_t stream) {
    OGMM_REQUIRE(B >= 0 && N >= 1 && M >= 1 && k >= 1 && N < (1ll << 31) && M < (1ll << 31) && B < 65536, OGMM_EINVAL,
                 "ogmm_knn_wide: bad sizes");
    OGMM_REQUIRE(C >= 32 && C <= 256, OGMM_EUNSUPPORTED, "ogmm_knn_wide: C=%lld outside [32, 256]", (long long)C);
    OGMM_REQUIRE(k <= M, OGMM_EINVAL, "ogmm_knn_wide: k=%lld exceeds M=%lld", (long long)k, (long long)M);
    if (B == 0) return OGMM_OK;
    OGMM_REQUIRE(src && dst && idx_out, OGMM_EINVAL, "ogmm_knn_wide: null pointer");
    return ogmm_launch_knn_wide(src, s_sb, s_sn, s_sc, dst, d_sb, d_sn, d_sc, B, N, M, C, k, normalize, idx_out, dist_out,
                                fallback_count, as_stream(stream));
}

extern "C" __attribute__((visibility("default"))) int ogmm_square_distance(const float* src, int64_t s_sb, int64_t s_sn, int64_t s_sc,
                                                                         const float* dst, int64_t d_sb, int64_t d_sn, int64_t d_sc,
                                                                         int64_t B, int64_t N, int64_t M, int64_t C, int normalize,
                                                                         float* dist_out, ogmm_stream_t stream) {
    OGMM_REQUIRE(B >= 0 && N >= 1 && M >= 1 && C >= 1 && N < (1ll << 31) && M < (1ll << 31) && B < 65536, OGMM_EINVAL,
                 "ogmm_square_distance: bad sizes B=%lld N=%lld M=%lld C=%lld", (long long)B, (long long)N, (long long)M, (long long)C);
    if (B == 0) return OGMM_OK;
    OGMM_REQUIRE(src && dst && dist_out, OGMM_EINVAL, "ogmm_square_distance: null pointer");
    OGMM_REQUIRE((N + kSqQ - 1) / kSqQ < 65536, OGMM_EUNSUPPORTED, "ogmm_square_distance: N=%lld too large", (long long)N);
    dim3 grid((unsigned)((M + kSqC - 1) / kSqC), (unsigned)((N + kSqQ - 1) / kSqQ), (unsigned)B);
    square_distance_kernel<<<grid, 256, 0, as_stream(stream)>>>(src, s_sb, s_sn, s_sc, dst, d_sb, d_sn, d_sc, (int)N, (int)M,
                                                                 (int)C, normalize, dist_out);
    OGMM_LAUNCH_CHECK("square_distance_kernel");
    return OGMM_OK;
}
