// K1 (3-D, 256 <= M <= 4096 candidates, k <= 24): exact kNN by threshold-then-collect selection (sm_100a).
//
// Same contract and the same distance arithmetic as the other 3-D kernels -- the reference's expanded form
// ((-2 s.d) + |s|^2) + |d|^2 with clamp 1e-12 (lib/utils.py:28-33), k smallest per row, ascending, ties to the lowest
// index (lib/utils.py:43) -- and the same sorted sweep with slab pruning as knn_sweep.cu, but the per-query list is
// no longer maintained by sorted inserts while sweeping (39 % of that kernel's instructions, ~190 warp-wide insert
// rounds per 32 queries for ~81 useful inserts per query).  Instead, per query thread:
//
//   pass 1  sweeps outwards in GROUPS of four consecutive sorted candidates and keeps only one number per group: the
//           smallest of its four distances, with the group id packed into the low mantissa bits.  Sixteen such group
//           keys per step are sorted by a register network (FMNMX pairs, no indices, no branches) and merged into the
//           32 smallest keys seen so far.  Its k-th entry is an upper bound of the true k-th distance (k different
//           groups each hold a candidate at or below it) and prunes the sweep.
//   pass 2  revisits ONLY the groups named by the first k (+ ties) keys -- about 80 candidates per query instead of the
//           ~450 of the slab, each lane walking its own groups -- eight groups at a time: the candidates at or below
//           the bound (at most 32 per chunk, so the shared-memory column cannot overflow) are appended to the column
//           and then inserted into the sorted 64-bit (distance, index) list; every lane has work in every round, so
//           ~30 rounds replace ~190.  Once the list is full its exact k-th distance tightens the bound for the
//           remaining chunks (the keys are in ascending order of group minimum, so the early chunks hold the winners).
//
// Exactness.  Let T be the key prefix (all bits above the group id) of the k-th smallest group key.  Any true
// neighbour c lies in a group whose minimum is <= d(c) <= (k-th distance); if that group's prefix exceeded T, the k
// group minima of the first k keys would all be strictly smaller than d(c), contradicting c being among the k smallest.
// So every true neighbour sits in a group whose key prefix is <= T: the first k keys plus any later keys with prefix
// == T (the list keeps 32, so up to 12 such ties are visible).  The bound used for pruning and collecting is the k-th
// key with its id bits set to one, i.e. rounded UP.  Warps that meet more prefix ties than fit (clouds of coincident
// points) or fewer than k groups redo their 32 queries exhaustively (exact, slow, rare).
#include "knn_common.cuh"

namespace ogmm {

constexpr int kSelCap = 32;                 // collected candidates per query and chunk (8 groups x 4) held in shared memory
constexpr int kSelKeep = 32;                // group keys kept per query (registers)
constexpr int kSelVisit = 24;               // groups revisited at most: k + prefix ties
constexpr float kSelBig = 1.0e38f;          // finite sentinel: stays finite with id bits OR-ed in

// Group key: min of the four distances (clamped like the reference), id bits replaced by the group id.
__device__ __forceinline__ float group_key(unsigned rec_base, int g, const QueryPack& q, unsigned gmask) {
    float v[4];
    group_distances(rec_base + 64u * (unsigned)g, q, v);
    const float m = fmaxf(fminf(fminf(v[0], v[1]), fminf(v[2], v[3])), 1e-12f);
    return __uint_as_float((__float_as_uint(m) & ~gmask) | (unsigned)g);
}

// ---- register sorting networks on float keys (finite, positive: float order == bit order) -----------------------
template <int N>
__device__ __forceinline__ void bitonic_sort_regs(float (&a)[N]) {
#pragma unroll
    for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const int l = i ^ j;
                if (l > i) {
                    const float lo = fminf(a[i], a[l]), hi = fmaxf(a[i], a[l]);
                    if ((i & k) == 0) { a[i] = lo; a[l] = hi; } else { a[i] = hi; a[l] = lo; }
                }
            }
        }
    }
}
// keep (sorted ascending, kSelKeep) <- the kSelKeep smallest of keep U fresh (fresh sorted ascending, 16)
__device__ __forceinline__ void merge_keys(float (&keep)[kSelKeep], const float (&fresh)[16]) {
#pragma unroll
    for (int i = 0; i < 16; ++i) keep[16 + i] = fminf(keep[16 + i], fresh[15 - i]);     // bitonic: asc 16, then min(asc, desc)
#pragma unroll
    for (int j = kSelKeep >> 1; j > 0; j >>= 1) {
#pragma unroll
        for (int i = 0; i < kSelKeep; ++i) {
            const int l = i ^ j;
            if (l > i) { const float lo = fminf(keep[i], keep[l]), hi = fmaxf(keep[i], keep[l]); keep[i] = lo; keep[l] = hi; }
        }
    }
}

__host__ __device__ inline size_t select_smem_bytes(int N, int M, bool self) {
    const int NG = (M + 3) / 4, Mp = pow2_ge(M), Np = self ? 0 : pow2_ge(N);
    return (size_t)64 * NG + (size_t)8 * Mp + (size_t)8 * Np + (size_t)6 * kSelCap * kSwThreads + 256;
}

template <int K>
__global__ void __launch_bounds__(kSwThreads, 2)
knn3_select_kernel(const float* __restrict__ src, int64_t s_sb, int64_t s_sn, int64_t s_sc,
                   const float* __restrict__ dst, int64_t d_sb, int64_t d_sn, int64_t d_sc,
                   int B, int N, int M, int k, int self, int items_per_cta,
                   int64_t* __restrict__ idx_out, float* __restrict__ dist_out, float* __restrict__ edge_out,
                   int32_t* __restrict__ stats) {
    extern __shared__ __align__(16) unsigned char sel_raw[];
    const int NG = (M + 3) / 4, Mp = pow2_ge(M), Np = self ? 0 : pow2_ge(N);
    float4* s_rec = reinterpret_cast<float4*>(sel_raw);                         // [NG][4] packed group records
    float* s_ckey = reinterpret_cast<float*>(s_rec + 4 * NG);                   // [Mp] sorted axis keys
    int* s_cord = reinterpret_cast<int*>(s_ckey + Mp);                          // [Mp] sorted position -> original index
    float* s_qkey = reinterpret_cast<float*>(s_cord + Mp);
    int* s_qord = reinterpret_cast<int*>(s_qkey + Np);
    float* s_buf_d = reinterpret_cast<float*>(s_qord + Np);                     // [kSelCap][256] collected distances
    unsigned short* s_buf_p = reinterpret_cast<unsigned short*>(s_buf_d + kSelCap * kSwThreads);    // sorted positions
    float* s_red = reinterpret_cast<float*>(s_buf_p + kSelCap * kSwThreads);    // [64]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned gmask = NG <= 256 ? 0xffu : 0x3ffu;                          // M <= 4096 -> NG <= 1024
    const int blocks_per_cloud = (N + kSwQueriesPerCta - 1) / kSwQueriesPerCta;
    const int64_t items = (int64_t)B * blocks_per_cloud;
    // contiguous item range per CTA: consecutive items share a cloud, so the cloud is staged and sorted once for them.
    // items_per_cta > 0: fixed-size ranges that start on cloud boundaries (or divide a cloud evenly), chosen by the host
    // when that costs no extra round; otherwise the items are split proportionally over the grid.
    const int64_t it_begin = items_per_cta > 0 ? (int64_t)blockIdx.x * items_per_cta : items * blockIdx.x / gridDim.x;
    const int64_t it_end = items_per_cta > 0 ? min(items, it_begin + items_per_cta) : items * (blockIdx.x + 1) / gridDim.x;
    int staged = -1;
    float cn_max = 0.f;
    int axis = 0;

    for (int64_t item = it_begin; item < it_end; ++item) {
        const int b = (int)(item / blocks_per_cloud), qb = (int)(item - (int64_t)b * blocks_per_cloud);
        const float* sb = src + (int64_t)b * s_sb;
        const float* db = dst + (int64_t)b * d_sb;
        if (b != staged) {
            __syncthreads();                               // every warp is done with the previous cloud
            // ---- axis of largest extent, largest |c|^2 (for the error margin) -------------------------------
            float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY}, nmax = 0.f;
            for (int m = tid; m < M; m += kSwThreads) {
                const float* p = db + (int64_t)m * d_sn;
                const float x = p[0], y = p[d_sc], z = p[2 * d_sc];
                lo[0] = fminf(lo[0], x); hi[0] = fmaxf(hi[0], x);
                lo[1] = fminf(lo[1], y); hi[1] = fmaxf(hi[1], y);
                lo[2] = fminf(lo[2], z); hi[2] = fmaxf(hi[2], z);
                nmax = fmaxf(nmax, sqn3(x, y, z));
            }
#pragma unroll
            for (int a = 0; a < 3; ++a) { lo[a] = -warp_max(-lo[a]); hi[a] = warp_max(hi[a]); }
            nmax = warp_max(nmax);
            if (lane == 0) {
#pragma unroll
                for (int a = 0; a < 3; ++a) { s_red[warp * 8 + a] = lo[a]; s_red[warp * 8 + 3 + a] = hi[a]; }
                s_red[warp * 8 + 6] = nmax;
            }
            __syncthreads();
            float ext[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                float l = INFINITY, h = -INFINITY;
                for (int w = 0; w < kSwThreads / 32; ++w) { l = fminf(l, s_red[w * 8 + a]); h = fmaxf(h, s_red[w * 8 + 3 + a]); }
                ext[a] = h - l;
            }
            cn_max = 0.f;
            for (int w = 0; w < kSwThreads / 32; ++w) cn_max = fmaxf(cn_max, s_red[w * 8 + 6]);
            axis = (ext[0] >= ext[1] && ext[0] >= ext[2]) ? 0 : (ext[1] >= ext[2] ? 1 : 2);
            const int64_t d_ax = (int64_t)axis * d_sc, s_ax = (int64_t)axis * s_sc;
            // ---- sort candidates (and, for a two-cloud call, queries) along the axis -----------------------------
            for (int m = tid; m < Mp; m += kSwThreads) {
                s_ckey[m] = m < M ? sort_key(db[(int64_t)m * d_sn + d_ax]) : INFINITY;
                s_cord[m] = m < M ? m : 0x7fffffff;
            }
            if (!self)
                for (int n = tid; n < Np; n += kSwThreads) {
                    s_qkey[n] = n < N ? sort_key(sb[(int64_t)n * s_sn + s_ax]) : INFINITY;
                    s_qord[n] = n < N ? n : 0x7fffffff;
                }
            __syncthreads();
            sort_pairs(s_ckey, s_cord, Mp);
            if (!self) sort_pairs(s_qkey, s_qord, Np);
            // ---- packed group records in sorted order; positions >= M are padding that never wins -------------
            float* recf = reinterpret_cast<float*>(s_rec);
            for (int p = tid; p < 4 * NG; p += kSwThreads) {
                float x = 0.f, y = 0.f, z = 0.f, w = kSelBig;
                if (p < M) {
                    const float* c = db + (int64_t)s_cord[p] * d_sn;
                    const float cx = c[0], cy = c[d_sc], cz = c[2 * d_sc];
                    x = -2.f * cx; y = -2.f * cy; z = -2.f * cz; w = sqn3(cx, cy, cz);
                }
                const int g = p >> 2, m = p & 3;
                float* r = recf + 16 * g + 8 * (m >> 1) + (m & 1);
                r[0] = x; r[2] = y; r[4] = z; r[6] = w;
            }
            __syncthreads();
            staged = b;
        }
        const float* qkeys = self ? s_ckey : s_qkey;
        const int* qords = self ? s_cord : s_qord;

        // ---- this item's 256 queries, one per thread, warps independent from here on ------------------------------
        const int r = qb * kSwQueriesPerCta + tid;
        const bool valid = r < N;
        if (!__any_sync(kFull, valid)) continue;
        int q = 0;
        float qx = 0.f, qy = 0.f, qz = 0.f, qk = 0.f;
        if (valid) {
            q = qords[r];
            const float* p = sb + (int64_t)q * s_sn;
            qx = p[0]; qy = p[s_sc]; qz = p[2 * s_sc];
            qk = qkeys[r];
        }
        const float qs = sqn3(qx, qy, qz);
        QueryPack Q;
        Q.x = pack2(qx, qx); Q.y = pack2(qy, qy); Q.z = pack2(qz, qz); Q.s = pack2(qs, qs);
        // fp32 error bound of the expanded form: the slab test must never prune a candidate that could still win
        const float margin = 1e-6f * (qs + cn_max) + 1e-30f;

        int pos;                                           // warp start: lower bound of the middle lane's key
        {
            const float mid = __shfl_sync(kFull, qk, 16);
            int lo_i = 0, hi_i = M;
            while (lo_i < hi_i) { const int md = (lo_i + hi_i) >> 1; if (s_ckey[md] < mid) lo_i = md + 1; else hi_i = md; }
            pos = lo_i;
        }

        // ================= pass 1: the 32 smallest group keys =================
        float keep[kSelKeep];
#pragma unroll
        for (int i = 0; i < kSelKeep; ++i) keep[i] = kSelBig;
        int gl = (pos >> 2) - 1, gr = pos >> 2;            // next group to visit on each side
        int steps = 0, merges = 0;                         // diagnostics (stats != nullptr only)
        const unsigned rec_base = (unsigned)__cvta_generic_to_shared(s_rec);
        while (true) {
            const float bound = __uint_as_float(__float_as_uint(keep[K - 1]) | gmask);      // rounded up
            const float tl = bound * (1.0f + 1e-6f) + margin;
            bool go_l = false, go_r = false;
            if (gl >= 0) { const float g = qk - s_ckey[min(4 * gl + 3, M - 1)]; go_l = valid && !(g > 0.f && g * g > tl); }
            if (gr < NG) { const float g = s_ckey[4 * gr] - qk; go_r = valid && !(g > 0.f && g * g > tl); }
            go_l = __any_sync(kFull, go_l);
            go_r = __any_sync(kFull, go_r);
            if (!go_l && !go_r) break;
            ++steps;
            float fresh[16];
            if (go_l) {
                if (gl >= 7) {                             // whole block in range (warp-uniform): no per-group checks
#pragma unroll
                    for (int u = 0; u < 8; ++u) fresh[u] = group_key(rec_base, gl - u, Q, gmask);
                } else {
#pragma unroll
                    for (int u = 0; u < 8; ++u) fresh[u] = (gl - u >= 0) ? group_key(rec_base, gl - u, Q, gmask) : kSelBig;
                }
                gl -= 8;
            } else {
#pragma unroll
                for (int u = 0; u < 8; ++u) fresh[u] = kSelBig;
            }
            if (go_r) {
                if (gr + 7 < NG) {
#pragma unroll
                    for (int u = 0; u < 8; ++u) fresh[8 + u] = group_key(rec_base, gr + u, Q, gmask);
                } else {
#pragma unroll
                    for (int u = 0; u < 8; ++u) fresh[8 + u] = (gr + u < NG) ? group_key(rec_base, gr + u, Q, gmask) : kSelBig;
                }
                gr += 8;
            } else {
#pragma unroll
                for (int u = 0; u < 8; ++u) fresh[8 + u] = kSelBig;
            }
            // Only the first kSelVisit + 1 kept keys have to be exact (the revisit list and the test for one tie too
            // many); a block in which no lane has a key below its (kSelVisit + 1)-th smallest changes none of them.
            float fmin16 = fresh[0];
#pragma unroll
            for (int u = 1; u < 16; ++u) fmin16 = fminf(fmin16, fresh[u]);
            if (!__any_sync(kFull, fmin16 < keep[kSelVisit])) continue;
            bitonic_sort_regs<16>(fresh);
            merge_keys(keep, fresh);
            ++merges;
        }

        // ================= pass 2: collect from the groups the first k (+ ties) keys name =================
        const unsigned kth = __float_as_uint(keep[K - 1]);
        const float tau = __uint_as_float(kth | gmask);
        int n_g = K;
#pragma unroll
        for (int j = K; j < kSelKeep; ++j) n_g += ((__float_as_uint(keep[j]) & ~gmask) == (kth & ~gmask)) ? 1 : 0;   // sorted: ties are contiguous
        bool redo = valid && (n_g > kSelVisit || !(keep[K - 1] < 0.5f * kSelBig));
        redo = __any_sync(kFull, redo);
        if (stats) {
            const int ng = __reduce_max_sync(kFull, valid ? n_g : 0);
            const unsigned ties = __ballot_sync(kFull, valid && n_g > kSelVisit);
            const unsigned few = __ballot_sync(kFull, valid && !(keep[K - 1] < 0.5f * kSelBig));
            if (lane == 0) {
                atomicAdd(stats + 1, 1); atomicAdd(stats + 2, steps); atomicAdd(stats + 3, merges);
                atomicAdd(stats + 5, ng); atomicAdd(stats + 7, ties != 0); atomicAdd(stats + 8, few != 0);
            }
        }

        TopK64<K> top;
        top.init(nullptr, nullptr, tid, valid);
        if (!redo) {
            float bound2 = tau;                            // collect bound; the exact k-th distance once the list is full
            int rounds = 0;
#pragma unroll
            for (int c0 = 0; c0 < kSelVisit; c0 += 8) {
                if (!__any_sync(kFull, valid && c0 < n_g)) break;
                int cnt = 0;
#pragma unroll
                for (int j = c0; j < c0 + 8; ++j) {
                    if (j < n_g) {
                        const int g = (int)(__float_as_uint(keep[j]) & gmask);
                        float v[4];
                        group_distances(rec_base + 64u * (unsigned)g, Q, v);
#pragma unroll
                        for (int m = 0; m < 4; ++m) {
                            if (v[m] <= bound2) {          // at most 32 per chunk: the column cannot overflow
                                s_buf_d[cnt * kSwThreads + tid] = fmaxf(v[m], 1e-12f);
                                s_buf_p[cnt * kSwThreads + tid] = (unsigned short)(4 * g + m);
                                ++cnt;
                            }
                        }
                    }
                }
                // exact (distance, index) order: every lane's s-th collected candidate goes in together
                const int most = __reduce_max_sync(kFull, valid ? cnt : 0);
                rounds += most;
                for (int s2 = 0; s2 < most; ++s2) {
                    u64 kv = kEmptyKey;
                    if (valid && s2 < cnt) {
                        const unsigned p = s_buf_p[s2 * kSwThreads + tid];
                        kv = ((u64)__float_as_uint(s_buf_d[s2 * kSwThreads + tid]) << 32) | ((unsigned)s_cord[p] << 16) | p;
                    }
                    if (__any_sync(kFull, kv < top.key[K - 1])) top.insert(kv);
                }
                bound2 = fminf(bound2, __uint_as_float((unsigned)(top.key[K - 1] >> 32)));      // +inf bits until the list is full
            }
            if (stats && lane == 0) atomicAdd(stats + 4, rounds);
        } else {
            // exhaustive redo of this warp's queries: every group, straight into the sorted list (exact by construction)
            if (lane == 0 && stats) atomicAdd(stats, 1);
            for (int g = 0; g < NG; ++g) {
                float v[4];
                group_distances(rec_base + 64u * (unsigned)g, Q, v);
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    const unsigned p = 4 * g + m;
                    u64 kv = kEmptyKey;
                    if (valid && p < (unsigned)M)
                        kv = ((u64)__float_as_uint(fmaxf(v[m], 1e-12f)) << 32) | ((unsigned)s_cord[p] << 16) | p;
                    if (__any_sync(kFull, kv < top.key[K - 1])) top.insert(kv);
                }
            }
        }

        // ---- results: idx (int64), optional distances, optional edge features [x_j - x_i ; x_i] -------------------
        if (valid) {
            int64_t* io = idx_out + ((int64_t)b * N + q) * k;
            float* dout = dist_out ? dist_out + ((int64_t)b * N + q) * k : nullptr;
            float* eo = edge_out ? edge_out + ((int64_t)b * N + q) * (int64_t)k * 6 : nullptr;
            const float* recf = reinterpret_cast<const float*>(s_rec);
            // rows of k int64 / 6k floats: 16-byte stores when k is even (row pitches 8k and 24k bytes are then 16-byte multiples)
            const bool vec = ((k & 1) == 0) && ((reinterpret_cast<uintptr_t>(idx_out) & 15) == 0) &&
                             (edge_out == nullptr || (reinterpret_cast<uintptr_t>(edge_out) & 15) == 0);
            float e6[12];
            long long nb2[2];
#pragma unroll
            for (int j = 0; j < K; ++j) {
                if (j < k) {
                    const unsigned low = (unsigned)(top.key[j] & 0xffffffffull);
                    // an unfilled slot (non-finite coordinates: no distance ever compares) stays in range
                    const unsigned p = min(low & 0xffffu, (unsigned)(M - 1));
                    const int nb = top.key[j] == kEmptyKey ? min(j, M - 1) : (int)(low >> 16);
                    if (dout) dout[j] = __uint_as_float((unsigned)(top.key[j] >> 32));
                    const int h = j & 1;
                    nb2[h] = nb;
                    if (eo) {
                        // self graph: the neighbour's coordinates are in the staged records (-2 x is exact, so is -0.5 * it)
                        const float* rc = recf + 16 * (p >> 2) + 8 * ((p & 3) >> 1) + (p & 1);
                        e6[6 * h + 0] = -0.5f * rc[0] - qx; e6[6 * h + 1] = -0.5f * rc[2] - qy; e6[6 * h + 2] = -0.5f * rc[4] - qz;
                        e6[6 * h + 3] = qx; e6[6 * h + 4] = qy; e6[6 * h + 5] = qz;
                    }
                    if (vec) {
                        if (h == 1) {
                            *reinterpret_cast<longlong2*>(io + j - 1) = make_longlong2(nb2[0], nb2[1]);
                            if (eo) {
                                float4* o4 = reinterpret_cast<float4*>(eo + 6 * (j - 1));
                                o4[0] = make_float4(e6[0], e6[1], e6[2], e6[3]);
                                o4[1] = make_float4(e6[4], e6[5], e6[6], e6[7]);
                                o4[2] = make_float4(e6[8], e6[9], e6[10], e6[11]);
                            }
                        }
                    } else {
                        io[j] = nb;
                        if (eo) {
#pragma unroll
                            for (int c = 0; c < 6; ++c) eo[6 * j + c] = e6[6 * h + c];
                        }
                    }
                }
            }
        }
        __syncwarp();
    }
}

}  // namespace ogmm

using namespace ogmm;

// Called by ogmm_knn_graph for C == 3, normalize == 0, 256 <= M <= 4096, N <= 4096, k <= 24.
// `stats` (optional, device int32[16], caller-zeroed): [0] warps that took the exhaustive redo path, [1] warps, [2] sweep
// steps, [3] steps that merged, [4] insert rounds, [5] sum over warps of the largest group count, [7] / [8] warps with
// too many prefix ties / fewer than k groups.
int ogmm_launch_knn3_select(const float* src, int64_t s_sb, int64_t s_sn, int64_t s_sc,
                            const float* dst, int64_t d_sb, int64_t d_sn, int64_t d_sc,
                            int64_t B, int64_t N, int64_t M, int64_t k,
                            int64_t* idx_out, float* dist_out, float* edge_out, int32_t* stats, cudaStream_t s) {
    const bool self = (src == dst) && s_sb == d_sb && s_sn == d_sn && s_sc == d_sc && N == M;
    const size_t smem = select_smem_bytes((int)N, (int)M, self);
    OGMM_REQUIRE(smem <= 200 * 1024, OGMM_EUNSUPPORTED, "knn select: %zu B of shared memory needed", smem);
    int dev = 0, sms = 0;
    int st = cuda_status(cudaGetDevice(&dev), "cudaGetDevice");
    if (st != OGMM_OK) return st;
    st = cuda_status(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev), "cudaDeviceGetAttribute");
    if (st != OGMM_OK) return st;
    const int64_t blocks = (N + kSwQueriesPerCta - 1) / kSwQueriesPerCta, items = B * blocks;
#define LAUNCH(KK)                                                                                                  \
    do {                                                                                                            \
        st = cuda_status(cudaFuncSetAttribute(knn3_select_kernel<KK>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                              (int)smem), "cudaFuncSetAttribute(knn3_select_kernel)");              \
        if (st != OGMM_OK) return st;                                                                               \
        int per_sm = 0;                                                                                             \
        st = cuda_status(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, knn3_select_kernel<KK>, kSwThreads, \
                                                                       smem), "cudaOccupancyMaxActiveBlocksPerMultiprocessor"); \
        if (st != OGMM_OK) return st;                                                                               \
        if (per_sm < 1) per_sm = 1;                                                                                 \
        int64_t grid = items < (int64_t)sms * per_sm ? items : (int64_t)sms * per_sm;                               \
        /* rounds every CTA needs anyway; if whole clouds (or even fractions of one) fit that count, give every CTA \
           such a range: each cloud is then sorted once per CTA that owns part of it instead of ~2.2 times */       \
        const int64_t per = (items + grid - 1) / grid;                                                              \
        int items_per_cta = 0;                                                                                      \
        if (per % blocks == 0 || blocks % per == 0) { items_per_cta = (int)per; grid = (items + per - 1) / per; }   \
        knn3_select_kernel<KK><<<(unsigned)grid, kSwThreads, smem, s>>>(src, s_sb, s_sn, s_sc, dst, d_sb, d_sn,     \
                                                                        d_sc, (int)B, (int)N, (int)M, (int)k,       \
                                                                        self ? 1 : 0, items_per_cta, idx_out,       \
                                                                        dist_out, edge_out, stats);                 \
    } while (0)
    if (k <= 4) LAUNCH(4);
    else if (k <= 8) LAUNCH(8);
    else if (k <= 16) LAUNCH(16);
    else if (k <= 20) LAUNCH(20);
    else LAUNCH(24);
#undef LAUNCH
    OGMM_LAUNCH_CHECK("knn3_select_kernel");
    return OGMM_OK;
}
