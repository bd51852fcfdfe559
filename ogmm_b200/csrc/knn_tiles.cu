// K1 for LARGE 3-D self graphs (4096 < N = M <= 16384, k <= 24): exact kNN by a tiled sorted sweep (sm_100a).
//
// Same contract and the same distance arithmetic as the other 3-D kernels (lib/utils.py:12-44; ties to the lowest
// index).  The exhaustive kernel visits all N candidates per query; at N = 16384, k = 20 the k-th neighbour sits at
// ~0.1 of the cloud's extent, so along the widest axis only ~1/5 of the points can matter.  Two kernels:
//
//   knn3_presort_kernel      one CTA per cloud: widest axis, counting sort into 4096 axis bins in shared memory, then the
//                            cloud is written to the workspace IN THAT ORDER as packed group records (knn_common.cuh),
//                            with the axis keys, the position -> original-index map and monotone per-tile key bounds.
//   knn3_tile_sweep_kernel   one CTA per 256 consecutive positions = queries (neighbours in space along the axis).  The CTA
//                            walks the sorted cloud in tiles of 256 candidates outwards from its own tile, left and
//                            right alternately, each tile staged in shared memory and consumed like the exhaustive
//                            kernel's (packed FP32x2 distances, eight candidates per warp vote, sorted 64-bit
//                            (distance, index) lists).  A side is abandoned as soon as NO query of the CTA can still be
//                            improved there: axis gap to the tile's nearest key, squared, above the query's current
//                            k-th distance plus the fp32 error margin of the expanded form.
//
// Exactness: a candidate is skipped only if (x_c - x_q)^2 alone already exceeds the query's k-th distance (plus margin),
// and squared distances only grow with the other two axes; lists break ties by ORIGINAL index, so the visiting order does
// not matter.
#include "knn_common.cuh"

namespace ogmm {

constexpr int kTsTile = 256;                   // candidates per tile (64 group records, 4 KB)
constexpr float kTsBig = 1.0e38f;              // |c|^2 of padding positions: finite, above any real distance

// workspace per cloud (floats): rec [Mp * 4] | key [Mp] | ord [Mp] (int) | meta [4] | hi_bound [kTsMaxTiles] | lo_bound [kTsMaxTiles]
constexpr int kTsMaxTiles = 16384 / kTsTile;
__host__ __device__ inline size_t tile_ws_floats(int Mp) { return (size_t)Mp * 6 + 4 + 2 * kTsMaxTiles; }

constexpr int kPsThreads = 1024;                // the pre-pass is one CTA per cloud: as many threads as a CTA can have
constexpr int kPsBins = 4096;                   // axis bins of the counting sort (+ one bin for non-finite keys)

// The sweep needs the cloud ORDERED BY TILE along the axis, not sorted inside a tile: a counting sort into 4096 equal-width
// axis bins (two passes of shared-memory atomics and one block scan) replaces a 105-stage bitonic sort of 16384 keys.
// Positions are non-decreasing in bin, arbitrary (atomic arrival order) inside a bin; the sweep never relies on more:
// per tile it gets hi_bound[t] = the largest key at or before tile t and lo_bound[t] = the smallest key at or after it,
// both monotone in t, as the exact pruning bounds.  The results do not depend on the order inside a bin (every reachable
// tile is visited whole, ties break by original index).
__global__ void __launch_bounds__(kPsThreads)
knn3_presort_kernel(const float* __restrict__ dst, int64_t d_sb, int64_t d_sn, int64_t d_sc, int M, int Mp,
                    float* __restrict__ ws) {
    extern __shared__ __align__(16) unsigned char ps_raw[];
    float* s_key = reinterpret_cast<float*>(ps_raw);               // [Mp] keys in output order
    int* s_ord = reinterpret_cast<int*>(s_key + Mp);               // [Mp] output position -> original index
    int* s_hist = s_ord + Mp;                                      // [kPsBins + 2] counts, then running cursors
    __shared__ float s_red[kPsThreads / 32 * 8];
    __shared__ int s_scan[kPsThreads / 32];
    __shared__ float s_tmax[kTsMaxTiles], s_tmin[kTsMaxTiles];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* db = dst + (int64_t)b * d_sb;
    float* w = ws + (size_t)b * tile_ws_floats(Mp);
    float* w_rec = w;
    float* w_key = w + (size_t)Mp * 4;
    int* w_ord = reinterpret_cast<int*>(w_key + Mp);
    float* w_meta = reinterpret_cast<float*>(w_ord + Mp);
    float* w_hi = w_meta + 4;
    float* w_lo = w_hi + kTsMaxTiles;

    // ---- extents (finite coordinates only), largest |c|^2 ----
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY}, nmax = 0.f;
    for (int m = tid; m < M; m += kPsThreads) {
        const float* p = db + (int64_t)m * d_sn;
        const float x = p[0], y = p[d_sc], z = p[2 * d_sc];
        if (fabsf(x) < INFINITY) { lo[0] = fminf(lo[0], x); hi[0] = fmaxf(hi[0], x); }
        if (fabsf(y) < INFINITY) { lo[1] = fminf(lo[1], y); hi[1] = fmaxf(hi[1], y); }
        if (fabsf(z) < INFINITY) { lo[2] = fminf(lo[2], z); hi[2] = fmaxf(hi[2], z); }
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
    for (int i = tid; i < kPsBins + 2; i += kPsThreads) s_hist[i] = 0;
    __syncthreads();
    float ext[3], alo[3], cn_max = 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float l = INFINITY, h = -INFINITY;
        for (int ww = 0; ww < kPsThreads / 32; ++ww) { l = fminf(l, s_red[ww * 8 + a]); h = fmaxf(h, s_red[ww * 8 + 3 + a]); }
        ext[a] = h - l; alo[a] = l;
        if (!(ext[a] >= 0.f)) ext[a] = 0.f;                        // no finite coordinate on this axis
    }
    for (int ww = 0; ww < kPsThreads / 32; ++ww) cn_max = fmaxf(cn_max, s_red[ww * 8 + 6]);
    const int axis = (ext[0] >= ext[1] && ext[0] >= ext[2]) ? 0 : (ext[1] >= ext[2] ? 1 : 2);
    const float k_lo = alo[axis];
    const float k_scale = ext[axis] > 0.f ? (float)kPsBins / ext[axis] : 0.f;
    auto key_of = [&](int m) { return sort_key(db[(int64_t)m * d_sn + (int64_t)axis * d_sc]); };
    auto bin_of = [&](float key) {                                  // monotone in key; non-finite keys (NaN -> +inf, -inf) at the ends
        if (!(key < INFINITY)) return kPsBins;
        if (!(key > -INFINITY)) return 0;
        const int bi = (int)((key - k_lo) * k_scale);
        return bi < 0 ? 0 : (bi > kPsBins - 1 ? kPsBins - 1 : bi);
    };
    // ---- histogram, exclusive scan, scatter ----
    for (int m = tid; m < M; m += kPsThreads) atomicAdd(&s_hist[bin_of(key_of(m))], 1);
    __syncthreads();
    {
        constexpr int per = (kPsBins + 1 + kPsThreads - 1) / kPsThreads;      // 5 bins per thread
        int local[per], sum = 0;
#pragma unroll
        for (int i = 0; i < per; ++i) { const int bi = tid * per + i; local[i] = bi <= kPsBins ? s_hist[bi] : 0; sum += local[i]; }
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(kFull, incl, o); if (lane >= o) incl += v; }
        if (lane == 31) s_scan[warp] = incl;
        __syncthreads();
        int base = 0;
        for (int ww = 0; ww < warp; ++ww) base += s_scan[ww];
        int run = base + incl - sum;
#pragma unroll
        for (int i = 0; i < per; ++i) { const int bi = tid * per + i; if (bi <= kPsBins) s_hist[bi] = run; run += local[i]; }
    }
    __syncthreads();
    for (int m = tid; m < M; m += kPsThreads) {
        const float key = key_of(m);
        const int pos = atomicAdd(&s_hist[bin_of(key)], 1);
        s_key[pos] = key; s_ord[pos] = m;
    }
    for (int p = M + tid; p < Mp; p += kPsThreads) { s_key[p] = INFINITY; s_ord[p] = 0; }
    __syncthreads();
    // ---- the cloud in output order: packed group records, keys, index map ----
    for (int p = tid; p < Mp; p += kPsThreads) {
        float x = 0.f, y = 0.f, z = 0.f, wv = kTsBig;
        if (p < M) {
            const float* c = db + (int64_t)s_ord[p] * d_sn;
            const float cx = c[0], cy = c[d_sc], cz = c[2 * d_sc];
            x = -2.f * cx; y = -2.f * cy; z = -2.f * cz; wv = sqn3(cx, cy, cz);
        }
        float* r = w_rec + 16 * (size_t)(p >> 2) + 8 * ((p & 3) >> 1) + (p & 1);
        r[0] = x; r[2] = y; r[4] = z; r[6] = wv;
        w_key[p] = s_key[p];
        w_ord[p] = s_ord[p];
    }
    // ---- per-tile key range, then the monotone bounds ----
    const int n_tiles = Mp / kTsTile;
    for (int t = warp; t < n_tiles; t += kPsThreads / 32) {
        float mx = -INFINITY, mn = INFINITY;
        for (int e = lane; e < kTsTile; e += 32) { const float kk = s_key[t * kTsTile + e]; mx = fmaxf(mx, kk); mn = fminf(mn, kk); }
        mx = warp_max(mx); mn = -warp_max(-mn);
        if (lane == 0) { s_tmax[t] = mx; s_tmin[t] = mn; }
    }
    __syncthreads();
    if (tid == 0) {
        float run = -INFINITY;
        for (int t = 0; t < n_tiles; ++t) { run = fmaxf(run, s_tmax[t]); w_hi[t] = run; }
        run = INFINITY;
        for (int t = n_tiles - 1; t >= 0; --t) { run = fminf(run, s_tmin[t]); w_lo[t] = run; }
        w_meta[0] = cn_max; w_meta[1] = (float)axis;
    }
}

template <int K>
__global__ void __launch_bounds__(kSwThreads, 2)
knn3_tile_sweep_kernel(const float* __restrict__ src, int64_t s_sb, int64_t s_sn, int64_t s_sc, int M, int Mp, int k,
                       const float* __restrict__ ws, int64_t* __restrict__ idx_out, float* __restrict__ dist_out,
                       float* __restrict__ edge_out) {
    __shared__ __align__(16) float s_rec[4 * kTsTile];
    __shared__ unsigned short s_ord[kTsTile];
    __shared__ float s_stage_d[kSwStage * kSwThreads];
    __shared__ unsigned short s_stage_i[kSwStage * kSwThreads];
    const int b = blockIdx.y, tid = threadIdx.x;
    const float* sb = src + (int64_t)b * s_sb;
    const float* w = ws + (size_t)b * tile_ws_floats(Mp);
    const float* w_rec = w;
    const float* w_key = w + (size_t)Mp * 4;
    const int* w_ord = reinterpret_cast<const int*>(w_key + Mp);
    const float* w_meta = reinterpret_cast<const float*>(w_ord + Mp);
    const float cn_max = w_meta[0];
    const float* w_hi = w_meta + 4;                         // largest key at or before tile t
    const float* w_lo = w_hi + kTsMaxTiles;                 // smallest key at or after tile t

    const int p0 = blockIdx.x * kSwThreads;             // first sorted position of this CTA's queries
    const int p = p0 + tid;
    const bool valid = p < M;
    int q = 0;
    float qx = 0.f, qy = 0.f, qz = 0.f, qk = 0.f;
    if (valid) {
        q = w_ord[p];
        const float* pp = sb + (int64_t)q * s_sn;
        qx = pp[0]; qy = pp[s_sc]; qz = pp[2 * s_sc];
        qk = w_key[p];
    }
    const float qs = sqn3(qx, qy, qz);
    QueryPack Q;
    Q.x = pack2(qx, qx); Q.y = pack2(qy, qy); Q.z = pack2(qz, qz); Q.s = pack2(qs, qs);
    const float margin = 1e-6f * (qs + cn_max) + 1e-30f;   // fp32 error of the expanded form: never prune a possible winner
    const unsigned rec_base = (unsigned)__cvta_generic_to_shared(s_rec);

    TopK64<K> top;
    top.init(s_stage_d, s_stage_i, tid, valid);

    auto process = [&](int t, bool mine) {                  // all threads; tile t of the sorted cloud; `mine`: this lane needs it
        __syncthreads();                                    // previous tile fully consumed
        const float4* g_rec = reinterpret_cast<const float4*>(w_rec + (size_t)t * kTsTile * 4);
        float4* s4 = reinterpret_cast<float4*>(s_rec);
        for (int e = tid; e < kTsTile; e += kSwThreads) s4[e] = __ldg(g_rec + e);
        for (int e = tid; e < kTsTile; e += kSwThreads) s_ord[e] = (unsigned short)w_ord[(size_t)t * kTsTile + e];
        __syncthreads();
        // a warp whose 32 queries (neighbours along the axis) are all out of reach of this tile sits it out
        if (!__any_sync(kFull, mine)) return;
        for (int m = 0; m < kTsTile; m += 8) {
            float v[4], u[4];
            group_distances(rec_base + 16u * (unsigned)m, Q, v);
            group_distances(rec_base + 16u * (unsigned)m + 64u, Q, u);
#pragma unroll
            for (int i = 0; i < 4; ++i) { v[i] = fmaxf(v[i], 1e-12f); u[i] = fmaxf(u[i], 1e-12f); }
            const float mn = fminf(fminf(fminf(v[0], v[1]), fminf(v[2], v[3])), fminf(fminf(u[0], u[1]), fminf(u[2], u[3])));
            if (__any_sync(kFull, mn <= top.thr)) {
#pragma unroll
                for (int i = 0; i < 4; ++i) top.offer(v[i], s_ord[m + i]);
                top.maybe_merge();
#pragma unroll
                for (int i = 0; i < 4; ++i) top.offer(u[i], s_ord[m + 4 + i]);
                top.maybe_merge();
            }
        }
        top.merge();                                        // thr = the exact k-th distance so far
    };

    const int n_tiles = Mp / kTsTile;
    const int c = p0 / kTsTile;
    process(c, true);
    int l = c - 1, r = c + 1;
    while (true) {
        const float tl = top.thr * (1.0f + 1e-6f) + margin;
        bool need_l = false, need_r = false;
        if (l >= 0) { const float g = qk - w_hi[l]; need_l = valid && !(g > 0.f && g * g > tl); }
        if (r < n_tiles) { const float g = w_lo[r] - qk; need_r = valid && !(g > 0.f && g * g > tl); }
        const bool go_l = __syncthreads_or(need_l ? 1 : 0) != 0;      // (the builtin returns "any", not the OR of the values)
        const bool go_r = __syncthreads_or(need_r ? 1 : 0) != 0;
        if (!go_l && !go_r) break;
        if (go_l) { process(l, need_l); --l; } else l = -1; // a side nobody needs any more stays closed (thr only shrinks)
        if (go_r) {                                         // (the left tile may have tightened thr: test again, cheaply)
            if (go_l && need_r) { const float g = w_lo[r] - qk; const float t2 = top.thr * (1.0f + 1e-6f) + margin; need_r = !(g > 0.f && g * g > t2); }
            process(r, need_r); ++r;
        } else r = n_tiles;
    }

    if (!valid) return;
    int64_t* io = idx_out + ((int64_t)b * M + q) * k;
    float* dout = dist_out ? dist_out + ((int64_t)b * M + q) * k : nullptr;
    float* eo = edge_out ? edge_out + ((int64_t)b * M + q) * (int64_t)k * 6 : nullptr;
#pragma unroll
    for (int j = 0; j < K; ++j) {
        if (j < k) {
            const unsigned low = (unsigned)(top.key[j] & 0xffffffffull);
            const int nb = top.key[j] == kEmptyKey ? min(j, M - 1) : (int)low;       // unfilled slots (NaN rows): in range
            io[j] = nb;
            if (dout) dout[j] = __uint_as_float((unsigned)(top.key[j] >> 32));
            if (eo) {
                const float* pp = sb + (int64_t)nb * s_sn;
                eo[6 * j + 0] = pp[0] - qx; eo[6 * j + 1] = pp[s_sc] - qy; eo[6 * j + 2] = pp[2 * s_sc] - qz;
                eo[6 * j + 3] = qx; eo[6 * j + 4] = qy; eo[6 * j + 5] = qz;
            }
        }
    }
}

}  // namespace ogmm

using namespace ogmm;

// Host entry for knn.cu.  Returns OGMM_EUNSUPPORTED when the call is outside this path's range (the caller falls back).
int ogmm_launch_knn3_tiles(const float* src, int64_t s_sb, int64_t s_sn, int64_t s_sc, int64_t B, int64_t M, int64_t k,
                           int64_t* idx_out, float* dist_out, float* edge_out, cudaStream_t s) {
    if (M <= 4096 || M > 16384 || k > 24) return OGMM_EUNSUPPORTED;      // the pre-sort holds 8 bytes per point in shared memory
    int Mp = 8192;
    while (Mp < M) Mp <<= 1;
    const size_t ws_bytes = sizeof(float) * tile_ws_floats(Mp) * (size_t)B;
    float* ws = nullptr;
    int st = cuda_status(cudaMallocAsync(reinterpret_cast<void**>(&ws), ws_bytes, s), "cudaMallocAsync(knn tiles workspace)");
    if (st != OGMM_OK) return st;
    const size_t sort_smem = (size_t)Mp * 8 + sizeof(int) * (kPsBins + 2);
    st = cuda_status(cudaFuncSetAttribute(knn3_presort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sort_smem),
                     "cudaFuncSetAttribute(knn3_presort_kernel)");
    if (st == OGMM_OK) {
        knn3_presort_kernel<<<(unsigned)B, kPsThreads, sort_smem, s>>>(src, s_sb, s_sn, s_sc, (int)M, Mp, ws);
        st = cuda_status(cudaGetLastError(), "knn3_presort_kernel");
    }
    if (st == OGMM_OK) {
        dim3 grid((unsigned)((M + kSwThreads - 1) / kSwThreads), (unsigned)B);
#define LAUNCH(KK) knn3_tile_sweep_kernel<KK><<<grid, kSwThreads, 0, s>>>(src, s_sb, s_sn, s_sc, (int)M, Mp, (int)k, ws, idx_out, dist_out, edge_out)
        if (k <= 4) LAUNCH(4);
        else if (k <= 8) LAUNCH(8);
        else if (k <= 16) LAUNCH(16);
        else if (k <= 20) LAUNCH(20);
        else LAUNCH(24);
#undef LAUNCH
        st = cuda_status(cudaGetLastError(), "knn3_tile_sweep_kernel");
    }
    const int st2 = cuda_status(cudaFreeAsync(ws, s), "cudaFreeAsync(knn tiles workspace)");
    return st != OGMM_OK ? st : st2;
}
