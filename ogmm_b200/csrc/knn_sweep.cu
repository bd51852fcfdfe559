// K1 (3-D, up to 4096 points per cloud): exact kNN by a sorted sweep with slab pruning (sm_100a).
//
// Same contract and the same distance arithmetic as knn3_kernel in knn.cu -- the reference's expanded form
// ((-2 s.d) + |s|^2) + |d|^2 with clamp 1e-12 (lib/utils.py:28-33), k smallest per row, ascending, ties to
// the lowest index (lib/utils.py:43) -- but a query no longer looks at every candidate:
//
//   1. the CTA sorts the cloud's candidates along the coordinate axis of largest extent (bitonic sort of
//      (key, index) pairs in shared memory) and stores them, in that order, as (-2x, -2y, -2z, |c|^2);
//   2. queries are taken in sorted order too, so the 32 queries of a warp are neighbours along the axis;
//   3. the warp sweeps outwards from its position, four candidates per side per step.  A side stops as soon
//      as, for EVERY lane, the squared axis gap to the next candidate exceeds that lane's current k-th best
//      distance (plus the fp32 error bound of the expanded form): since d^2 >= gap^2 and candidates are
//      sorted, nothing further on that side can enter any lane's list.  Typically ~25 % of the cloud is
//      visited, nearest slabs first, so the k-th-best threshold is tight almost immediately.
//
// The result is bit-identical to the exhaustive kernel: pruning only skips candidates that provably fail
// the (distance, index) comparison.  Because candidates no longer arrive in index order, list entries are
// 64-bit keys (distance bits << 32 | index) compared as integers -- lexicographic (distance, index) order in
// two instructions (distances are >= 1e-12 > 0, so their bit patterns order like the floats).
#include "knn_common.cuh"

namespace ogmm {

struct SweepSmem {
    float* ckey; int* cord; float4* c4; float* qkey; int* qord; float* stage_d; int* stage_i; float* red;
};
__host__ __device__ inline size_t sweep_smem_bytes(int N, int M, bool self) {
    const int Mp = pow2_ge(M), Np = self ? 0 : pow2_ge(N);
    return (size_t)16 * M + (size_t)8 * Mp + (size_t)8 * Np + (size_t)6 * kSwStage * kSwThreads + 256;
}

template <int K>
__global__ void __launch_bounds__(kSwThreads)
knn3_sweep_kernel(const float* __restrict__ src, int64_t s_sb, int64_t s_sn, int64_t s_sc,
                  const float* __restrict__ dst, int64_t d_sb, int64_t d_sn, int64_t d_sc,
                  int N, int M, int k, int self,
                  int64_t* __restrict__ idx_out, float* __restrict__ dist_out, float* __restrict__ edge_out) {
    extern __shared__ __align__(16) unsigned char sw_raw[];
    const int Mp = pow2_ge(M), Np = self ? 0 : pow2_ge(N);
    float4* s_c4 = reinterpret_cast<float4*>(sw_raw);
    float* s_ckey = reinterpret_cast<float*>(s_c4 + M);
    int* s_cord = reinterpret_cast<int*>(s_ckey + Mp);
    float* s_qkey = reinterpret_cast<float*>(s_cord + Mp);
    int* s_qord = reinterpret_cast<int*>(s_qkey + Np);
    float* s_stage_d = reinterpret_cast<float*>(s_qord + Np);
    unsigned short* s_stage_i = reinterpret_cast<unsigned short*>(s_stage_d + kSwStage * kSwThreads);
    float* s_red = reinterpret_cast<float*>(s_stage_i + kSwStage * kSwThreads);     // [64]

    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* sb = src + (int64_t)b * s_sb;
    const float* db = dst + (int64_t)b * d_sb;

    // ---- axis of largest extent, largest |c|^2 (for the error margin) ---------------------------------------
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
    float cn_max = 0.f;
    for (int w = 0; w < kSwThreads / 32; ++w) cn_max = fmaxf(cn_max, s_red[w * 8 + 6]);
    const int axis = (ext[0] >= ext[1] && ext[0] >= ext[2]) ? 0 : (ext[1] >= ext[2] ? 1 : 2);
    const int64_t d_ax = (int64_t)axis * d_sc, s_ax = (int64_t)axis * s_sc;

    // ---- sort candidates along the axis ----------------------------------------------------------------------
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
    for (int m = tid; m < M; m += kSwThreads) {
        const float* p = db + (int64_t)s_cord[m] * d_sn;
        const float x = p[0], y = p[d_sc], z = p[2 * d_sc];
        s_c4[m] = make_float4(-2.f * x, -2.f * y, -2.f * z, sqn3(x, y, z));
    }
    __syncthreads();
    const float* qkeys = self ? s_ckey : s_qkey;
    const int* qords = self ? s_cord : s_qord;

    // ---- queries, 256 sorted ranks at a time -------------------------------------------------------------------
    const int r_begin = blockIdx.x * kSwQueriesPerCta, r_end = min(N, r_begin + kSwQueriesPerCta);
    for (int r0 = r_begin; r0 < r_end; r0 += kSwThreads) {
        const int r = r0 + tid;
        const bool valid = r < r_end;
        int q = 0;
        float qx = 0.f, qy = 0.f, qz = 0.f, qk = 0.f;
        if (valid) {
            q = qords[r];
            const float* p = sb + (int64_t)q * s_sn;
            qx = p[0]; qy = p[s_sc]; qz = p[2 * s_sc];
            qk = qkeys[r];
        }
        const float qs = sqn3(qx, qy, qz);
        // fp32 error bound of the expanded form: the slab test must never prune a candidate that could still win
        const float margin = 1e-6f * (qs + cn_max) + 1e-30f;

        TopK64<K> top;
        top.init(s_stage_d, s_stage_i, tid, valid);

        // warp start: lower bound of the middle lane's key among the sorted candidates
        int pos;
        {
            const float mid = __shfl_sync(kFull, qk, 16);
            int lo_i = 0, hi_i = M;
            while (lo_i < hi_i) { const int md = (lo_i + hi_i) >> 1; if (s_ckey[md] < mid) lo_i = md + 1; else hi_i = md; }
            pos = lo_i;
        }
        if (!__any_sync(kFull, valid)) continue;          // whole warp idle (only in a ragged last round)
        int L = pos - 1, R = pos;
        while (true) {
            const float tl = top.thr * (1.0f + 1e-6f) + margin;
            bool go_l = false, go_r = false;
            if (L >= 0) { const float g = qk - s_ckey[L]; go_l = valid && !(g > 0.f && g * g > tl); }
            if (R < M) { const float g = s_ckey[R] - qk; go_r = valid && !(g > 0.f && g * g > tl); }
            go_l = __any_sync(kFull, go_l);
            go_r = __any_sync(kFull, go_r);
            if (!go_l && !go_r) break;
            if (go_l) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int p = L - u;
                    if (p >= 0) {
                        const float4 c = s_c4[p];
                        float v = __fmul_rn(qx, c.x);
                        v = fmaf(qy, c.y, v);
                        v = fmaf(qz, c.z, v);
                        v = fmaxf(__fadd_rn(__fadd_rn(v, qs), c.w), 1e-12f);
                        if (v <= top.thr) { top.sd[top.cnt * kSwThreads + tid] = v; top.si[top.cnt * kSwThreads + tid] = (unsigned short)s_cord[p]; ++top.cnt; }
                    }
                }
                L -= 4;
            }
            if (go_r) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int p = R + u;
                    if (p < M) {
                        const float4 c = s_c4[p];
                        float v = __fmul_rn(qx, c.x);
                        v = fmaf(qy, c.y, v);
                        v = fmaf(qz, c.z, v);
                        v = fmaxf(__fadd_rn(__fadd_rn(v, qs), c.w), 1e-12f);
                        if (v <= top.thr) { top.sd[top.cnt * kSwThreads + tid] = v; top.si[top.cnt * kSwThreads + tid] = (unsigned short)s_cord[p]; ++top.cnt; }
                    }
                }
                R += 4;
            }
            top.maybe_merge();
        }
        top.merge();

        if (valid) {
            int64_t* io = idx_out + ((int64_t)b * N + q) * k;
            float* dout = dist_out ? dist_out + ((int64_t)b * N + q) * k : nullptr;
            float* eo = edge_out ? edge_out + ((int64_t)b * N + q) * (int64_t)k * 6 : nullptr;
#pragma unroll
            for (int j = 0; j < K; ++j) {
                if (j < k) {
                    const int nb = (int)min((unsigned)(top.key[j] & 0xffffffffull), (unsigned)(M - 1));   // an unfilled slot (NaN keys) stays in range
                    io[j] = nb;
                    if (dout) dout[j] = __uint_as_float((unsigned)(top.key[j] >> 32));
                    if (eo) {
                        const float* p = sb + (int64_t)nb * s_sn;
                        eo[6 * j + 0] = p[0] - qx; eo[6 * j + 1] = p[s_sc] - qy; eo[6 * j + 2] = p[2 * s_sc] - qz;
                        eo[6 * j + 3] = qx; eo[6 * j + 4] = qy; eo[6 * j + 5] = qz;
                    }
                }
            }
        }
        __syncwarp();
    }
}

}  // namespace ogmm

using namespace ogmm;

// Called by ogmm_knn_graph for C == 3, normalize == 0, N, M <= 4096.  Returns OGMM_OK or an error code.
int ogmm_launch_knn3_sweep(const float* src, int64_t s_sb, int64_t s_sn, int64_t s_sc,
                           const float* dst, int64_t d_sb, int64_t d_sn, int64_t d_sc,
                           int64_t B, int64_t N, int64_t M, int64_t k,
                           int64_t* idx_out, float* dist_out, float* edge_out, cudaStream_t s) {
    const bool self = (src == dst) && s_sb == d_sb && s_sn == d_sn && s_sc == d_sc && N == M;
    const size_t smem = sweep_smem_bytes((int)N, (int)M, self);
    OGMM_REQUIRE(smem <= 200 * 1024, OGMM_EUNSUPPORTED, "knn sweep: %zu B of shared memory needed", smem);
    dim3 grid((unsigned)((N + kSwQueriesPerCta - 1) / kSwQueriesPerCta), (unsigned)B);
#define LAUNCH(KK)                                                                                                  \
    do {                                                                                                            \
        if (smem > 48 * 1024) {                                                                                     \
            int st = cuda_status(cudaFuncSetAttribute(knn3_sweep_kernel<KK>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                                      (int)smem), "cudaFuncSetAttribute(knn3_sweep_kernel)");       \
            if (st != OGMM_OK) return st;                                                                           \
        }                                                                                                           \
        knn3_sweep_kernel<KK><<<grid, kSwThreads, smem, s>>>(src, s_sb, s_sn, s_sc, dst, d_sb, d_sn, d_sc, (int)N,  \
                                                             (int)M, (int)k, self ? 1 : 0, idx_out, dist_out,       \
                                                             edge_out);                                             \
    } while (0)
    if (k <= 4) LAUNCH(4);
    else if (k <= 8) LAUNCH(8);
    else if (k <= 16) LAUNCH(16);
    else if (k <= 20) LAUNCH(20);
    else if (k <= 32) LAUNCH(32);
    else LAUNCH(64);
#undef LAUNCH
    OGMM_LAUNCH_CHECK("knn3_sweep_kernel");
    return OGMM_OK;
}
