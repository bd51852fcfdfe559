// K3, register-operand tensor-core form: out[j][d] = sum_n gamma[n][j] * f[d][n] / npi[j] for J == 16 and features in
// their native (B,D,N) layout.  The feature matrix is the only large operand (B*D*N*4 bytes, read once); this kernel
// takes it from HBM straight into mma.sync A fragments -- no shared-memory staging, no transposition, no block-level
// synchronisation in the main loop -- so the SM's work per feature byte is a 16-byte load, two ALU ops and 3/8 of an
// MMA, and the kernel runs at whatever HBM delivers.
//
//   * Warp tile: 32 feature rows (2 m16 tiles) x all 16 clusters (2 n8 tiles), K = points.  A step covers 32 points:
//     lane (g = lane / 4, t = lane % 4) loads the float4s f[row g (+8)][n0 + 16 h + 4t .. + 3], h = 0, 1, of each m16
//     tile back to back -- one full 128-byte line per row per step, every sector used.  K is a summation index, so
//     the fragment's k slots may hold any permutation of the points as long as B uses the same one: k-step ks (0..3)
//     takes points 16 (ks / 2) + 4t + 2 (ks % 2) (slot k = t) and the next one (slot k = t + 4) straight from the
//     loaded float4.  A tile's registers are refilled for the warp's next step as soon as they are consumed, so each
//     lane keeps 4 to 8 16-byte loads in flight (16 warps per SM: 32 to 64 KB) with no block-level synchronisation.
//   * FP32 fidelity through the error-compensated TF32 split (reference accumulates in FP32; parity budget 1e-4):
//     x = hi + lo, hi = x with the low 13 mantissa bits cleared, lo = x - hi (exact); acc += hi*hi + lo*hi + hi*lo.
//     The dropped lo*lo term and the truncation of lo are ~2^-21 relative; accumulation is FP32.
//   * gamma (N x 16) is staged once per CTA in shared memory, 1024 points at a time, as [n / 4][j'][n % 4] with
//     j' = (j + 2 (n / 4 % 4)) % 8 + (j & 8): the B fragment of a step is two conflict-free 128-bit loads per lane
//     (clusters g and g + 8, four consecutive points each) per 16 points, reused by both m16 tiles.
//   * CTA = 8 warps laid out WR (rows) x WN (point interleave), WR = min(8, ceil(D / 32)): D >= 256 gives each warp
//     its own 32 rows over every point; smaller D splits the points of a chunk across warps step by step.  Partial
//     sums meet in shared memory after the loop; npi comes from the gamma staging pass.
//
// Taken by ogmm_gmm_moments_feat for J == 16, N % 4 == 0, D % 32 == 0, native layout, contiguous gamma, 16-byte aligned rows
// (OGMM_FEAT_NO_MMA=1 disables it); everything else runs the FP32 FFMA2 kernel in moments.cu.
#include "common.cuh"

namespace ogmm {

constexpr int kMmThreads = 256;
constexpr int kMmChunk = 1024;                    // points of gamma resident in shared memory
constexpr int kMmRowsWarp = 32;
constexpr int kMmRedPitch = 17;
constexpr size_t kMmSmem = sizeof(float) * ((size_t)kMmChunk * 16 + 8 * 16 + 16);

__device__ __forceinline__ uint32_t tf32_hi(float x) { return __float_as_uint(x) & 0xffffe000u; }
__device__ __forceinline__ uint32_t tf32_lo(float x, uint32_t hi) { return __float_as_uint(x - __uint_as_float(hi)); }

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t b0, const uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ float4 ldg_nc4(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// float offset of gamma[n][j] inside the staged chunk (n relative to the chunk)
__device__ __forceinline__ int gpos(int n, int j) {
    const int q = n >> 2;
    return q * 64 + ((((j + 2 * (q & 3)) & 7) + (j & 8)) << 2) + (n & 3);
}

__global__ void __launch_bounds__(kMmThreads, 2)
gmm_moments_feat_mma_kernel(const float* __restrict__ gamma, const float* __restrict__ feats, int64_t f_sb, int64_t f_sd,
                            int N, int D, int rows_cta, float* __restrict__ pi_out, float* __restrict__ mu_out) {
    extern __shared__ __align__(16) float mm_smem[];
    float* s_g = mm_smem;                            // [kMmChunk / 4][16][4] swizzled gamma; later the fold buffer
    float* s_gs = mm_smem + kMmChunk * 16;           // [8][16] gamma column sums per warp
    float* s_npi = s_gs + 8 * 16;                    // [16]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int b = blockIdx.y, d0 = blockIdx.x * rows_cta;
    const int WR = rows_cta / kMmRowsWarp, WN = 8 / WR;
    const int wr = warp % WR, wn = warp / WR;
    const float* gam = gamma + (int64_t)b * N * 16;

    // this lane's rows: d0 + 32 wr + 8 m + g, m = 0..3 (tile mt = m / 2, fragment row half m % 2); D % 32 == 0, so a warp
    // is either fully inside the matrix or fully outside (then it re-reads row 0 and its sums are dropped)
    const bool warp_live = d0 + kMmRowsWarp * wr < D;
    const float* p0 = feats + (int64_t)b * f_sb + (warp_live ? (int64_t)(d0 + kMmRowsWarp * wr + g) * f_sd : 0) + 4 * t;
    const int64_t stride8 = warp_live ? 8 * f_sd : 0;

    float acc[2][2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int jt = 0; jt < 2; ++jt)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[mt][jt][e] = 0.f;
    float4 gsum = make_float4(0.f, 0.f, 0.f, 0.f);   // clusters 4 (tid & 3) .. + 3 over this thread's staged points

    // a[m][h]: row m of this lane, points n0 + 16 h + 4 t .. + 3 of the current 32-point step
    float4 a[4][2];
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    auto load_row = [&](int m, int n0) {
        const float* p = p0 + m * stride8 + n0;
        a[m][0] = (n0 + 4 * t < N) ? ldg_nc4(p) : zero4;
        a[m][1] = (n0 + 16 + 4 * t < N) ? ldg_nc4(p + 16) : zero4;
    };

    for (int c0 = 0; c0 < N; c0 += kMmChunk) {
        const int clen = min(kMmChunk, N - c0);
        const int csteps = (clen + 31) >> 5;
        // first feature step of this warp in flight while gamma is staged
        if (wn < csteps) {
#pragma unroll
            for (int m = 0; m < 4; ++m) load_row(m, c0 + 32 * wn);
        }
        __syncthreads();                             // previous chunk fully consumed
        const int cpad = csteps << 5;
#pragma unroll 4
        for (int u = tid; u < cpad * 4; u += kMmThreads) {   // u = n * 4 + q: clusters 4q .. 4q + 3 of point n
            const int n = u >> 2, q = u & 3;
            float4 v = zero4;
            if (n < clen) v = *reinterpret_cast<const float4*>(gam + (int64_t)(c0 + n) * 16 + 4 * q);
            gsum.x += v.x; gsum.y += v.y; gsum.z += v.z; gsum.w += v.w;
            s_g[gpos(n, 4 * q)] = v.x; s_g[gpos(n, 4 * q + 1)] = v.y;
            s_g[gpos(n, 4 * q + 2)] = v.z; s_g[gpos(n, 4 * q + 3)] = v.w;
        }
        __syncthreads();

        for (int s = wn; s < csteps; s += WN) {
            // B of the step: bq[h][jt] = gamma[points 32 s + 16 h + 4 t .. + 3][cluster 8 jt + g]
            const float* gq = s_g + (8 * s + t) * 64 + (((g + 2 * t) & 7) << 2);
            float4 bq[2][2];
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int jt = 0; jt < 2; ++jt) bq[h][jt] = *reinterpret_cast<const float4*>(gq + h * 256 + jt * 32);
            const int n_next = c0 + 32 * (s + WN);
            const bool more = s + WN < csteps;
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                const float4 x[2][2] = {{a[2 * mt][0], a[2 * mt][1]}, {a[2 * mt + 1][0], a[2 * mt + 1][1]}};   // [row half][h]
                if (more) { load_row(2 * mt, n_next); load_row(2 * mt + 1, n_next); }   // refill for this warp's next step
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const int h = ks >> 1;
                    // k slots t and t + 4 of k-step ks <- points 16 h + 4 t + 2 (ks & 1) and + 1
                    float av[4], bv[2][2];
                    if (ks & 1) {
                        av[0] = x[0][h].z; av[1] = x[1][h].z; av[2] = x[0][h].w; av[3] = x[1][h].w;
                        bv[0][0] = bq[h][0].z; bv[0][1] = bq[h][0].w; bv[1][0] = bq[h][1].z; bv[1][1] = bq[h][1].w;
                    } else {
                        av[0] = x[0][h].x; av[1] = x[1][h].x; av[2] = x[0][h].y; av[3] = x[1][h].y;
                        bv[0][0] = bq[h][0].x; bv[0][1] = bq[h][0].y; bv[1][0] = bq[h][1].x; bv[1][1] = bq[h][1].y;
                    }
                    uint32_t ah[4], al[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) { ah[e] = tf32_hi(av[e]); al[e] = tf32_lo(av[e], ah[e]); }
#pragma unroll
                    for (int jt = 0; jt < 2; ++jt) {
                        const uint32_t bh0 = tf32_hi(bv[jt][0]), bh1 = tf32_hi(bv[jt][1]);
                        const uint32_t bl0 = tf32_lo(bv[jt][0], bh0), bl1 = tf32_lo(bv[jt][1], bh1);
                        mma_tf32(acc[mt][jt], al, bh0, bh1);
                        mma_tf32(acc[mt][jt], ah, bl0, bl1);
                        mma_tf32(acc[mt][jt], ah, bh0, bh1);
                    }
                }
            }
        }
    }

    // ---- gamma column sums -> npi -------------------------------------------------------------------------
    gsum.x += __shfl_xor_sync(kFull, gsum.x, 4); gsum.y += __shfl_xor_sync(kFull, gsum.y, 4);
    gsum.z += __shfl_xor_sync(kFull, gsum.z, 4); gsum.w += __shfl_xor_sync(kFull, gsum.w, 4);
    gsum.x += __shfl_xor_sync(kFull, gsum.x, 8); gsum.y += __shfl_xor_sync(kFull, gsum.y, 8);
    gsum.z += __shfl_xor_sync(kFull, gsum.z, 8); gsum.w += __shfl_xor_sync(kFull, gsum.w, 8);
    gsum.x += __shfl_xor_sync(kFull, gsum.x, 16); gsum.y += __shfl_xor_sync(kFull, gsum.y, 16);
    gsum.z += __shfl_xor_sync(kFull, gsum.z, 16); gsum.w += __shfl_xor_sync(kFull, gsum.w, 16);
    if (lane < 4) *reinterpret_cast<float4*>(s_gs + warp * 16 + 4 * lane) = gsum;
    __syncthreads();                                 // also: every warp is done with s_g
    if (tid < 16) {
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) tot += s_gs[w * 16 + tid];
        const float pi = __fdiv_rn(tot, (float)N);
        s_npi[tid] = __fadd_rn(__fmul_rn(pi, (float)N), 1e-5f);
        if (blockIdx.x == 0 && pi_out) pi_out[(int64_t)b * 16 + tid] = pi;
    }
    // ---- partial sums of the WN point groups meet in shared memory: s_red[wn][row][j], pitch 17 ------------------
    float* s_red = s_g;                              // 8 warps x 32 rows x 17 floats = 17,408 B <= 64 KB
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int jt = 0; jt < 2; ++jt) {
            const int row = wn * rows_cta + kMmRowsWarp * wr + 16 * mt + g, j = 8 * jt + 2 * t;
            s_red[row * kMmRedPitch + j] = acc[mt][jt][0];
            s_red[row * kMmRedPitch + j + 1] = acc[mt][jt][1];
            s_red[(row + 8) * kMmRedPitch + j] = acc[mt][jt][2];
            s_red[(row + 8) * kMmRedPitch + j + 1] = acc[mt][jt][3];
        }
    __syncthreads();
    for (int e = tid; e < 16 * rows_cta; e += kMmThreads) {
        const int j = e / rows_cta, row = e - j * rows_cta;
        if (d0 + row < D) {
            float v = 0.f;
            for (int w = 0; w < WN; ++w) v += s_red[(w * rows_cta + row) * kMmRedPitch + j];
            mu_out[((int64_t)b * 16 + j) * D + d0 + row] = __fdiv_rn(v, s_npi[j]);
        }
    }
}

}  // namespace ogmm

using namespace ogmm;

// Returns OGMM_EUNSUPPORTED (without touching the error string) when the call does not fit; the caller falls back.
int ogmm_launch_moments_feat_mma(const float* gamma, int64_t g_sb, int64_t g_sn, int64_t g_sj,
                                 const float* feats, int64_t f_sb, int64_t f_sn, int64_t f_sd,
                                 int64_t B, int64_t N, int64_t J, int64_t D, float* pi_out, float* mu_out, cudaStream_t s) {
    if (J != 16 || (N & 3) != 0 || N < 32 || (D % kMmRowsWarp) != 0 || f_sn != 1 || (f_sd & 3) != 0 || (f_sb & 3) != 0 || g_sj != 1 || g_sn != 16 ||
        g_sb != N * 16 || (reinterpret_cast<uintptr_t>(feats) & 15) != 0 || (reinterpret_cast<uintptr_t>(gamma) & 15) != 0)
        return OGMM_EUNSUPPORTED;
    static bool configured = false;
    if (!configured) {
        int st = cuda_status(cudaFuncSetAttribute(gmm_moments_feat_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                  (int)kMmSmem), "cudaFuncSetAttribute(gmm_moments_feat_mma_kernel)");
        if (st != OGMM_OK) return st;
        configured = true;
    }
    int wr = (int)((D + kMmRowsWarp - 1) / kMmRowsWarp);
    wr = wr >= 8 ? 8 : (wr >= 4 ? 4 : (wr >= 2 ? 2 : 1));
    const int rows_cta = wr * kMmRowsWarp;
    dim3 grid((unsigned)((D + rows_cta - 1) / rows_cta), (unsigned)B);
    gmm_moments_feat_mma_kernel<<<grid, kMmThreads, kMmSmem, s>>>(gamma, feats, f_sb, f_sd, (int)N, (int)D, rows_cta, pi_out,
                                                                  mu_out);
    OGMM_LAUNCH_CHECK("gmm_moments_feat_mma_kernel");
    return OGMM_OK;
}
