// K3 backward: gradient of the feature M-step with respect to the point features (sm_100a).
//
// Forward (lib/utils.py:137-140, called with wide features at :289): pi = mean_n gamma, npi = pi N + 1e-5,
// mu[b,j,d] = sum_n gamma[b,n,j] feats[b,n,d] / npi[b,j].  In the reference's training step (train.py:57-75) gamma is
// detached (lib/utils.py:286) and autograd flows into `feats` only:
//
//     dL/dfeats[b,n,d] = sum_j gamma[b,n,j] * (dL/dmu[b,j,d] / npi[b,j])
//
// a (N x J) x (J x D) product per cloud whose output is as large as the feature tensor: 4 (N J + J D + N D) bytes per
// cloud, HBM WRITE bound (2 MB out of 2.1 MB at N=1024, J=16, D=512), 2 J flop per output element.
//
// CTA tile = 128 points x 64 feature rows, 256 threads; thread (tx, ty) owns 4 consecutive points x 8 rows.  gamma is
// staged transposed ([j][point]: the 4 points of a thread are one conflict-free 16-byte read), dL/dmu / npi is staged
// as [j][row] (warp-uniform reads).  Accumulation is FP32, packed FFMA2 over point pairs, j ascending -- the same
// summation order for every output element, so the result does not depend on the launch geometry.  Stores are
// 16-byte, 512 contiguous bytes per warp and row when the output is in the model's native (B,D,N) layout.
#include "common.cuh"

namespace ogmm {

constexpr int kBwdPts = 128, kBwdRows = 64, kBwdThreads = 256, kBwdJ = 16;

__global__ void __launch_bounds__(kBwdThreads)
gmm_moments_feat_bwd_kernel(const float* __restrict__ gamma, int64_t g_sb, int64_t g_sn, int64_t g_sj,
                            const float* __restrict__ dmu, const float* __restrict__ pi,
                            int N, int J, int D,
                            float* __restrict__ dfeats, int64_t o_sb, int64_t o_sn, int64_t o_sd) {
    __shared__ __align__(16) float s_g[kBwdJ][kBwdPts + 4];      // gamma chunk, transposed (pitch 132: the transposing store is 2-way conflicted at worst)
    __shared__ __align__(16) float s_s[kBwdJ][kBwdRows];         // dL/dmu / npi chunk
    const int b = blockIdx.z, n0 = blockIdx.x * kBwdPts, d0 = blockIdx.y * kBwdRows;
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const float* gb = gamma + (int64_t)b * g_sb;
    const float* mb = dmu + (int64_t)b * J * D;
    const float* pb = pi + (int64_t)b * J;

    float2 acc[8][2];
#pragma unroll
    for (int r = 0; r < 8; ++r) { acc[r][0] = make_float2(0.f, 0.f); acc[r][1] = make_float2(0.f, 0.f); }

    for (int j0 = 0; j0 < J; j0 += kBwdJ) {
        __syncthreads();
        // gamma chunk: element e -> (point e / 16, column e % 16); consecutive threads walk a point's columns (coalesced
        // for the contiguous (B,N,J) layout the clustering kernel writes)
        for (int e = tid; e < kBwdPts * kBwdJ; e += kBwdThreads) {
            const int p = e / kBwdJ, jj = e - p * kBwdJ;
            float v = 0.f;
            if (n0 + p < N && j0 + jj < J) v = gb[(int64_t)(n0 + p) * g_sn + (int64_t)(j0 + jj) * g_sj];
            s_g[jj][p] = v;
        }
        for (int e = tid; e < kBwdJ * kBwdRows; e += kBwdThreads) {
            const int jj = e / kBwdRows, r = e - jj * kBwdRows;
            float v = 0.f;
            if (j0 + jj < J && d0 + r < D) {
                const float npi = __fadd_rn(__fmul_rn(pb[j0 + jj], (float)N), 1e-5f);      // lib/utils.py:138
                v = __fdiv_rn(mb[(int64_t)(j0 + jj) * D + d0 + r], npi);
            }
            s_s[jj][r] = v;
        }
        __syncthreads();
#pragma unroll
        for (int jj = 0; jj < kBwdJ; ++jj) {
            const float4 g = *reinterpret_cast<const float4*>(&s_g[jj][4 * tx]);
            const float4 sa = *reinterpret_cast<const float4*>(&s_s[jj][8 * ty]);
            const float4 sb = *reinterpret_cast<const float4*>(&s_s[jj][8 * ty + 4]);
            const float s[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
            const float2 g01 = make_float2(g.x, g.y), g23 = make_float2(g.z, g.w);
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const float2 ss = make_float2(s[r], s[r]);
                ffma2_pair(acc[r][0], ss, g01);
                ffma2_pair(acc[r][1], ss, g23);
            }
        }
    }
    float* ob = dfeats + (int64_t)b * o_sb;
    const int n = n0 + 4 * tx;
    const bool vec = (o_sn == 1) && ((o_sd & 3) == 0) && ((o_sb & 3) == 0) && ((reinterpret_cast<uintptr_t>(dfeats) & 15) == 0) && n + 3 < N;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int d = d0 + 8 * ty + r;
        if (d >= D) continue;
        float* o = ob + (int64_t)d * o_sd + (int64_t)n * o_sn;
        if (vec) {
            *reinterpret_cast<float4*>(o) = make_float4(acc[r][0].x, acc[r][0].y, acc[r][1].x, acc[r][1].y);
        } else {
            const float v[4] = {acc[r][0].x, acc[r][0].y, acc[r][1].x, acc[r][1].y};
#pragma unroll
            for (int p = 0; p < 4; ++p)
                if (n + p < N) o[(int64_t)p * o_sn] = v[p];
        }
    }
}

}  // namespace ogmm

using namespace ogmm;

extern "C" __attribute__((visibility("default"))) int ogmm_gmm_moments_feat_backward(
    const float* gamma, int64_t g_sb, int64_t g_sn, int64_t g_sj, const float* grad_mu, const float* pi,
    int64_t B, int64_t N, int64_t J, int64_t D, float* grad_feats, int64_t o_sb, int64_t o_sn, int64_t o_sd,
    ogmm_stream_t stream) {
    OGMM_REQUIRE(B >= 0 && N >= 1 && J >= 1 && D >= 1 && N < (1ll << 31) && D < (1ll << 31) && B < 65536, OGMM_EINVAL,
                 "ogmm_gmm_moments_feat_backward: bad sizes B=%lld N=%lld J=%lld D=%lld", (long long)B, (long long)N,
                 (long long)J, (long long)D);
    if (B == 0) return OGMM_OK;
    OGMM_REQUIRE(gamma && grad_mu && pi && grad_feats, OGMM_EINVAL, "ogmm_gmm_moments_feat_backward: null pointer");
    OGMM_REQUIRE((D + kBwdRows - 1) / kBwdRows < 65536, OGMM_EUNSUPPORTED, "ogmm_gmm_moments_feat_backward: D=%lld too large", (long long)D);
    dim3 grid((unsigned)((N + kBwdPts - 1) / kBwdPts), (unsigned)((D + kBwdRows - 1) / kBwdRows), (unsigned)B);
    gmm_moments_feat_bwd_kernel<<<grid, kBwdThreads, 0, as_stream(stream)>>>(gamma, g_sb, g_sn, g_sj, grad_mu, pi, (int)N, (int)J,
                                                                              (int)D, grad_feats, o_sb, o_sn, o_sd);
    OGMM_LAUNCH_CHECK("gmm_moments_feat_bwd_kernel");
    return OGMM_OK;
}
