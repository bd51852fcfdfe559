// Backward of the DeepGMR path's two kernels (sm_100a): what autograd gives the reference when baseline/deepgmr.py:64-79
// is trained -- gradients flow from the transform through gmm_register (:17-38) into (pi, mu, sigma) and from there
// through gmm_params(gamma, pts, return_sigma=True) (lib/utils.py:130-149) into gamma = softmax(logits).
//
//   gmm_moments_small_bwd_kernel   dL/dgamma of the xyz moments.  With npi_j = pi_j N + 1e-5, d_nj = x_n - mu_j:
//       pi_j    = sum_n gamma_nj / N                         ->  1 / N
//       mu_j    = sum_n gamma_nj x_n / npi_j                 ->  d_nj / npi_j
//       sigma_j = sum_n gamma_nj |x_n - mu_j|^2 / npi_j      ->  (|d_nj|^2 - c_j . d_nj - sigma_j) / npi_j,
//       c_j = 2 (npi_j - sum_n gamma_nj) mu_j / npi_j = 2e-5 mu_j / npi_j   (the term through mu_j; it vanishes but for
//       the 1e-5 in npi).  Sigma_j = sigma_j I, so the upstream scalar is the trace of dL/dSigma_j.
//       One thread per point, the per-component constants in shared memory; the result is written with the caller's
//       strides (DeepGMR holds gamma as (B,J,N): consecutive threads write consecutive addresses).
//   gmm_register_bwd_kernel        one warp per pair.  M = sum_j pi_j A_j r_j^T + 1e-4 with A_j = mu_s_j - c_s,
//       r_j = Sigma_j^-T (mu_t_j - c_t), c_s = sum pi mu_s, c_t = sum pi mu_t; R = V diag(1,1,det) U^T is the orthogonal
//       polar factor of M^T (rotation_backward, procrustes_bwd.cu -- det(V U^T) is +-1 and its ambient autograd term is
//       cancelled by the antisymmetrisation in the SVD backward, so it contributes nothing); t = c_t - R c_s.
#include "common.cuh"
#include "svd3.cuh"

namespace ogmm {

// ---- shared with procrustes_bwd.cu by inclusion of the same few lines (kept private to each translation unit) ----------
__device__ __forceinline__ double dg_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// M (conditioned, row-major) and the upstream gradient gR (row-major) -> R = V diag(1,1,d) U^T and gM = dL/dM.
__device__ __forceinline__ void dg_rotation_backward(const double* M, const double* gR, double* R, double* gM) {
    double U[9], S[3], V[9];
    svd3(M, U, S, V);
    double P[9];
    v_d_ut(V, U, 1.0, P);
    const double d = det3(P) > 0.0 ? 1.0 : -1.0;
    v_d_ut(V, U, d, R);
    const double dd[3] = {1.0, 1.0, d};
    const double sp[3] = {S[0], S[1], d * S[2]};
    double T1[9];                                      // gR^T V
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
        for (int j = 0; j < 3; ++j) T1[3 * p + j] = gR[p] * V[j] + gR[3 + p] * V[3 + j] + gR[6 + p] * V[6 + j];
    double G[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) G[3 * i + j] = dd[i] * (U[i] * T1[j] + U[3 + i] * T1[3 + j] + U[6 + i] * T1[6 + j]);
    double H[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            double den = sp[i] + sp[j];
            if (fabs(den) < 1e-30) den = den < 0.0 ? -1e-30 : 1e-30;
            H[3 * i + j] = i == j ? 0.0 : (G[3 * i + j] - G[3 * j + i]) / den;
        }
    double T2[9];                                      // D H V^T
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) T2[3 * i + j] = dd[i] * (H[3 * i] * V[3 * j] + H[3 * i + 1] * V[3 * j + 1] + H[3 * i + 2] * V[3 * j + 2]);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) gM[3 * i + j] = U[3 * i] * T2[j] + U[3 * i + 1] * T2[3 + j] + U[3 * i + 2] * T2[6 + j];
}

__device__ __forceinline__ void dg_inverse3(const double* m, double* inv) {
    const double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
    const double r = 1.0 / (m[0] * c00 + m[1] * c01 + m[2] * c02);
    inv[0] = c00 * r; inv[1] = (m[2] * m[7] - m[1] * m[8]) * r; inv[2] = (m[1] * m[5] - m[2] * m[4]) * r;
    inv[3] = c01 * r; inv[4] = (m[0] * m[8] - m[2] * m[6]) * r; inv[5] = (m[2] * m[3] - m[0] * m[5]) * r;
    inv[6] = c02 * r; inv[7] = (m[1] * m[6] - m[0] * m[7]) * r; inv[8] = (m[0] * m[4] - m[1] * m[3]) * r;
}

// ---------------------------------------------------------------------------------------------
constexpr int kMbThreads = 256;
constexpr int kMbConst = 12;        // floats per component: mu (3), 1 / npi, sigma, gpi / N, gmu (3), gsigma, pad

__global__ void __launch_bounds__(kMbThreads)
gmm_moments_small_bwd_kernel(const float* __restrict__ pts, int64_t p_sb, int64_t p_sn, int64_t p_sc,
                             const float* __restrict__ pi, const float* __restrict__ mu, const float* __restrict__ sigma,
                             const float* __restrict__ grad_pi, const float* __restrict__ grad_mu,
                             const float* __restrict__ grad_sigma, int N, int J,
                             float* __restrict__ grad_gamma, int64_t o_sb, int64_t o_sn, int64_t o_sj) {
    extern __shared__ __align__(16) float mb_c[];
    const int b = blockIdx.y, tid = threadIdx.x;
    for (int j = tid; j < J; j += kMbThreads) {
        const float npi = __fadd_rn(__fmul_rn(pi[(int64_t)b * J + j], (float)N), 1e-5f);      // lib/utils.py:138
        const float inv = 1.0f / npi;
        float* c = mb_c + (size_t)j * kMbConst;
        const float* m = mu + ((int64_t)b * J + j) * 3;
        c[0] = m[0]; c[1] = m[1]; c[2] = m[2];
        c[3] = inv;
        c[4] = sigma ? sigma[((int64_t)b * J + j) * 9] : 0.f;                                   // Sigma_j = sigma_j I
        c[5] = grad_pi ? grad_pi[(int64_t)b * J + j] / (float)N : 0.f;
        const float* gm = grad_mu ? grad_mu + ((int64_t)b * J + j) * 3 : nullptr;
        c[6] = gm ? gm[0] * inv : 0.f; c[7] = gm ? gm[1] * inv : 0.f; c[8] = gm ? gm[2] * inv : 0.f;
        const float* gs = grad_sigma ? grad_sigma + ((int64_t)b * J + j) * 9 : nullptr;
        c[9] = gs ? (gs[0] + gs[4] + gs[8]) * inv : 0.f;                                       // trace of dL/dSigma_j, / npi
        c[10] = 2e-5f * inv;                                                                   // c_j = c[10] * mu_j
        c[11] = 0.f;
    }
    __syncthreads();
    const int n = blockIdx.x * kMbThreads + tid;
    if (n >= N) return;
    const float* p = pts + (int64_t)b * p_sb + (int64_t)n * p_sn;
    const float x = p[0], y = p[p_sc], z = p[2 * p_sc];
    float* o = grad_gamma + (int64_t)b * o_sb + (int64_t)n * o_sn;
    for (int j = 0; j < J; ++j) {
        const float4 c0 = *reinterpret_cast<const float4*>(mb_c + (size_t)j * kMbConst);
        const float4 c1 = *reinterpret_cast<const float4*>(mb_c + (size_t)j * kMbConst + 4);
        const float4 c2 = *reinterpret_cast<const float4*>(mb_c + (size_t)j * kMbConst + 8);
        const float dx = x - c0.x, dy = y - c0.y, dz = z - c0.z;
        const float sq = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        const float cd = c2.z * fmaf(c0.z, dz, fmaf(c0.y, dy, c0.x * dx));                     // c_j . d
        float v = c1.y;                                                                        // gpi / N
        v = fmaf(c1.z, dx, fmaf(c1.w, dy, fmaf(c2.x, dz, v)));                                 // gmu . d / npi
        v = fmaf(c2.y, sq - cd - c1.x, v);                                                     // gsigma (|d|^2 - c.d - sigma) / npi
        o[(int64_t)j * o_sj] = v;
    }
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
gmm_register_bwd_kernel(const float* __restrict__ pi_s, const float* __restrict__ mu_s, const float* __restrict__ mu_t,
                        const float* __restrict__ sigma_t, int B, int J, const float* __restrict__ grad_tf,
                        float* __restrict__ g_pi, float* __restrict__ g_mu_s, float* __restrict__ g_mu_t,
                        float* __restrict__ g_sigma) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= B) return;
    const float* pi = pi_s + (int64_t)warp * J;
    const float* ms = mu_s + (int64_t)warp * J * 3;
    const float* mt = mu_t + (int64_t)warp * J * 3;
    const float* sg = sigma_t + (int64_t)warp * J * 9;
    // forward: c_s, c_t, M
    double cs[3] = {0.0, 0.0, 0.0}, ct[3] = {0.0, 0.0, 0.0};
    for (int j = lane; j < J; j += 32) {
        const double p = (double)pi[j];
#pragma unroll
        for (int a = 0; a < 3; ++a) { cs[a] += p * (double)ms[3 * j + a]; ct[a] += p * (double)mt[3 * j + a]; }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) { cs[a] = dg_warp_sum(cs[a]); ct[a] = dg_warp_sum(ct[a]); }
    double M[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) M[i] = 0.0;
    for (int j = lane; j < J; j += 32) {
        const double p = (double)pi[j];
        double A[3], Bv[3], S[9], W[9], r[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) { A[c] = (double)ms[3 * j + c] - cs[c]; Bv[c] = (double)mt[3 * j + c] - ct[c]; }
#pragma unroll
        for (int c = 0; c < 9; ++c) S[c] = (double)sg[9 * j + c];
        dg_inverse3(S, W);
#pragma unroll
        for (int c = 0; c < 3; ++c) r[c] = Bv[0] * W[c] + Bv[1] * W[3 + c] + Bv[2] * W[6 + c];      // r = W^T B
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int c = 0; c < 3; ++c) M[3 * a + c] += p * A[a] * r[c];
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        double m = dg_warp_sum(M[i]);
        if (m != m) m = 0.0;                                         // nan_to_num(Ms, nan=0) + 1e-4 on all nine entries
        else if (m > 3.402823466e+38) m = 3.402823466e+38;
        else if (m < -3.402823466e+38) m = -3.402823466e+38;
        M[i] = m + 1e-4;
    }
    // upstream: T = [[R t], [0 0 0 1]]
    const float* gT = grad_tf + (int64_t)warp * 16;
    double gR[9], gt[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) gt[a] = (double)gT[4 * a + 3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int c = 0; c < 3; ++c) gR[3 * a + c] = (double)gT[4 * a + c] - gt[a] * cs[c];          // t = c_t - R c_s
    double R[9], gM[9];
    dg_rotation_backward(M, gR, R, gM);                              // every lane, redundantly
    double gcs[3], gct[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) { gcs[a] = -(R[a] * gt[0] + R[3 + a] * gt[1] + R[6 + a] * gt[2]); gct[a] = gt[a]; }
    // pass 2: the sums of dL/dA_j and dL/dB_j feed the centroids
    double sA[3] = {0.0, 0.0, 0.0}, sB[3] = {0.0, 0.0, 0.0};
    for (int j = lane; j < J; j += 32) {
        const double p = (double)pi[j];
        double A[3], Bv[3], S[9], W[9], r[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) { A[c] = (double)ms[3 * j + c] - cs[c]; Bv[c] = (double)mt[3 * j + c] - ct[c]; }
#pragma unroll
        for (int c = 0; c < 9; ++c) S[c] = (double)sg[9 * j + c];
        dg_inverse3(S, W);
#pragma unroll
        for (int c = 0; c < 3; ++c) r[c] = Bv[0] * W[c] + Bv[1] * W[3 + c] + Bv[2] * W[6 + c];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            sA[a] += p * (gM[3 * a] * r[0] + gM[3 * a + 1] * r[1] + gM[3 * a + 2] * r[2]);          // dL/dA_j = pi gM r
            double gr[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) gr[c] = p * (gM[c] * A[0] + gM[3 + c] * A[1] + gM[6 + c] * A[2]);   // pi gM^T A
            sB[a] += W[3 * a] * gr[0] + W[3 * a + 1] * gr[1] + W[3 * a + 2] * gr[2];                // dL/dB_j = W gr
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) { gcs[a] -= dg_warp_sum(sA[a]); gct[a] -= dg_warp_sum(sB[a]); }
    // pass 3: per component
    for (int j = lane; j < J; j += 32) {
        const double p = (double)pi[j];
        double A[3], Bv[3], S[9], W[9], r[3], gA[3], gr[3], gB[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) { A[c] = (double)ms[3 * j + c] - cs[c]; Bv[c] = (double)mt[3 * j + c] - ct[c]; }
#pragma unroll
        for (int c = 0; c < 9; ++c) S[c] = (double)sg[9 * j + c];
        dg_inverse3(S, W);
#pragma unroll
        for (int c = 0; c < 3; ++c) r[c] = Bv[0] * W[c] + Bv[1] * W[3 + c] + Bv[2] * W[6 + c];
        double gp = 0.0;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double Mr = gM[3 * a] * r[0] + gM[3 * a + 1] * r[1] + gM[3 * a + 2] * r[2];
            gp += A[a] * Mr;                                                                        // A^T gM r
            gA[a] = p * Mr;
            gr[a] = p * (gM[a] * A[0] + gM[3 + a] * A[1] + gM[6 + a] * A[2]);
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) gB[a] = W[3 * a] * gr[0] + W[3 * a + 1] * gr[1] + W[3 * a + 2] * gr[2];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            gp += (double)ms[3 * j + a] * gcs[a] + (double)mt[3 * j + a] * gct[a];
            g_mu_s[((int64_t)warp * J + j) * 3 + a] = (float)(gA[a] + p * gcs[a]);
            g_mu_t[((int64_t)warp * J + j) * 3 + a] = (float)(gB[a] + p * gct[a]);
        }
        g_pi[(int64_t)warp * J + j] = (float)gp;
        // W = Sigma^-1, dL/dW[k][c] = B[k] gr[c]  ->  dL/dSigma = -W^T (dL/dW) W^T
        double T[9];                                         // (dL/dW) W^T : T[k][d] = B[k] sum_c gr[c] W[d][c]
        double wr[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) wr[d] = gr[0] * W[3 * d] + gr[1] * W[3 * d + 1] + gr[2] * W[3 * d + 2];
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int d = 0; d < 3; ++d) T[3 * k + d] = Bv[k] * wr[d];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int d = 0; d < 3; ++d)
                g_sigma[((int64_t)warp * J + j) * 9 + 3 * a + d] =
                    (float)(-(W[a] * T[d] + W[3 + a] * T[3 + d] + W[6 + a] * T[6 + d]));           // -(W^T T)[a][d]
    }
}

}  // namespace ogmm

using namespace ogmm;

extern "C" __attribute__((visibility("default"))) int ogmm_gmm_moments_backward(
    const float* pts, int64_t p_sb, int64_t p_sn, int64_t p_sc, const float* pi, const float* mu, const float* sigma,
    const float* grad_pi, const float* grad_mu, const float* grad_sigma, int64_t B, int64_t N, int64_t J,
    float* grad_gamma, int64_t o_sb, int64_t o_sn, int64_t o_sj, ogmm_stream_t stream) {
    OGMM_REQUIRE(B >= 0 && N >= 1 && J >= 1 && B < 65536 && N < (1ll << 31), OGMM_EINVAL,
                 "ogmm_gmm_moments_backward: bad sizes B=%lld N=%lld J=%lld", (long long)B, (long long)N, (long long)J);
    OGMM_REQUIRE(J <= 1024, OGMM_EUNSUPPORTED, "ogmm_gmm_moments_backward: J=%lld > 1024", (long long)J);
    if (B == 0) return OGMM_OK;
    OGMM_REQUIRE(pts && pi && mu && grad_gamma, OGMM_EINVAL, "ogmm_gmm_moments_backward: null pointer");
    OGMM_REQUIRE(grad_sigma == nullptr || sigma != nullptr, OGMM_EINVAL, "ogmm_gmm_moments_backward: grad_sigma without sigma");
    const size_t smem = sizeof(float) * kMbConst * (size_t)J;
    if (smem > 48 * 1024) {
        int st = cuda_status(cudaFuncSetAttribute(gmm_moments_small_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                             "cudaFuncSetAttribute(gmm_moments_small_bwd_kernel)");
        if (st != OGMM_OK) return st;
    }
    dim3 grid((unsigned)((N + kMbThreads - 1) / kMbThreads), (unsigned)B);
    gmm_moments_small_bwd_kernel<<<grid, kMbThreads, smem, as_stream(stream)>>>(pts, p_sb, p_sn, p_sc, pi, mu, sigma, grad_pi, grad_mu,
                                                                                grad_sigma, (int)N, (int)J, grad_gamma, o_sb, o_sn, o_sj);
    OGMM_LAUNCH_CHECK("gmm_moments_small_bwd_kernel");
    return OGMM_OK;
}

extern "C" __attribute__((visibility("default"))) int ogmm_gmm_register_backward(
    const float* pi_s, const float* mu_s, const float* mu_t, const float* sigma_t, int64_t B, int64_t J,
    const float* grad_transform, float* grad_pi_s, float* grad_mu_s, float* grad_mu_t, float* grad_sigma_t,
    ogmm_stream_t stream) {
    OGMM_REQUIRE(B >= 0 && J >= 1 && B < (1ll << 31) && J < (1ll << 31), OGMM_EINVAL, "ogmm_gmm_register_backward: bad sizes");
    if (B == 0) return OGMM_OK;
    OGMM_REQUIRE(pi_s && mu_s && mu_t && sigma_t && grad_transform && grad_pi_s && grad_mu_s && grad_mu_t && grad_sigma_t,
                 OGMM_EINVAL, "ogmm_gmm_register_backward: null pointer");
    const int threads = 128;
    const int blocks = (int)((B * 32 + threads - 1) / threads);
    gmm_register_bwd_kernel<<<blocks, threads, 0, as_stream(stream)>>>(pi_s, mu_s, mu_t, sigma_t, (int)B, (int)J, grad_transform,
                                                                       grad_pi_s, grad_mu_s, grad_mu_t, grad_sigma_t);
    OGMM_LAUNCH_CHECK("gmm_register_bwd_kernel");
    return OGMM_OK;
}
