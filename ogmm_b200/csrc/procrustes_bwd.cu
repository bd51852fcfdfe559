// K4 backward: gradients of the weighted-Procrustes head (sm_100a).
//
//   rigid_transform_bwd_kernel   lib/se3.py:256-289 (compute_rigid_transformation)   one warp per batch element
//   soft_procrustes_bwd_kernel   models/dgcnn.py:96-115 (GMMSVD.forward, is_sk=False) one CTA per pair
//
// The reference differentiates this head with autograd through torch.svd (train.py:69-75 back-propagates the
// registration loss on (R, t) into the cluster descriptors).  Here the chain is written out in closed form and
// evaluated in one launch; the forward quantities are recomputed from the inputs (a few hundred KB per pair), so the
// forward kernel saves nothing.
//
// Rotation.  M = cov + 1e-5 I = U S V^T, R = V D U^T with D = diag(1, 1, d), d = sign(det(V U^T)).  Q = R^T = U D V^T
// is the orthogonal polar factor of M = Q P, P = V (D S) V^T.  Differentiating M = Q P with Q^T dQ antisymmetric:
//
//     dQ = U D Omega V^T,   Omega_ij = (B_ij - B_ji) / (s'_i + s'_j),   B = D U^T dM V,   s' = (s1, s2, d s3)
//
// so for an upstream gradient gR (gQ = gR^T):  G = D U^T gQ V,  H_ij = (G_ij - G_ji) / (s'_i + s'_j),  gM = U D H V^T.
// This is the derivative of the same function torch.svd's backward differentiates, without its 1 / (s_i^2 - s_j^2)
// terms (which cancel analytically but lose digits when two singular values are close).
//
// Covariance.  cov = sum_n w_n (s_n - a)(c_n - b)^T with a, b the weighted centroids: sum_n w_n (s_n - a) = 0 makes the
// terms through a and b vanish, leaving  g w_n = (s_n - a)^T gM (c_n - b),  g s_n = w_n gM (c_n - b),
// g c_n = w_n gM^T (s_n - a).  t = -R a + b adds gR -= gt a^T, ga = -R^T gt, gb = gt, spread over the points by
// a = sum w s / W, b = sum w c / W.
//
// Soft correspondences (GMMSVD).  c_n = sum_m P_nm tgt_m, w_n = sum_m P_nm, P = softmax_m(sim / T),
// sim_nm = <x_n / |x_n|, y_m / |y_m|>:
//     gP_nm = <g c_n, tgt_m> + g w_n,   gsim = P (gP - rowsum(P gP)) / T,
//     g x_n = (sum_m gsim_nm yh_m - xh_n sum_m gsim_nm sim_nm) / |x_n|      (and symmetrically for y),
// where the projection term uses  <xh_n, sum_m gsim_nm yh_m> = sum_m gsim_nm sim_nm,  so no second reduction over the
// descriptor dimension is needed.  All 3x3 algebra runs in fp64 registers, the J x J stage in fp32.
#include "common.cuh"
#include "svd3.cuh"

namespace ogmm {

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// The conditioned covariance the forward decomposes (procrustes.cu finish_procrustes: nan_to_num, + 1e-5 on the diagonal).
__device__ __forceinline__ void condition_cov(double* cov) {
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        double c = cov[i];
        if (c != c) c = 0.0;
        else if (c > 3.402823466e+38) c = 3.402823466e+38;
        else if (c < -3.402823466e+38) c = -3.402823466e+38;
        if (i % 4 == 0) c = c + 1e-5;
        cov[i] = c;
    }
}

// cov (conditioned, row-major) and the upstream gradient gR (row-major) -> R and gM = dL/dcov.
__device__ __forceinline__ void rotation_backward(const double* cov, const double* gR, double* R, double* gM) {
    double U[9], S[3], V[9];
    svd3(cov, U, S, V);
    v_d_ut(V, U, 1.0, R);
    double d = 1.0;
    if (!(det3(R) > 0.0)) { d = -1.0; v_d_ut(V, U, -1.0, R); }
    const double dd[3] = {1.0, 1.0, d};
    const double sp[3] = {S[0], S[1], d * S[2]};
    // G = D U^T gQ V with gQ = gR^T:  G_ij = dd_i sum_{p,q} U_pi gR_qp V_qj
    double T1[9];                                      // T1 = gR^T V  (T1_pj = sum_q gR_qp V_qj)
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
    // gM = U D H V^T
    double T2[9];                                      // T2 = D H V^T
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) T2[3 * i + j] = dd[i] * (H[3 * i] * V[3 * j] + H[3 * i + 1] * V[3 * j + 1] + H[3 * i + 2] * V[3 * j + 2]);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) gM[3 * i + j] = U[3 * i] * T2[j] + U[3 * i + 1] * T2[3 + j] + U[3 * i + 2] * T2[6 + j];
}

// Everything between (centroids, covariance, upstream gR / gt) and the per-point gradient coefficients.
struct HeadGrad {
    double gM[9], a[3], b[3], ga[3], gb[3], invW;
};

__device__ __forceinline__ void head_grad(double* cov, const double* a, const double* b, double W, const float* gR_in,
                                          const float* gt_in, HeadGrad& h) {
    condition_cov(cov);
    double gR[9], gt[3], R[9];
#pragma unroll
    for (int i = 0; i < 3; ++i) gt[i] = gt_in ? (double)gt_in[i] : 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) gR[3 * i + j] = (gR_in ? (double)gR_in[3 * i + j] : 0.0) - gt[i] * a[j];   // t = -R a + b
    rotation_backward(cov, gR, R, h.gM);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        h.a[i] = a[i]; h.b[i] = b[i];
        h.ga[i] = -(R[i] * gt[0] + R[3 + i] * gt[1] + R[6 + i] * gt[2]);
        h.gb[i] = gt[i];
    }
    h.invW = 1.0 / W;
}

// Per-point gradients from the coefficients: s, c the point and its correspondence, w its weight.
__device__ __forceinline__ void point_grad(const HeadGrad& h, const double* s, const double* c, double w, double* gs,
                                           double* gc, double& gw) {
    double ds[3], dc[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) { ds[i] = s[i] - h.a[i]; dc[i] = c[i] - h.b[i]; }
    double Mdc[3], Mtds[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        Mdc[i] = h.gM[3 * i] * dc[0] + h.gM[3 * i + 1] * dc[1] + h.gM[3 * i + 2] * dc[2];
        Mtds[i] = h.gM[i] * ds[0] + h.gM[3 + i] * ds[1] + h.gM[6 + i] * ds[2];
    }
    gw = ds[0] * Mdc[0] + ds[1] * Mdc[1] + ds[2] * Mdc[2]
       + (ds[0] * h.ga[0] + ds[1] * h.ga[1] + ds[2] * h.ga[2] + dc[0] * h.gb[0] + dc[1] * h.gb[1] + dc[2] * h.gb[2]) * h.invW;
    const double wW = w * h.invW;
#pragma unroll
    for (int i = 0; i < 3; ++i) { gs[i] = w * Mdc[i] + wW * h.ga[i]; gc[i] = w * Mtds[i] + wW * h.gb[i]; }
}

__global__ void __launch_bounds__(128)
rigid_transform_bwd_kernel(const float* __restrict__ src, int64_t s_sb, int64_t s_sc, int64_t s_sn,
                           const float* __restrict__ corr, int64_t c_sb, int64_t c_sc, int64_t c_sn,
                           const float* __restrict__ weight, int64_t w_sb, int64_t w_sn, int B, int n,
                           const float* __restrict__ grad_rot, const float* __restrict__ grad_trans,
                           float* __restrict__ grad_src, float* __restrict__ grad_corr, float* __restrict__ grad_weight) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= B) return;
    const float* s = src + (int64_t)warp * s_sb;
    const float* c = corr + (int64_t)warp * c_sb;
    const float* w = weight + (int64_t)warp * w_sb;

    double sw = 0.0, ss[3] = {0.0, 0.0, 0.0}, sc[3] = {0.0, 0.0, 0.0};
    for (int i = lane; i < n; i += 32) {
        const double wi = (double)w[i * w_sn];
        sw += wi;
#pragma unroll
        for (int a = 0; a < 3; ++a) { ss[a] += (double)s[a * s_sc + i * s_sn] * wi; sc[a] += (double)c[a * c_sc + i * c_sn] * wi; }
    }
    sw = warp_sum_f64(sw);
    double ca[3], cb[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) { ca[a] = warp_sum_f64(ss[a]) / sw; cb[a] = warp_sum_f64(sc[a]) / sw; }
    double cov[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) cov[i] = 0.0;
    for (int i = lane; i < n; i += 32) {
        const double wi = (double)w[i * w_sn];
        double sa[3], cc[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) { sa[a] = ((double)s[a * s_sc + i * s_sn] - ca[a]) * wi; cc[a] = (double)c[a * c_sc + i * c_sn] - cb[a]; }
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int q = 0; q < 3; ++q) cov[3 * a + q] += sa[a] * cc[q];
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) cov[i] = warp_sum_f64(cov[i]);
    // every lane holds the same covariance: the 3x3 algebra runs redundantly instead of being broadcast
    HeadGrad h;
    head_grad(cov, ca, cb, sw, grad_rot ? grad_rot + (int64_t)warp * 9 : nullptr,
              grad_trans ? grad_trans + (int64_t)warp * 3 : nullptr, h);
    float* gs_out = grad_src + (int64_t)warp * 3 * n;
    float* gc_out = grad_corr + (int64_t)warp * 3 * n;
    float* gw_out = grad_weight + (int64_t)warp * n;
    for (int i = lane; i < n; i += 32) {
        double sp[3], cp[3], gs[3], gc[3], gw;
#pragma unroll
        for (int a = 0; a < 3; ++a) { sp[a] = (double)s[a * s_sc + i * s_sn]; cp[a] = (double)c[a * c_sc + i * c_sn]; }
        point_grad(h, sp, cp, (double)w[i * w_sn], gs, gc, gw);
#pragma unroll
        for (int a = 0; a < 3; ++a) { gs_out[a * n + i] = (float)gs[a]; gc_out[a * n + i] = (float)gc[a]; }
        gw_out[i] = (float)gw;
    }
}

// ---------------------------------------------------------------------------------------------
// GMMSVD (is_sk = False) backward.  Dynamic shared memory (floats):
//   sim [Js][Jt], P [Js][Jt] (scores, then gsim in place), tile [(Js+Jt)][kBT+4] normalised descriptor chunk,
//   den [Js+Jt] (max(|row|, 1e-12)), proj [Js+Jt] (sum gsim*sim per row / column; 0 for rows under the 1e-12 clamp),
//   corr [3][Js], wgt [Js], mus [Js][3], mut [Jt][3], gcn [3][Js], gwn [Js]
constexpr int kBT = 64, kBTP = kBT + 4, kBThreads = 256;

__host__ __device__ inline size_t soft_procrustes_bwd_smem(int Js, int Jt) {
    return sizeof(float) * (2 * (((size_t)Js * Jt + 3) & ~(size_t)3) + (size_t)(Js + Jt) * kBTP + 2 * (size_t)(Js + Jt) + 3 * (size_t)Js + Js
                            + 3 * (size_t)(Js + Jt) + 3 * (size_t)Js + Js + 32);
}

__global__ void __launch_bounds__(kBThreads)
soft_procrustes_bwd_kernel(const float* __restrict__ src_mu, const float* __restrict__ tgt_mu,
                           const float* __restrict__ src_desc, const float* __restrict__ tgt_desc,
                           int Js, int Jt, int D, float temperature,
                           const float* __restrict__ grad_rot, const float* __restrict__ grad_trans,
                           const float* __restrict__ grad_corr,
                           float* __restrict__ grad_src_mu, float* __restrict__ grad_tgt_mu,
                           float* __restrict__ grad_src_desc, float* __restrict__ grad_tgt_desc) {
    extern __shared__ __align__(16) float smem[];
    float* sim = smem;
    const size_t jj4 = ((size_t)Js * Jt + 3) & ~(size_t)3;      // the tile is read as float4: keep it 16-byte aligned
    float* P = sim + jj4;
    float* tile = P + jj4;
    float* den = tile + (size_t)(Js + Jt) * kBTP;
    float* proj = den + (Js + Jt);
    float* corr = proj + (Js + Jt);
    float* wgt = corr + 3 * Js;
    float* mus = wgt + Js;
    float* mut = mus + 3 * Js;
    float* gcn = mut + 3 * Jt;
    float* gwn = gcn + 3 * Js;
    __shared__ HeadGrad sh;

    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = kBThreads / 32;
    const int R = Js + Jt, npairs = Js * Jt;
    const float* xd = src_desc + (int64_t)b * Js * D;
    const float* yd = tgt_desc + (int64_t)b * Jt * D;

    // ---- forward recomputation: row norms, similarity, scores, soft correspondences (same arithmetic as the forward) ----
    for (int r = warp; r < R; r += NW) {
        const float* row = r < Js ? xd + (int64_t)r * D : yd + (int64_t)(r - Js) * D;
        float acc = 0.f;
        for (int d = lane; d < D; d += 32) { const float v = row[d]; acc += v * v; }
        acc = warp_sum(acc);
        if (lane == 0) den[r] = fmaxf(sqrtf(acc), 1e-12f);
    }
    for (int i = tid; i < 3 * Js; i += kBThreads) mus[i] = src_mu[(int64_t)b * Js * 3 + i];
    for (int i = tid; i < 3 * Jt; i += kBThreads) mut[i] = tgt_mu[(int64_t)b * Jt * 3 + i];
    for (int p = tid; p < npairs; p += kBThreads) sim[p] = 0.f;
    __syncthreads();

    auto load_tile = [&](int d0) {
        const int dw = min(kBT, D - d0);
        for (int e = tid; e < R * kBT; e += kBThreads) {
            const int r = e / kBT, d = e - r * kBT;
            float v = 0.f;
            if (d < dw) v = (r < Js ? xd[(int64_t)r * D + d0 + d] : yd[(int64_t)(r - Js) * D + d0 + d]) / den[r];
            tile[r * kBTP + d] = v;
        }
    };
    for (int d0 = 0; d0 < D; d0 += kBT) {
        __syncthreads();
        load_tile(d0);
        __syncthreads();
        for (int p = tid; p < npairs; p += kBThreads) {
            const int i = p / Jt, j = p - i * Jt;
            const float4* xa = reinterpret_cast<const float4*>(tile + i * kBTP);
            const float4* ya = reinterpret_cast<const float4*>(tile + (Js + j) * kBTP);
            float a = sim[p];
#pragma unroll
            for (int d = 0; d < kBT / 4; ++d) {
                const float4 u = xa[d], v = ya[d];
                a = fmaf(u.x, v.x, a); a = fmaf(u.y, v.y, a); a = fmaf(u.z, v.z, a); a = fmaf(u.w, v.w, a);
            }
            sim[p] = a;
        }
    }
    __syncthreads();
    for (int i = warp; i < Js; i += NW) {
        const float* srow = sim + (size_t)i * Jt;
        float* prow = P + (size_t)i * Jt;
        float m = -INFINITY;
        for (int j = lane; j < Jt; j += 32) { const float z = __fdiv_rn(srow[j], temperature); prow[j] = z; m = fmaxf(m, z); }
        m = warp_max(m);
        float sum = 0.f;
        for (int j = lane; j < Jt; j += 32) { const float e = expf(prow[j] - m); prow[j] = e; sum += e; }
        sum = warp_sum(sum);
        float c0 = 0.f, c1 = 0.f, c2 = 0.f, w = 0.f;
        for (int j = lane; j < Jt; j += 32) {
            const float sc = prow[j] / sum;
            prow[j] = sc;
            w += sc;
            c0 = fmaf(mut[3 * j], sc, c0); c1 = fmaf(mut[3 * j + 1], sc, c1); c2 = fmaf(mut[3 * j + 2], sc, c2);
        }
        c0 = warp_sum(c0); c1 = warp_sum(c1); c2 = warp_sum(c2); w = warp_sum(w);
        if (lane == 0) { corr[i] = c0; corr[Js + i] = c1; corr[2 * Js + i] = c2; wgt[i] = w; }
    }
    __syncthreads();

    // ---- 3x3 stage in fp64 (warp 0): centroids, covariance, SVD, dL/dcov ----
    if (warp == 0) {
        double sw = 0.0, ss[3] = {0.0, 0.0, 0.0}, sc[3] = {0.0, 0.0, 0.0};
        for (int i = lane; i < Js; i += 32) {
            const double wi = (double)wgt[i];
            sw += wi;
#pragma unroll
            for (int a = 0; a < 3; ++a) { ss[a] += (double)mus[3 * i + a] * wi; sc[a] += (double)corr[a * Js + i] * wi; }
        }
        sw = warp_sum_f64(sw);
        double ca[3], cb[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) { ca[a] = warp_sum_f64(ss[a]) / sw; cb[a] = warp_sum_f64(sc[a]) / sw; }
        double cov[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) cov[i] = 0.0;
        for (int i = lane; i < Js; i += 32) {
            const double wi = (double)wgt[i];
            double sa[3], cc[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) { sa[a] = ((double)mus[3 * i + a] - ca[a]) * wi; cc[a] = (double)corr[a * Js + i] - cb[a]; }
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int q = 0; q < 3; ++q) cov[3 * a + q] += sa[a] * cc[q];
        }
#pragma unroll
        for (int i = 0; i < 9; ++i) cov[i] = warp_sum_f64(cov[i]);
        if (lane == 0)
            head_grad(cov, ca, cb, sw, grad_rot ? grad_rot + (int64_t)b * 9 : nullptr,
                      grad_trans ? grad_trans + (int64_t)b * 3 : nullptr, sh);
    }
    __syncthreads();

    // ---- per source component: dL/dw_n, dL/dcorr_n (plus the upstream gradient on src_corr), dL/dsrc_mu ----
    for (int i = tid; i < Js; i += kBThreads) {
        double sp[3], cp[3], gs[3], gc[3], gw;
#pragma unroll
        for (int a = 0; a < 3; ++a) { sp[a] = (double)mus[3 * i + a]; cp[a] = (double)corr[a * Js + i]; }
        point_grad(sh, sp, cp, (double)wgt[i], gs, gc, gw);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            if (grad_corr) gc[a] += (double)grad_corr[(int64_t)b * 3 * Js + a * Js + i];
            gcn[a * Js + i] = (float)gc[a];
            grad_src_mu[(int64_t)b * Js * 3 + 3 * i + a] = (float)gs[a];
        }
        gwn[i] = (float)gw;
    }
    __syncthreads();
    // dL/dtgt_mu[m] = sum_n P[n][m] dL/dcorr_n   (before P is overwritten)
    for (int e = tid; e < 3 * Jt; e += kBThreads) {
        const int j = e / 3, a = e - 3 * j;
        float acc = 0.f;
        for (int i = 0; i < Js; ++i) acc = fmaf(P[(size_t)i * Jt + j], gcn[a * Js + i], acc);
        grad_tgt_mu[(int64_t)b * Jt * 3 + e] = acc;
    }
    __syncthreads();
    // softmax backward per row, in place: P <- gsim; proj[n] = sum_m gsim sim.  Written as
    // gz_j = P_j sum_k P_k (gP_j - gP_k) / T  rather than  P_j (gP_j - sum_k P_k gP_k) / T:  at T = 0.05 rows are
    // often nearly one-hot, and for the dominant entry the second form cancels to fp32 noise while the differences of
    // the first stay exact (the k = j term is zero).  J^2 work per row; the idle descriptor tile is the scratch.
    for (int i = warp; i < Js; i += NW) {
        float* prow = P + (size_t)i * Jt;
        const float* srow = sim + (size_t)i * Jt;
        float* gprow = tile + (size_t)warp * Jt;
        float* gzrow = tile + (size_t)(NW + warp) * Jt;
        const float g0 = gcn[i], g1 = gcn[Js + i], g2 = gcn[2 * Js + i], gw = gwn[i];
        for (int j = lane; j < Jt; j += 32)
            gprow[j] = fmaf(g0, mut[3 * j], fmaf(g1, mut[3 * j + 1], fmaf(g2, mut[3 * j + 2], gw)));
        __syncwarp();
        float pr = 0.f;
        for (int j = lane; j < Jt; j += 32) {
            const float gpj = gprow[j];
            float acc = 0.f;
            for (int k = 0; k < Jt; ++k) acc = fmaf(prow[k], gpj - gprow[k], acc);
            const float gz = __fdiv_rn(prow[j] * acc, temperature);
            gzrow[j] = gz;
            pr = fmaf(gz, srow[j], pr);
        }
        pr = warp_sum(pr);
        __syncwarp();
        for (int j = lane; j < Jt; j += 32) prow[j] = gzrow[j];
        __syncwarp();
        // F.normalize divides by max(|x|, 1e-12): below the clamp x_hat = x / 1e-12 is linear, no projection term
        if (lane == 0) proj[i] = den[i] > 1e-12f ? pr : 0.f;
    }
    __syncthreads();
    for (int j = tid; j < Jt; j += kBThreads) {
        float pr = 0.f;
        for (int i = 0; i < Js; ++i) pr = fmaf(P[(size_t)i * Jt + j], sim[(size_t)i * Jt + j], pr);
        proj[Js + j] = den[Js + j] > 1e-12f ? pr : 0.f;
    }

    // ---- descriptor gradients, chunk by chunk: thread = (column d of the chunk, every 4th row) ----
    const int dcol = tid & (kBT - 1), rr = tid / kBT;
    constexpr int kRowStep = kBThreads / kBT;
    for (int d0 = 0; d0 < D; d0 += kBT) {
        __syncthreads();
        load_tile(d0);
        __syncthreads();
        if (d0 + dcol < D) {
            for (int r = rr; r < R; r += kRowStep) {
                float acc = 0.f;
                if (r < Js) {
                    const float* g = P + (size_t)r * Jt;
                    for (int j = 0; j < Jt; ++j) acc = fmaf(g[j], tile[(Js + j) * kBTP + dcol], acc);
                    grad_src_desc[((int64_t)b * Js + r) * D + d0 + dcol] = (acc - tile[r * kBTP + dcol] * proj[r]) / den[r];
                } else {
                    const int j = r - Js;
                    for (int i = 0; i < Js; ++i) acc = fmaf(P[(size_t)i * Jt + j], tile[i * kBTP + dcol], acc);
                    grad_tgt_desc[((int64_t)b * Jt + j) * D + d0 + dcol] = (acc - tile[r * kBTP + dcol] * proj[r]) / den[r];
                }
            }
        }
    }
}

}  // namespace ogmm

using namespace ogmm;

extern "C" __attribute__((visibility("default"))) int ogmm_rigid_transform_backward(
    const float* src, int64_t s_sb, int64_t s_sc, int64_t s_sn, const float* corr, int64_t c_sb, int64_t c_sc, int64_t c_sn,
    const float* weight, int64_t w_sb, int64_t w_sn, int64_t B, int64_t n, const float* grad_rot, const float* grad_trans,
    float* grad_src, float* grad_corr, float* grad_weight, ogmm_stream_t stream) {
    OGMM_REQUIRE(B >= 0 && n >= 1 && B < (1ll << 31) && n < (1ll << 31), OGMM_EINVAL,
                 "ogmm_rigid_transform_backward: bad sizes B=%lld n=%lld", (long long)B, (long long)n);
    if (B == 0) return OGMM_OK;
    OGMM_REQUIRE(src && corr && weight && grad_src && grad_corr && grad_weight, OGMM_EINVAL,
                 "ogmm_rigid_transform_backward: null pointer");
    const int threads = 128;
    const int blocks = (int)((B * 32 + threads - 1) / threads);
    rigid_transform_bwd_kernel<<<blocks, threads, 0, as_stream(stream)>>>(src, s_sb, s_sc, s_sn, corr, c_sb, c_sc, c_sn,
                                                                          weight, w_sb, w_sn, (int)B, (int)n, grad_rot,
                                                                          grad_trans, grad_src, grad_corr, grad_weight);
    OGMM_LAUNCH_CHECK("rigid_transform_bwd_kernel");
    return OGMM_OK;
}

extern "C" __attribute__((visibility("default"))) int ogmm_soft_procrustes_backward(
    const float* src_mu, const float* tgt_mu, const float* src_desc, const float* tgt_desc, int64_t B, int64_t Js, int64_t Jt,
    int64_t D, float temperature, const float* grad_rot, const float* grad_trans, const float* grad_corr,
    float* grad_src_mu, float* grad_tgt_mu, float* grad_src_desc, float* grad_tgt_desc, ogmm_stream_t stream) {
    OGMM_REQUIRE(B >= 0 && Js >= 1 && Jt >= 1 && D >= 1 && B < (1ll << 31), OGMM_EINVAL,
                 "ogmm_soft_procrustes_backward: bad sizes B=%lld Js=%lld Jt=%lld D=%lld", (long long)B, (long long)Js,
                 (long long)Jt, (long long)D);
    if (B == 0) return OGMM_OK;
    OGMM_REQUIRE(src_mu && tgt_mu && src_desc && tgt_desc && grad_src_mu && grad_tgt_mu && grad_src_desc && grad_tgt_desc,
                 OGMM_EINVAL, "ogmm_soft_procrustes_backward: null pointer");
    OGMM_REQUIRE(temperature > 0.f, OGMM_EINVAL, "ogmm_soft_procrustes_backward: temperature must be > 0");
    const size_t smem = soft_procrustes_bwd_smem((int)Js, (int)Jt);
    OGMM_REQUIRE(smem <= 200 * 1024, OGMM_EUNSUPPORTED,
                 "soft_procrustes_backward: Js=%lld Jt=%lld needs %zu B of shared memory (> 200 KiB)", (long long)Js,
                 (long long)Jt, smem);
    int st = cuda_status(cudaFuncSetAttribute(soft_procrustes_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)smem), "cudaFuncSetAttribute(soft_procrustes_bwd_kernel)");
    if (st != OGMM_OK) return st;
    soft_procrustes_bwd_kernel<<<(unsigned)B, kBThreads, smem, as_stream(stream)>>>(
        src_mu, tgt_mu, src_desc, tgt_desc, (int)Js, (int)Jt, (int)D, temperature, grad_rot, grad_trans, grad_corr,
        grad_src_mu, grad_tgt_mu, grad_src_desc, grad_tgt_desc);
    OGMM_LAUNCH_CHECK("soft_procrustes_bwd_kernel");
    return OGMM_OK;
}
