// K4: weighted Procrustes and the two registration heads (sm_100a).
//
//   rigid_transform_kernel   lib/se3.py:256-289          one warp per batch element
//   soft_procrustes_kernel   models/dgcnn.py:96-115       one CTA per pair (is_sk=False head)
//   cos_similarity_kernel    lib/utils.py:222-226         one CTA per (batch, 16x16 tile)
//   gmm_register_kernel      baseline/deepgmr.py:17-38    one warp per batch element
//
// These are latency-bound (tens of KB per pair); the point is to remove the reference's
// device->host->device SVD round trip and its ~20 tiny launches, not to chase a roofline.
// Everything after the fp32 inputs -- weighted centroids, the 3x3 cross-covariance, the SVD and t = c_t - R c_s --
// runs in fp64 registers (a few hundred flops per pair): the result is the correctly rounded answer for the fp32
// inputs, so its distance to the fp32 reference is the REFERENCE's own rounding error (tests print that spread:
// the fp32-vs-fp64 difference of the reference's op sequence) and rotation / translation stay inside the
// 1e-3 deg / 1e-5 x scale budget against the fp64 arbiter.
#include "common.cuh"
#include "svd3.cuh"

namespace ogmm {

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// Shared tail: covariance (row-major 9) + centroids -> R, t.  Called by one lane.  The reference adds 1e-5 * eye to an
// fp32 covariance (lib/se3.py:275): the diagonal is rounded to fp32 before the add so that term enters identically.
__device__ __forceinline__ void finish_procrustes(const double* cov_in, const double* cs, const double* ct,
                                                  float* rot_out, float* trans_out) {
    double cov[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        double c = cov_in[i];
        if (c != c) c = 0.0;                                     // nan_to_num(cov)
        else if (c > 3.402823466e+38) c = 3.402823466e+38;
        else if (c < -3.402823466e+38) c = -3.402823466e+38;
        if (i % 4 == 0) c = c + 1e-5;
        cov[i] = c;
    }
    double R[9];
    rotation_from_cov_ogmm(cov, R);
#pragma unroll
    for (int i = 0; i < 9; ++i) rot_out[i] = (float)R[i];
#pragma unroll
    for (int i = 0; i < 3; ++i)
        trans_out[i] = (float)(-(R[3 * i] * cs[0] + R[3 * i + 1] * cs[1] + R[3 * i + 2] * cs[2]) + ct[i]);
}

__global__ void __launch_bounds__(128)
rigid_transform_kernel(const float* __restrict__ src, int64_t s_sb, int64_t s_sc, int64_t s_sn,
                       const float* __restrict__ corr, int64_t c_sb, int64_t c_sc, int64_t c_sn,
                       const float* __restrict__ weight, int64_t w_sb, int64_t w_sn,
                       int B, int n, float* __restrict__ rot_out, float* __restrict__ trans_out) {
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
        for (int a = 0; a < 3; ++a) {
            ss[a] += (double)s[a * s_sc + i * s_sn] * wi;
            sc[a] += (double)c[a * c_sc + i * c_sn] * wi;
        }
    }
    sw = warp_sum_d(sw);
    double cs[3], ct[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        cs[a] = warp_sum_d(ss[a]) / sw;
        ct[a] = warp_sum_d(sc[a]) / sw;
    }
    double cov[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) cov[i] = 0.0;
    for (int i = lane; i < n; i += 32) {
        const double wi = (double)w[i * w_sn];
        double sa[3], cb[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            sa[a] = ((double)s[a * s_sc + i * s_sn] - cs[a]) * wi;
            cb[a] = (double)c[a * c_sc + i * c_sn] - ct[a];
        }
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) cov[3 * a + b] += sa[a] * cb[b];
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) cov[i] = warp_sum_d(cov[i]);
    if (lane == 0) finish_procrustes(cov, cs, ct, rot_out + (int64_t)warp * 9, trans_out + (int64_t)warp * 3);
}

// ---------------------------------------------------------------------------------------------
// GMMSVD (is_sk = False).  Dynamic smem layout (floats):
//   sim   [Js][Jt]            similarity, then scores in place
//   xs    [Js][DT+4]          normalised src descriptor chunk
//   yt    [Jt][DT+4]          normalised tgt descriptor chunk
//   den   [Js+Jt]             max(|row|, 1e-12)
//   corr  [3][Js], wgt [Js], mus [Js][3], mut [Jt][3]
constexpr int kDT = 64;
constexpr int kDTP = kDT + 4;
constexpr int kSPThreads = 256;
constexpr int kPairsPerThread = 16;

__host__ __device__ inline size_t soft_procrustes_smem(int Js, int Jt) {
    return sizeof(float) * ((((size_t)Js * Jt + 3) & ~(size_t)3) + (size_t)(Js + Jt) * kDTP + (Js + Jt) + 3 * (size_t)Js + Js + 3 * (size_t)(Js + Jt) + 32);
}

// Steps 3-5 of GMMSVD shared by both similarity kernels: softmax(sim / T) rows, soft correspondences, weights,
// weighted Procrustes over the Js components.  `sim` holds the cosine similarities on entry (all threads synchronised).
__device__ __forceinline__ void soft_head_tail(float* sim, float* corr, float* wgt, const float* mus, const float* mut,
                                               int Js, int Jt, float temperature, int b, float* __restrict__ rot_out,
                                               float* __restrict__ trans_out, float* __restrict__ corr_out) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = kSPThreads / 32;
    // 3+4. per source row: softmax(sim / T) over j, soft correspondence, weight
    for (int i = warp; i < Js; i += NW) {
        float* row = sim + (size_t)i * Jt;
        float m = -INFINITY;
        for (int j = lane; j < Jt; j += 32) { float z = __fdiv_rn(row[j], temperature); row[j] = z; m = fmaxf(m, z); }
        m = warp_max(m);
        float sum = 0.f;
        for (int j = lane; j < Jt; j += 32) { float e = expf(row[j] - m); row[j] = e; sum += e; }
        sum = warp_sum(sum);
        float c0 = 0.f, c1 = 0.f, c2 = 0.f, w = 0.f;
        for (int j = lane; j < Jt; j += 32) {
            float sc = row[j] / sum;
            w += sc;
            c0 = fmaf(mut[3 * j], sc, c0); c1 = fmaf(mut[3 * j + 1], sc, c1); c2 = fmaf(mut[3 * j + 2], sc, c2);
        }
        c0 = warp_sum(c0); c1 = warp_sum(c1); c2 = warp_sum(c2); w = warp_sum(w);
        if (lane == 0) { corr[i] = c0; corr[Js + i] = c1; corr[2 * Js + i] = c2; wgt[i] = w; }
    }
    __syncthreads();
    for (int e = tid; e < 3 * Js; e += kSPThreads) corr_out[(int64_t)b * 3 * Js + e] = corr[e];

    // 5. weighted Procrustes over the Js components (warp 0)
    if (warp == 0) {
        double sw = 0.0, ss[3] = {0.0, 0.0, 0.0}, sc[3] = {0.0, 0.0, 0.0};
        for (int i = lane; i < Js; i += 32) {
            const double wi = (double)wgt[i];
            sw += wi;
#pragma unroll
            for (int a = 0; a < 3; ++a) { ss[a] += (double)mus[3 * i + a] * wi; sc[a] += (double)corr[a * Js + i] * wi; }
        }
        sw = warp_sum_d(sw);
        double cs[3], ct[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) { cs[a] = warp_sum_d(ss[a]) / sw; ct[a] = warp_sum_d(sc[a]) / sw; }
        double cov[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) cov[i] = 0.0;
        for (int i = lane; i < Js; i += 32) {
            const double wi = (double)wgt[i];
            double sa[3], cb[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) { sa[a] = ((double)mus[3 * i + a] - cs[a]) * wi; cb[a] = (double)corr[a * Js + i] - ct[a]; }
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int c = 0; c < 3; ++c) cov[3 * a + c] += sa[a] * cb[c];
        }
#pragma unroll
        for (int i = 0; i < 9; ++i) cov[i] = warp_sum_d(cov[i]);
        if (lane == 0) finish_procrustes(cov, cs, ct, rot_out + (int64_t)b * 9, trans_out + (int64_t)b * 3);
    }
}

__global__ void __launch_bounds__(kSPThreads)
soft_procrustes_kernel(const float* __restrict__ src_mu, const float* __restrict__ tgt_mu,
                       const float* __restrict__ src_desc, const float* __restrict__ tgt_desc,
                       int Js, int Jt, int D, float temperature,
                       float* __restrict__ rot_out, float* __restrict__ trans_out,
                       float* __restrict__ corr_out, float* __restrict__ sim_out, int head) {
    extern __shared__ __align__(16) float smem[];
    float* sim = smem;
    float* xs = sim + (((size_t)Js * Jt + 3) & ~(size_t)3);      // the descriptor tiles are read as float4: 16-byte aligned
    float* yt = xs + (size_t)Js * kDTP;
    float* den = yt + (size_t)Jt * kDTP;
    float* corr = den + (Js + Jt);
    float* wgt = corr + 3 * Js;
    float* mus = wgt + Js;
    float* mut = mus + 3 * Js;

    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = kSPThreads / 32;
    const float* xd = src_desc + (int64_t)b * Js * D;
    const float* yd = tgt_desc + (int64_t)b * Jt * D;

    // 1. row norms (F.normalize: x / max(|x|_2, 1e-12))
    for (int r = warp; r < Js + Jt; r += NW) {
        const float* row = r < Js ? xd + (int64_t)r * D : yd + (int64_t)(r - Js) * D;
        float acc = 0.f;
        for (int d = lane; d < D; d += 32) { float v = row[d]; acc += v * v; }
        acc = warp_sum(acc);
        if (lane == 0) den[r] = fmaxf(sqrtf(acc), 1e-12f);
    }
    if (head != 0) {                                   // cosine-only calls pass no centroids
        for (int i = tid; i < 3 * Js; i += kSPThreads) mus[i] = src_mu[(int64_t)b * Js * 3 + i];
        for (int i = tid; i < 3 * Jt; i += kSPThreads) mut[i] = tgt_mu[(int64_t)b * Jt * 3 + i];
    }
    __syncthreads();

    // 2. similarity, pair-block by pair-block, descriptor chunk by chunk
    const int npairs = Js * Jt;
    for (int pb = 0; pb < npairs; pb += kSPThreads * kPairsPerThread) {
        float acc[kPairsPerThread];
#pragma unroll
        for (int q = 0; q < kPairsPerThread; ++q) acc[q] = 0.f;
        for (int d0 = 0; d0 < D; d0 += kDT) {
            const int dw = min(kDT, D - d0);
            __syncthreads();
            for (int e = tid; e < (Js + Jt) * kDT; e += kSPThreads) {
                int r = e / kDT, d = e - r * kDT;
                float v = 0.f;
                if (d < dw) {
                    v = r < Js ? xd[(int64_t)r * D + d0 + d] : yd[(int64_t)(r - Js) * D + d0 + d];
                    v = v / den[r];
                }
                if (r < Js) xs[r * kDTP + d] = v; else yt[(r - Js) * kDTP + d] = v;
            }
            __syncthreads();
#pragma unroll
            for (int q = 0; q < kPairsPerThread; ++q) {
                int p = pb + tid + q * kSPThreads;
                if (p < npairs) {
                    int i = p / Jt, j = p - i * Jt;
                    const float4* xa = reinterpret_cast<const float4*>(xs + i * kDTP);
                    const float4* ya = reinterpret_cast<const float4*>(yt + j * kDTP);
                    float a = acc[q];
#pragma unroll
                    for (int d = 0; d < kDT / 4; ++d) {
                        float4 u = xa[d], v = ya[d];
                        a = fmaf(u.x, v.x, a); a = fmaf(u.y, v.y, a); a = fmaf(u.z, v.z, a); a = fmaf(u.w, v.w, a);
                    }
                    acc[q] = a;
                }
            }
        }
#pragma unroll
        for (int q = 0; q < kPairsPerThread; ++q) {
            int p = pb + tid + q * kSPThreads;
            if (p < npairs) sim[p] = acc[q];
        }
    }
    __syncthreads();
    if (sim_out != nullptr)
        for (int p = tid; p < npairs; p += kSPThreads) sim_out[(int64_t)b * npairs + p] = sim[p];
    if (head == 0) return;      // cos_similarity only
    __syncthreads();

    soft_head_tail(sim, corr, wgt, mus, mut, Js, Jt, temperature, b, rot_out, trans_out, corr_out);
}

// Same head with BOTH descriptor sets resident in shared memory ((Js + Jt) x (D + 4) floats; 66 KB at J = 16, D = 512):
// one load phase with every byte in flight at once instead of D / 64 load -> barrier -> compute rounds, whose exposed
// global-memory latency made the chunked kernel 43 us for 256 pairs.  Same arithmetic in the same order (row norm,
// x / max(|x|, 1e-12) per element, one FMA chain over d ascending), so the results are bit-identical.
__host__ __device__ inline size_t soft_procrustes_full_smem(int Js, int Jt, int D) {
    return sizeof(float) * ((size_t)Js * Jt + (size_t)(Js + Jt) * (D + 4) + (Js + Jt) + 3 * (size_t)Js + Js + 3 * (size_t)(Js + Jt) + 32);
}

__global__ void __launch_bounds__(kSPThreads)
soft_procrustes_full_kernel(const float* __restrict__ src_mu, const float* __restrict__ tgt_mu,
                            const float* __restrict__ src_desc, const float* __restrict__ tgt_desc,
                            int Js, int Jt, int D, float temperature,
                            float* __restrict__ rot_out, float* __restrict__ trans_out,
                            float* __restrict__ corr_out, float* __restrict__ sim_out, int head) {
    extern __shared__ __align__(16) float smem[];
    const int P = D + 4;                               // row pitch: D % 4 == 0, so rows stay 16-byte aligned
    float* sim = smem;
    float* rows = sim + (size_t)Js * Jt;               // [Js + Jt][P]; Js * Jt % 4 == 0 is checked by the launcher
    float* den = rows + (size_t)(Js + Jt) * P;
    float* corr = den + (Js + Jt);
    float* wgt = corr + 3 * Js;
    float* mus = wgt + Js;
    float* mut = mus + 3 * Js;
    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = kSPThreads / 32;
    const float* xd = src_desc + (int64_t)b * Js * D;
    const float* yd = tgt_desc + (int64_t)b * Jt * D;
    const int R = Js + Jt, D4 = D >> 2;

    for (int e = tid; e < R * D4; e += kSPThreads) {
        const int r = e / D4, c = e - r * D4;
        const float* row = r < Js ? xd + (int64_t)r * D : yd + (int64_t)(r - Js) * D;
        *reinterpret_cast<float4*>(rows + (size_t)r * P + 4 * c) = __ldg(reinterpret_cast<const float4*>(row) + c);
    }
    if (head != 0) {                                   // cosine-only calls pass no centroids
        for (int i = tid; i < 3 * Js; i += kSPThreads) mus[i] = src_mu[(int64_t)b * Js * 3 + i];
        for (int i = tid; i < 3 * Jt; i += kSPThreads) mut[i] = tgt_mu[(int64_t)b * Jt * 3 + i];
    }
    __syncthreads();
    // row norms (F.normalize: x / max(|x|_2, 1e-12)), lane-strided partial sums like the chunked kernel
    for (int r = warp; r < R; r += NW) {
        const float* row = rows + (size_t)r * P;
        float acc = 0.f;
        for (int d = lane; d < D; d += 32) { const float v = row[d]; acc += v * v; }
        acc = warp_sum(acc);
        if (lane == 0) den[r] = fmaxf(sqrtf(acc), 1e-12f);
    }
    __syncthreads();
    for (int e = tid; e < R * D4; e += kSPThreads) {
        const int r = e / D4, c = e - r * D4;
        float4* q = reinterpret_cast<float4*>(rows + (size_t)r * P + 4 * c);
        float4 v = *q;
        const float dn = den[r];
        v.x = v.x / dn; v.y = v.y / dn; v.z = v.z / dn; v.w = v.w / dn;
        *q = v;
    }
    __syncthreads();
    const int npairs = Js * Jt;
    for (int p = tid; p < npairs; p += kSPThreads) {
        const int i = p / Jt, j = p - i * Jt;
        const float4* xa = reinterpret_cast<const float4*>(rows + (size_t)i * P);
        const float4* ya = reinterpret_cast<const float4*>(rows + (size_t)(Js + j) * P);
        float a = 0.f;
        for (int d = 0; d < D4; ++d) {
            const float4 u = xa[d], v = ya[d];
            a = fmaf(u.x, v.x, a); a = fmaf(u.y, v.y, a); a = fmaf(u.z, v.z, a); a = fmaf(u.w, v.w, a);
        }
        sim[p] = a;
    }
    __syncthreads();
    if (sim_out != nullptr)
        for (int p = tid; p < npairs; p += kSPThreads) sim_out[(int64_t)b * npairs + p] = sim[p];
    if (head == 0) return;      // cos_similarity only
    __syncthreads();
    soft_head_tail(sim, corr, wgt, mus, mut, Js, Jt, temperature, b, rot_out, trans_out, corr_out);
}

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void inverse3(const double* m, double* inv) {
    const double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
    const double det = m[0] * c00 + m[1] * c01 + m[2] * c02;
    const double r = 1.0 / det;
    inv[0] = c00 * r; inv[1] = (m[2] * m[7] - m[1] * m[8]) * r; inv[2] = (m[1] * m[5] - m[2] * m[4]) * r;
    inv[3] = c01 * r; inv[4] = (m[0] * m[8] - m[2] * m[6]) * r; inv[5] = (m[2] * m[3] - m[0] * m[5]) * r;
    inv[6] = c02 * r; inv[7] = (m[1] * m[6] - m[0] * m[7]) * r; inv[8] = (m[0] * m[4] - m[1] * m[3]) * r;
}

__global__ void __launch_bounds__(128)
gmm_register_kernel(const float* __restrict__ pi_s, const float* __restrict__ mu_s, const float* __restrict__ mu_t,
                    const float* __restrict__ sigma_t, int B, int J, float* __restrict__ tf_out) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= B) return;
    const float* pi = pi_s + (int64_t)warp * J;
    const float* ms = mu_s + (int64_t)warp * J * 3;
    const float* mt = mu_t + (int64_t)warp * J * 3;
    const float* sg = sigma_t + (int64_t)warp * J * 9;
    double cs[3] = {0.0, 0.0, 0.0}, ct[3] = {0.0, 0.0, 0.0};
    for (int j = lane; j < J; j += 32) {
        const double p = (double)pi[j];
#pragma unroll
        for (int a = 0; a < 3; ++a) { cs[a] += p * (double)ms[3 * j + a]; ct[a] += p * (double)mt[3 * j + a]; }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) { cs[a] = warp_sum_d(cs[a]); ct[a] = warp_sum_d(ct[a]); }
    double M[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) M[i] = 0.0;
    for (int j = lane; j < J; j += 32) {
        const double p = (double)pi[j];
        double a[3], bt[3], sgd[9], inv[9];
#pragma unroll
        for (int c = 0; c < 3; ++c) { a[c] = p * ((double)ms[3 * j + c] - cs[c]); bt[c] = (double)mt[3 * j + c] - ct[c]; }
#pragma unroll
        for (int c = 0; c < 9; ++c) sgd[c] = (double)sg[9 * j + c];
        inverse3(sgd, inv);
        // (a b^T) Sigma^-1 = a (b^T Sigma^-1)
        double row[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) row[c] = bt[0] * inv[c] + bt[1] * inv[3 + c] + bt[2] * inv[6 + c];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) M[3 * r + c] += a[r] * row[c];
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) M[i] = warp_sum_d(M[i]);
    if (lane == 0) {
        double Md[9], R[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            double m = M[i];
            if (m != m) m = 0.0;                                   // nan_to_num(Ms, nan=0) + 1e-4 on all nine entries
            else if (m > 3.402823466e+38) m = 3.402823466e+38;
            else if (m < -3.402823466e+38) m = -3.402823466e+38;
            Md[i] = m + 1e-4;
        }
        rotation_from_cov_deepgmr(Md, R);
        float* T = tf_out + (int64_t)warp * 16;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            T[4 * r] = (float)R[3 * r]; T[4 * r + 1] = (float)R[3 * r + 1]; T[4 * r + 2] = (float)R[3 * r + 2];
            T[4 * r + 3] = (float)(ct[r] - (R[3 * r] * cs[0] + R[3 * r + 1] * cs[1] + R[3 * r + 2] * cs[2]));
        }
        T[12] = 0.f; T[13] = 0.f; T[14] = 0.f; T[15] = 1.f;
    }
}

}  // namespace ogmm

using namespace ogmm;

extern "C" __attribute__((visibility("default"))) int ogmm_rigid_transform(const float* src, int64_t s_sb, int64_t s_sc, int64_t s_sn,
                                    const float* corr, int64_t c_sb, int64_t c_sc, int64_t c_sn,
                                    const float* weight, int64_t w_sb, int64_t w_sn,
                                    int64_t B, int64_t n, float* rot_out, float* trans_out, ogmm_stream_t stream) {
    OGMM_REQUIRE(B >= 0 && n >= 1 && B < (1ll << 31) && n < (1ll << 31), OGMM_EINVAL,
                 "ogmm_rigid_transform: bad sizes B=%lld n=%lld", (long long)B, (long long)n);
    if (B == 0) return OGMM_OK;
    OGMM_REQUIRE(src && corr && weight && rot_out && trans_out, OGMM_EINVAL, "ogmm_rigid_transform: null pointer");
    const int threads = 128;
    const int blocks = (int)((B * 32 + threads - 1) / threads);
    rigid_transform_kernel<<<blocks, threads, 0, as_stream(stream)>>>(src, s_sb, s_sc, s_sn, corr, c_sb, c_sc, c_sn,
                                                                      weight, w_sb, w_sn, (int)B, (int)n, rot_out, trans_out);
    OGMM_LAUNCH_CHECK("rigid_transform_kernel");
    return OGMM_OK;
}

static int launch_soft(const float* src_mu, const float* tgt_mu, const float* src_desc, const float* tgt_desc,
                       int64_t B, int64_t Js, int64_t Jt, int64_t D, float temperature, float* rot_out,
                       float* trans_out, float* corr_out, float* sim_out, int head, ogmm_stream_t stream) {
    // both descriptor sets resident in shared memory when they fit three CTAs per SM; the chunked kernel otherwise
    const size_t full = soft_procrustes_full_smem((int)Js, (int)Jt, (int)D);
    const bool aligned = ((reinterpret_cast<uintptr_t>(src_desc) | reinterpret_cast<uintptr_t>(tgt_desc)) & 15) == 0;
    if ((D & 3) == 0 && ((Js * Jt) & 3) == 0 && aligned && full <= 72 * 1024) {
        int st = cuda_status(cudaFuncSetAttribute(soft_procrustes_full_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                  (int)full), "cudaFuncSetAttribute(soft_procrustes_full_kernel)");
        if (st != OGMM_OK) return st;
        soft_procrustes_full_kernel<<<(unsigned)B, kSPThreads, full, as_stream(stream)>>>(
            src_mu, tgt_mu, src_desc, tgt_desc, (int)Js, (int)Jt, (int)D, temperature, rot_out, trans_out, corr_out,
            sim_out, head);
        OGMM_LAUNCH_CHECK("soft_procrustes_full_kernel");
        return OGMM_OK;
    }
    size_t smem = soft_procrustes_smem((int)Js, (int)Jt);
    OGMM_REQUIRE(smem <= 200 * 1024, OGMM_EUNSUPPORTED,
                 "soft_procrustes: Js=%lld Jt=%lld needs %zu B of shared memory (> 200 KiB)", (long long)Js,
                 (long long)Jt, smem);
    int st = cuda_status(cudaFuncSetAttribute(soft_procrustes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)smem), "cudaFuncSetAttribute(soft_procrustes_kernel)");
    if (st != OGMM_OK) return st;
    soft_procrustes_kernel<<<(unsigned)B, kSPThreads, smem, as_stream(stream)>>>(
        src_mu, tgt_mu, src_desc, tgt_desc, (int)Js, (int)Jt, (int)D, temperature, rot_out, trans_out, corr_out,
        sim_out, head);
    OGMM_LAUNCH_CHECK("soft_procrustes_kernel");
    return OGMM_OK;
}

extern "C" __attribute__((visibility("default"))) int ogmm_soft_procrustes(const float* src_mu, const float* tgt_mu, const float* src_desc,
                                    const float* tgt_desc, int64_t B, int64_t Js, int64_t Jt, int64_t D,
                                    float temperature, float* rot_out, float* trans_out, float* corr_out,
                                    float* sim_out, ogmm_stream_t stream) {
    OGMM_REQUIRE(B >= 0 && Js >= 1 && Jt >= 1 && D >= 1 && B < (1ll << 31), OGMM_EINVAL,
                 "ogmm_soft_procrustes: bad sizes B=%lld Js=%lld Jt=%lld D=%lld", (long long)B, (long long)Js,
                 (long long)Jt, (long long)D);
    if (B == 0) return OGMM_OK;
    OGMM_REQUIRE(src_mu && tgt_mu && src_desc && tgt_desc && rot_out && trans_out && corr_out, OGMM_EINVAL,
                 "ogmm_soft_procrustes: null pointer");
    OGMM_REQUIRE(temperature > 0.f, OGMM_EINVAL, "ogmm_soft_procrustes: temperature must be > 0");
    return launch_soft(src_mu, tgt_mu, src_desc, tgt_desc, B, Js, Jt, D, temperature, rot_out, trans_out, corr_out,
                       sim_out, 1, stream);
}

extern "C" __attribute__((visibility("default"))) int ogmm_cos_similarity(const float* x, const float* y, int64_t B, int64_t N, int64_t M, int64_t D,
                                   float* sim_out, ogmm_stream_t stream) {
    OGMM_REQUIRE(B >= 0 && N >= 1 && M >= 1 && D >= 1 && B < (1ll << 31), OGMM_EINVAL,
                 "ogmm_cos_similarity: bad sizes");
    if (B == 0) return OGMM_OK;
    OGMM_REQUIRE(x && y && sim_out, OGMM_EINVAL, "ogmm_cos_similarity: null pointer");
    return launch_soft(nullptr, nullptr, x, y, B, N, M, D, 1.f, nullptr, nullptr, nullptr, sim_out, 0, stream);
}

extern "C" __attribute__((visibility("default"))) int ogmm_gmm_register(const float* pi_s, const float* mu_s, const float* mu_t, const float* sigma_t,
                                 int64_t B, int64_t J, float* transform_out, ogmm_stream_t stream) {
    OGMM_REQUIRE(B >= 0 && J >= 1 && B < (1ll << 31) && J < (1ll << 31), OGMM_EINVAL, "ogmm_gmm_register: bad sizes");
    if (B == 0) return OGMM_OK;
    OGMM_REQUIRE(pi_s && mu_s && mu_t && sigma_t && transform_out, OGMM_EINVAL, "ogmm_gmm_register: null pointer");
    const int threads = 128;
    const int blocks = (int)((B * 32 + threads - 1) / threads);
    gmm_register_kernel<<<blocks, threads, 0, as_stream(stream)>>>(pi_s, mu_s, mu_t, sigma_t, (int)B, (int)J,
                                                                   transform_out);
    OGMM_LAUNCH_CHECK("gmm_register_kernel");
    return OGMM_OK;
}
