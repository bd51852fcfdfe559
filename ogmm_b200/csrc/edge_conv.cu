// N3: edge gather fused into the first EdgeConv layer of DGCNN (sm_100a).
//
// Replaces, for inference, the opening of DGCNN.forward (models/dgcnn.py:135-141):
//
//     x  = get_graph_feature(x, k, idx)            (B,6,N,k)   [x_j - x_i ; x_i]        lib/utils.py:47-66
//     x  = relu(bn1(conv1(x)))                     (B,C,N,k)   1x1 Conv2d(6 -> C, bias=False) + BatchNorm2d (eval)
//     x1 = x.max(dim=-1, keepdim=True)[0]          (B,C,N,1)
//
// The (B,6,N,k) edge tensor is never materialised: each thread gathers the k neighbours of one point from the cloud
// (staged in shared memory), keeps the k offsets x_j - x_i in registers and evaluates its 8 output channels from
// them.  BatchNorm in eval mode is the per-channel affine map y -> y * scale + shift with
// scale = gamma / sqrt(running_var + eps), shift = beta - running_mean * scale, folded by the caller.
//
//     act[b,c,n,kk] = relu((W[c,0:3] . (x_j - x_i) + W[c,3:6] . x_i) * scale[c] + shift[c])
//     max[b,c,n]    = max_kk act[b,c,n,kk]
//
// `act` (what conv2 of the reference consumes) is optional: with act == NULL only the 256-byte-per-point pooled
// output is written.  Algorithmic bytes per cloud: 12 N + 8 N k in, 4 C N (+ 4 C N k) out -- HBM write bound when act
// is requested (5.2 MB per cloud at N=1024, k=20, C=64), latency bound otherwise.
//
// CTA = 32 points x (C / 8) channel groups: lane <-> point, warp <-> 8 channels, so the weights are warp-uniform
// shared-memory broadcasts and each lane's act row (k contiguous floats per channel) is written with 16-byte stores.
#include "common.cuh"

namespace ogmm {

constexpr int kEcPts = 32;        // points per CTA (one per lane)
constexpr int kEcCh = 8;          // channels per warp
constexpr int kEcMaxK = 32;

template <int KP>                 // neighbours held in registers (k <= KP)
__global__ void __launch_bounds__(256)
edge_conv_max_kernel(const float* __restrict__ x, int64_t x_sb, int64_t x_sc, int64_t x_sn,
                     const int64_t* __restrict__ idx, const float* __restrict__ weight,
                     const float* __restrict__ scale, const float* __restrict__ shift,
                     int N, int k, int C, float* __restrict__ act, float* __restrict__ pooled) {
    extern __shared__ __align__(16) float ec_sm[];
    float* s_w = ec_sm;                      // [C][8]: W[c][0..5], scale, shift
    float* s_xyz = ec_sm + (size_t)C * 8;    // [N][3] the whole cloud (neighbour gathers)
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwarps = blockDim.x >> 5;
    const float* xb = x + (int64_t)b * x_sb;
    for (int e = tid; e < C; e += blockDim.x) {
#pragma unroll
        for (int i = 0; i < 6; ++i) s_w[e * 8 + i] = weight[e * 6 + i];
        s_w[e * 8 + 6] = scale[e];
        s_w[e * 8 + 7] = shift[e];
    }
    for (int e = tid; e < 3 * N; e += blockDim.x) {
        const int c = e / N, n = e - c * N;                      // coalesced along n for the native (B,3,N) layout
        s_xyz[3 * n + c] = xb[(int64_t)c * x_sc + (int64_t)n * x_sn];
    }
    __syncthreads();

    const int n = blockIdx.x * kEcPts + lane;
    const bool valid = n < N;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    float dx[KP], dy[KP], dz[KP];
    if (valid) {
        qx = s_xyz[3 * n]; qy = s_xyz[3 * n + 1]; qz = s_xyz[3 * n + 2];
        const int64_t* ip = idx + ((int64_t)b * N + n) * k;
#pragma unroll
        for (int kk = 0; kk < KP; ++kk) {
            dx[kk] = dy[kk] = dz[kk] = 0.f;
            if (kk < k) {
                int64_t j = ip[kk];
                j = j < 0 ? 0 : (j >= N ? N - 1 : j);
                dx[kk] = s_xyz[3 * j] - qx; dy[kk] = s_xyz[3 * j + 1] - qy; dz[kk] = s_xyz[3 * j + 2] - qz;
            }
        }
    }
    for (int c0 = warp * kEcCh; c0 < C; c0 += nwarps * kEcCh) {
#pragma unroll
        for (int cc = 0; cc < kEcCh; ++cc) {
            const int c = c0 + cc;
            if (c >= C) break;                                       // warp-uniform
            const float4 wa = *reinterpret_cast<const float4*>(s_w + c * 8);
            const float4 wb = *reinterpret_cast<const float4*>(s_w + c * 8 + 4);
            // the centre part of the 1x1 convolution is the same for every neighbour; the accumulation order over the six
            // input channels is the convolution's own (0..5): ((((w0 d0 + w1 d1) + w2 d2) + w3 q0) + w4 q1) + w5 q2
            float v[KP];
            float best = 0.f;                                       // ReLU output is >= 0
#pragma unroll
            for (int kk = 0; kk < KP; ++kk) {
                float a = __fmul_rn(wa.x, dx[kk]);
                a = fmaf(wa.y, dy[kk], a);
                a = fmaf(wa.z, dz[kk], a);
                a = fmaf(wa.w, qx, a);
                a = fmaf(wb.x, qy, a);
                a = fmaf(wb.y, qz, a);
                a = fmaxf(fmaf(a, wb.z, wb.w), 0.f);
                v[kk] = a;
                if (kk < k) best = fmaxf(best, a);
            }
            if (!valid) continue;
            pooled[((int64_t)b * C + c) * N + n] = best;
            if (act != nullptr) {
                float* o = act + (((int64_t)b * C + c) * N + n) * k;
                if ((k & 3) == 0 && (reinterpret_cast<uintptr_t>(act) & 15) == 0) {
#pragma unroll
                    for (int kk = 0; kk < KP; kk += 4)
                        if (kk < k) *reinterpret_cast<float4*>(o + kk) = make_float4(v[kk], v[kk + 1], v[kk + 2], v[kk + 3]);
                } else {
#pragma unroll
                    for (int kk = 0; kk < KP; ++kk)
                        if (kk < k) o[kk] = v[kk];
                }
            }
        }
    }
}


// ---------------------------------------------------------------------------------------------
// N3, second half: the k = 5 angle feature of PositionEncoding.forward (models/attn.py:65-73) in one launch:
//
//     p2gc  = points - mean_n(points)                                   (B,3,N)
//     p2lc  = get_graph_feature(points, k)[:, :3]                       (B,3,N,k)    x_j - x_i
//     alpha = <normalize(p2lc, dim=1), normalize(p2gc, dim=1)>          (B,1,N,k)    F.normalize: v / max(|v|, 1e-12)
//     out   = conv_ang1(alpha).max(dim=-1)[0]                           (B,C,N)      Conv2d(1 -> C, bias=False) + BN (eval) + LeakyReLU
//
// Neither the (B,6,N,k) edge tensor nor the (B,C,N,k) activation is materialised.  Per channel the map
// alpha -> leaky(fma(w_c * alpha, scale_c, shift_c)) is monotone (rounded products and sums are monotone), so the max
// over the k neighbours is that map applied to the largest alpha (w_c * scale_c >= 0) or the smallest one (< 0): each
// thread keeps two numbers per point and evaluates C channels from them.  lane <-> point: every channel row is written
// as 128 contiguous bytes per warp.
__global__ void __launch_bounds__(256)
edge_angle_max_kernel(const float* __restrict__ x, int64_t x_sb, int64_t x_sc, int64_t x_sn,
                      const float* __restrict__ centroid, const int64_t* __restrict__ idx,
                      const float* __restrict__ weight, const float* __restrict__ scale, const float* __restrict__ shift,
                      float slope, int N, int k, int C, float* __restrict__ alpha_out, float* __restrict__ pooled) {
    extern __shared__ __align__(16) float ea_sm[];
    float* s_w = ea_sm;                      // [C][4]: w, scale, shift, pad
    float* s_xyz = ea_sm + (size_t)C * 4;    // [N][3]
    const int b = blockIdx.y, tid = threadIdx.x;
    const float* xb = x + (int64_t)b * x_sb;
    for (int e = tid; e < C; e += blockDim.x) {
        s_w[e * 4] = weight[e]; s_w[e * 4 + 1] = scale[e]; s_w[e * 4 + 2] = shift[e]; s_w[e * 4 + 3] = 0.f;
    }
    for (int e = tid; e < 3 * N; e += blockDim.x) {
        const int c = e / N, n = e - c * N;
        s_xyz[3 * n + c] = xb[(int64_t)c * x_sc + (int64_t)n * x_sn];
    }
    __syncthreads();
    const int n = blockIdx.x * blockDim.x + tid;
    if (n >= N) return;
    const float qx = s_xyz[3 * n], qy = s_xyz[3 * n + 1], qz = s_xyz[3 * n + 2];
    // normalize(p2gc): three independent divisions by max(|v|, 1e-12), as F.normalize does
    float gx = qx - centroid[3 * b], gy = qy - centroid[3 * b + 1], gz = qz - centroid[3 * b + 2];
    {
        const float gn = fmaxf(sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)), __fmul_rn(gz, gz))), 1e-12f);
        gx = __fdiv_rn(gx, gn); gy = __fdiv_rn(gy, gn); gz = __fdiv_rn(gz, gn);
    }
    float amax = -INFINITY, amin = INFINITY;
    const int64_t* ip = idx + ((int64_t)b * N + n) * k;
    for (int kk = 0; kk < k; ++kk) {
        int64_t j = ip[kk];
        j = j < 0 ? 0 : (j >= N ? N - 1 : j);
        float dx = s_xyz[3 * j] - qx, dy = s_xyz[3 * j + 1] - qy, dz = s_xyz[3 * j + 2] - qz;
        const float dn = fmaxf(sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz))), 1e-12f);
        dx = __fdiv_rn(dx, dn); dy = __fdiv_rn(dy, dn); dz = __fdiv_rn(dz, dn);
        const float a = __fadd_rn(__fadd_rn(__fmul_rn(dx, gx), __fmul_rn(dy, gy)), __fmul_rn(dz, gz));
        if (alpha_out != nullptr) alpha_out[((int64_t)b * N + n) * k + kk] = a;
        amax = fmaxf(amax, a); amin = fminf(amin, a);
    }
    for (int c = 0; c < C; ++c) {
        const float4 w = *reinterpret_cast<const float4*>(s_w + c * 4);
        // the side that wins the max: the affine map has the sign of w * scale (LeakyReLU is increasing for slope >= 0)
        const bool up = (w.x >= 0.f) == (w.y >= 0.f);
        float v = fmaf(__fmul_rn(w.x, up ? amax : amin), w.y, w.z);
        v = v >= 0.f ? v : __fmul_rn(v, slope);
        pooled[((int64_t)b * C + c) * N + n] = v;
    }
}

}  // namespace ogmm

using namespace ogmm;

extern "C" __attribute__((visibility("default"))) int ogmm_edge_conv_max(
    const float* x, int64_t x_sb, int64_t x_sc, int64_t x_sn, const int64_t* idx, const float* weight, const float* scale,
    const float* shift, int64_t B, int64_t N, int64_t k, int64_t C, float* act_out, float* max_out, ogmm_stream_t stream) {
    OGMM_REQUIRE(B >= 0 && N >= 1 && k >= 1 && C >= 1 && B < 65536 && N < (1ll << 24), OGMM_EINVAL,
                 "ogmm_edge_conv_max: bad sizes B=%lld N=%lld k=%lld C=%lld", (long long)B, (long long)N, (long long)k, (long long)C);
    OGMM_REQUIRE(k <= kEcMaxK, OGMM_EUNSUPPORTED, "ogmm_edge_conv_max: k=%lld > %d", (long long)k, kEcMaxK);
    if (B == 0) return OGMM_OK;
    OGMM_REQUIRE(x && idx && weight && scale && shift && max_out, OGMM_EINVAL, "ogmm_edge_conv_max: null pointer");
    const size_t smem = sizeof(float) * ((size_t)C * 8 + (size_t)3 * N);
    OGMM_REQUIRE(smem <= 200 * 1024, OGMM_EUNSUPPORTED, "ogmm_edge_conv_max: N=%lld, C=%lld need %zu B of shared memory", (long long)N,
                 (long long)C, smem);
    int threads = (int)((C + kEcCh - 1) / kEcCh) * 32;
    if (threads > 256) threads = 256;
    dim3 grid((unsigned)((N + kEcPts - 1) / kEcPts), (unsigned)B);
    cudaStream_t s = as_stream(stream);
#define LAUNCH(KP)                                                                                                      \
    do {                                                                                                                \
        if (smem > 48 * 1024) {                                                                                         \
            int st = cuda_status(cudaFuncSetAttribute(edge_conv_max_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                                      (int)smem), "cudaFuncSetAttribute(edge_conv_max_kernel)");        \
            if (st != OGMM_OK) return st;                                                                               \
        }                                                                                                               \
        edge_conv_max_kernel<KP><<<grid, threads, smem, s>>>(x, x_sb, x_sc, x_sn, idx, weight, scale, shift, (int)N,    \
                                                            (int)k, (int)C, act_out, max_out);                         \
    } while (0)
    if (k <= 8) LAUNCH(8);
    else if (k <= 20) LAUNCH(20);
    else LAUNCH(32);
#undef LAUNCH
    OGMM_LAUNCH_CHECK("edge_conv_max_kernel");
    return OGMM_OK;
}

extern "C" __attribute__((visibility("default"))) int ogmm_edge_angle_max(
    const float* x, int64_t x_sb, int64_t x_sc, int64_t x_sn, const float* centroid, const int64_t* idx, const float* weight,
    const float* scale, const float* shift, float slope, int64_t B, int64_t N, int64_t k, int64_t C, float* alpha_out,
    float* max_out, ogmm_stream_t stream) {
    OGMM_REQUIRE(B >= 0 && N >= 1 && k >= 1 && C >= 1 && B < 65536 && N < (1ll << 24), OGMM_EINVAL,
                 "ogmm_edge_angle_max: bad sizes B=%lld N=%lld k=%lld C=%lld", (long long)B, (long long)N, (long long)k, (long long)C);
    OGMM_REQUIRE(slope >= 0.f, OGMM_EINVAL, "ogmm_edge_angle_max: negative LeakyReLU slope %g (the max shortcut needs a monotone activation)", (double)slope);
    if (B == 0) return OGMM_OK;
    OGMM_REQUIRE(x && centroid && idx && weight && scale && shift && max_out, OGMM_EINVAL, "ogmm_edge_angle_max: null pointer");
    const size_t smem = sizeof(float) * ((size_t)C * 4 + (size_t)3 * N);
    OGMM_REQUIRE(smem <= 200 * 1024, OGMM_EUNSUPPORTED, "ogmm_edge_angle_max: N=%lld, C=%lld need %zu B of shared memory", (long long)N,
                 (long long)C, smem);
    if (smem > 48 * 1024) {
        int st = cuda_status(cudaFuncSetAttribute(edge_angle_max_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                             "cudaFuncSetAttribute(edge_angle_max_kernel)");
        if (st != OGMM_OK) return st;
    }
    const int threads = 128;
    dim3 grid((unsigned)((N + threads - 1) / threads), (unsigned)B);
    edge_angle_max_kernel<<<grid, threads, smem, as_stream(stream)>>>(x, x_sb, x_sc, x_sn, centroid, idx, weight, scale, shift,
                                                                      slope, (int)N, (int)k, (int)C, alpha_out, max_out);
    OGMM_LAUNCH_CHECK("edge_angle_max_kernel");
    return OGMM_OK;
}
