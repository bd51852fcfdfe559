// Host entry points: FPS and the overlap-guided Sinkhorn k-means (kernels in sinkhorn_impl.cuh).
#include "sinkhorn_impl.cuh"

namespace ogmm {

// Clouds above kMaxPoints (up to 16384 points, the clustering's own limit): the coordinates live in shared memory
// (SoA, 12 bytes per point), only the running minimum distances stay in registers (16 per thread at 1024 threads).
// Same arithmetic, same 64-bit (distance, ~index) arg-max keys, same tie-break as fps_kernel.
constexpr int kFpsBigThreads = 1024;
constexpr int kFpsBigPpt = 16;
constexpr int64_t kMaxFpsPoints = (int64_t)kFpsBigThreads * kFpsBigPpt;

__global__ void __launch_bounds__(kFpsBigThreads)
fps_smem_kernel(const float* __restrict__ xyz, int64_t sb, int64_t sn, int64_t sc, int N, int npoint,
                const int64_t* __restrict__ start, int64_t* __restrict__ ids_out, float* __restrict__ pts_out) {
    extern __shared__ __align__(16) float fps_pts[];
    __shared__ unsigned long long s_key[32];
    __shared__ float s_red[32];
    constexpr int NT = kFpsBigThreads, PPT = kFpsBigPpt, NW = NT / 32;
    const int Np = (N + 31) & ~31;
    float* sx = fps_pts; float* sy = fps_pts + Np; float* sz = fps_pts + 2 * Np;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* base = xyz + (int64_t)b * sb;
    for (int i = tid; i < N; i += NT) {
        sx[i] = base[(int64_t)i * sn]; sy[i] = base[(int64_t)i * sn + sc]; sz[i] = base[(int64_t)i * sn + 2 * sc];
    }
    float best[PPT];
#pragma unroll
    for (int p = 0; p < PPT; ++p) best[p] = 1e10f;
    __syncthreads();

    auto relax = [&](float cx, float cy, float cz) -> int {
        unsigned long long key = 0ull;
#pragma unroll
        for (int p = 0; p < PPT; ++p) {
            const int i = tid + p * NT;
            if (i < N) {
                const float d = sq3(sx[i] - cx, sy[i] - cy, sz[i] - cz);      // sum((xyz - c) ** 2, -1): no FMA
                if (d < best[p]) best[p] = d;
                const unsigned long long k = far_key(best[p], i);
                key = k > key ? k : key;
            }
        }
        key = warp_max_u64(key);
        __syncthreads();                       // previous readers of s_key are done
        if (lane == 0) s_key[warp] = key;
        __syncthreads();
        unsigned long long k2 = lane < NW ? s_key[lane] : 0ull;
        k2 = warp_max_u64(k2);
        return (int)(0xffffffffu - (unsigned)(k2 & 0xffffffffull));
    };

    int far;
    if (start == nullptr) {                    // is_center=True (lib/utils.py:183-188): farthest point from the centroid
        float ax = 0.f, ay = 0.f, az = 0.f;
#pragma unroll
        for (int p = 0; p < PPT; ++p) {
            const int i = tid + p * NT;
            if (i < N) { ax += sx[i]; ay += sy[i]; az += sz[i]; }
        }
        const float cx = block_sum<NT>(ax, s_red) / (float)N;
        const float cy = block_sum<NT>(ay, s_red) / (float)N;
        const float cz = block_sum<NT>(az, s_red) / (float)N;
        far = relax(cx, cy, cz);
    } else {
        far = (int)start[b];
        far = far < 0 ? 0 : (far >= N ? N - 1 : far);
    }
    int64_t* ids = ids_out + (int64_t)b * npoint;
    float* pts = pts_out ? pts_out + (int64_t)b * npoint * 3 : nullptr;
    for (int s = 0; s < npoint; ++s) {         // lib/utils.py:191-197
        const float cx = sx[far], cy = sy[far], cz = sz[far];
        if (tid == 0) {
            ids[s] = far;
            if (pts) { pts[3 * s] = cx; pts[3 * s + 1] = cy; pts[3 * s + 2] = cz; }
        }
        far = relax(cx, cy, cz);
    }
}

}  // namespace ogmm

using namespace ogmm;

extern "C" __attribute__((visibility("default"))) int ogmm_fps(const float* xyz, int64_t sb, int64_t sn, int64_t sc, int64_t B, int64_t N, int64_t npoint,
                        const int64_t* start, int64_t* ids_out, float* pts_out, ogmm_stream_t stream) {
    OGMM_REQUIRE(B >= 0 && N >= 1 && npoint >= 0 && B < (1ll << 31), OGMM_EINVAL, "ogmm_fps: bad sizes");
    OGMM_REQUIRE(N <= kMaxFpsPoints, OGMM_EUNSUPPORTED, "ogmm_fps: N=%lld > %lld", (long long)N, (long long)kMaxFpsPoints);
    if (B == 0 || npoint == 0) return OGMM_OK;
    OGMM_REQUIRE(xyz && ids_out, OGMM_EINVAL, "ogmm_fps: null pointer");
    if (N > kMaxPoints) {                      // coordinates in shared memory instead of registers
        const size_t smem = sizeof(float) * 3 * (size_t)((N + 31) & ~31ll);
        int st = cuda_status(cudaFuncSetAttribute(fps_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                             "cudaFuncSetAttribute(fps_smem_kernel)");
        if (st != OGMM_OK) return st;
        fps_smem_kernel<<<(unsigned)B, kFpsBigThreads, smem, as_stream(stream)>>>(xyz, sb, sn, sc, (int)N, (int)npoint, start,
                                                                                 ids_out, pts_out);
        OGMM_LAUNCH_CHECK("fps_smem_kernel");
        return OGMM_OK;
    }
#define CALL(NT, PPT)                                                                                              \
    fps_kernel<NT, PPT><<<(unsigned)B, NT, 0, as_stream(stream)>>>(xyz, sb, sn, sc, (int)N, (int)npoint, start,    \
                                                                   ids_out, pts_out)
    OGMM_DISPATCH_POINTS(N, CALL);
#undef CALL
    OGMM_LAUNCH_CHECK("fps_kernel");
    return OGMM_OK;
}

extern "C" __attribute__((visibility("default"))) int64_t ogmm_sinkhorn_cluster_workspace(int64_t B, int64_t N, int64_t J, int64_t iters, int64_t max_iter) {
    (void)N;
    if (B < 0 || J < 1 || iters < 1 || max_iter < 1) return 0;
    return cluster_ws_layout(B, J, iters, max_iter).total;
}
extern "C" __attribute__((visibility("default"))) int64_t ogmm_sinkhorn_workspace(int64_t B, int64_t N, int64_t M, int64_t max_iter) {
    (void)N; (void)M;
    if (B < 0 || max_iter < 1) return 0;
    return cluster_ws_layout(B, 1, 1, max_iter).total;
}

extern "C" __attribute__((visibility("default"))) int ogmm_sinkhorn_cluster(const float* xyz, int64_t sb, int64_t sn, int64_t sc, const float* o_scores,
                                     int64_t B, int64_t N, int64_t J, int64_t iters, float tau, float epsilon,
                                     float thresh, int64_t max_iter, float* gamma_out, float* pi_out, float* mu_out,
                                     int32_t* iters_run_out, void* workspace, int64_t workspace_bytes,
                                     ogmm_stream_t stream) {
    OGMM_REQUIRE(B >= 0 && N >= 1 && J >= 1 && iters >= 1 && max_iter >= 1 && B < (1ll << 31), OGMM_EINVAL,
                 "ogmm_sinkhorn_cluster: bad sizes B=%lld N=%lld J=%lld iters=%lld max_iter=%lld", (long long)B,
                 (long long)N, (long long)J, (long long)iters, (long long)max_iter);
    OGMM_REQUIRE(N <= kMaxClusterPoints, OGMM_EUNSUPPORTED, "ogmm_sinkhorn_cluster: N=%lld > %lld", (long long)N, (long long)kMaxClusterPoints);
    OGMM_REQUIRE(J <= N && J <= 1024, OGMM_EUNSUPPORTED, "ogmm_sinkhorn_cluster: need J <= min(N, 1024), got J=%lld N=%lld",
                 (long long)J, (long long)N);
    OGMM_REQUIRE(iters <= 64 && max_iter <= 1024, OGMM_EUNSUPPORTED, "ogmm_sinkhorn_cluster: iters <= 64 and max_iter <= 1024");
    OGMM_REQUIRE(epsilon > 0.f && tau > 0.f, OGMM_EINVAL, "ogmm_sinkhorn_cluster: epsilon and tau must be > 0");
    if (B == 0) return OGMM_OK;
    OGMM_REQUIRE(xyz && o_scores && gamma_out && pi_out && mu_out, OGMM_EINVAL, "ogmm_sinkhorn_cluster: null pointer");
    SinkhornParams P{};
    P.xyz = xyz; P.sb = sb; P.sn = sn; P.sc = sc; P.o_scores = o_scores;
    P.B = (int)B; P.N = (int)N; P.J = (int)J; P.iters = (int)iters; P.max_iter = (int)max_iter;
    P.tau = tau; P.eps = epsilon; P.thresh = thresh;
    P.gamma = gamma_out; P.pi = pi_out; P.mu = mu_out; P.iters_run = iters_run_out;
    return launch_sinkhorn<true>(P, workspace, workspace_bytes, stream);
}

