// Host entry points: FPS and the overlap-guided Sinkhorn k-means (kernels in sinkhorn_impl.cuh).
#include "sinkhorn_impl.cuh"

extern "C" __attribute__((visibility("default"))) int ogmm_fps(const float* xyz, int64_t sb, int64_t sn, int64_t sc, int64_t B, int64_t N, int64_t npoint,
                        const int64_t* start, int64_t* ids_out, float* pts_out, ogmm_stream_t stream) {
    OGMM_REQUIRE(B >= 0 && N >= 1 && npoint >= 0 && B < (1ll << 31), OGMM_EINVAL, "ogmm_fps: bad sizes");
    OGMM_REQUIRE(N <= kMaxPoints, OGMM_EUNSUPPORTED, "ogmm_fps: N=%lld > %lld", (long long)N, (long long)kMaxPoints);
    if (B == 0 || npoint == 0) return OGMM_OK;
    OGMM_REQUIRE(xyz && ids_out, OGMM_EINVAL, "ogmm_fps: null pointer");
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

