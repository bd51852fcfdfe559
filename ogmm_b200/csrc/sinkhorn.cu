// Host entry point: stand-alone Sinkhorn on a cost matrix (kernels in sinkhorn_impl.cuh).
#include "sinkhorn_impl.cuh"

extern "C" __attribute__((visibility("default"))) int ogmm_sinkhorn(const float* cost, const float* p, const float* q, int64_t B, int64_t N, int64_t M,
                             float epsilon, float thresh, int64_t max_iter, float* gamma_out, float* loss_out,
                             int32_t* iters_run_out, void* workspace, int64_t workspace_bytes, ogmm_stream_t stream) {
    OGMM_REQUIRE(B >= 0 && N >= 1 && M >= 1 && max_iter >= 1 && B < (1ll << 31), OGMM_EINVAL, "ogmm_sinkhorn: bad sizes");
    OGMM_REQUIRE(N <= kMaxPoints && M <= 4096, OGMM_EUNSUPPORTED, "ogmm_sinkhorn: need N <= %lld and M <= 4096",
                 (long long)kMaxPoints);
    OGMM_REQUIRE(max_iter <= 4096, OGMM_EUNSUPPORTED, "ogmm_sinkhorn: max_iter <= 4096");
    OGMM_REQUIRE(epsilon > 0.f, OGMM_EINVAL, "ogmm_sinkhorn: epsilon must be > 0");
    if (B == 0) return OGMM_OK;
    OGMM_REQUIRE(cost && gamma_out, OGMM_EINVAL, "ogmm_sinkhorn: null pointer");
    SinkhornParams P{};
    P.cost = cost; P.p = p; P.q = q;
    P.B = (int)B; P.N = (int)N; P.J = (int)M; P.iters = 1; P.max_iter = (int)max_iter;
    P.tau = 1.f; P.eps = epsilon; P.thresh = thresh;
    P.gamma = gamma_out; P.loss = loss_out; P.iters_run = iters_run_out;
    return launch_sinkhorn<false>(P, workspace, workspace_bytes, stream);
}
