// Clustering variant for clouds of 8193..16384 points (cfg 4): 1024 threads x 16 points per thread, log-domain path.
// Register-starved (64 registers per thread at 1024 threads), so it spills and is slow; the multi-CTA variant with
// distributed-shared-memory folds is the planned replacement.  Own translation unit on purpose.
#include "sinkhorn_impl.cuh"

int ogmm_launch_cluster_big(ogmm::SinkhornParams P, cudaStream_t s) {
    return launch_sinkhorn_variant<1024, 16, true, false>(P, s);
}
