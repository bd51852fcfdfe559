// Probe (GPU box): can a kernel tail-launch its own follow-up (CUDA dynamic parallelism, cudaStreamTailLaunch), also
// inside a captured CUDA graph?  And what do empty follow-up launches cost: ordinary vs programmatic dependent launch?
//   nvcc -rdc=true -gencode arch=compute_100a,code=sm_100a tools/cdp_probe.cu -lcudadevrt -o build/cdp_probe
#include <cstdio>
#include <cuda_runtime.h>

__global__ void round_kernel(int* state, int* log, int limit) {
    // state[0] = rounds done; the last thread of the grid decides whether another round is needed
    __shared__ int last;
    if (threadIdx.x == 0) last = (atomicAdd(&state[1], 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (last && threadIdx.x == 0) {
        state[1] = 0;
        int r = state[0];
        log[r] = r + 1;
        state[0] = r + 1;
        __threadfence();
        if (r + 1 < limit) round_kernel<<<gridDim.x, blockDim.x, 0, cudaStreamTailLaunch>>>(state, log, limit);
    }
}

__global__ void after_kernel(const int* state, int* out) { if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = state[0]; }

__global__ void empty_kernel(const int* state) {
    if (state[0] > 1000000) printf("never\n");
}
__global__ void empty_pdl_kernel(const int* state) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (state[0] > 1000000) printf("never\n");
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

int main() {
    int *state, *log, *out;
    CK(cudaMalloc(&state, 64)); CK(cudaMalloc(&log, 64 * 4)); CK(cudaMalloc(&out, 4));
    cudaStream_t s; CK(cudaStreamCreate(&s));
    int h[16], hout;
    // 1. eager tail-launch chain
    CK(cudaMemsetAsync(state, 0, 64, s)); CK(cudaMemsetAsync(log, 0, 256, s));
    round_kernel<<<256, 256, 0, s>>>(state, log, 6);
    after_kernel<<<1, 32, 0, s>>>(state, out);
    CK(cudaGetLastError()); CK(cudaStreamSynchronize(s));
    CK(cudaMemcpy(&hout, out, 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(h, log, 64, cudaMemcpyDeviceToHost));
    printf("eager: kernel after the chain saw %d rounds (expected 6); log %d %d %d %d %d %d\n", hout, h[0], h[1], h[2], h[3], h[4], h[5]);
    // 2. the same inside a captured graph, replayed three times
    cudaGraph_t g; cudaGraphExec_t ge;
    CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    CK(cudaMemsetAsync(state, 0, 64, s));
    round_kernel<<<256, 256, 0, s>>>(state, log, 5);
    after_kernel<<<1, 32, 0, s>>>(state, out);
    cudaError_t ce = cudaStreamEndCapture(s, &g);
    if (ce != cudaSuccess) { printf("graph: capture failed: %s\n", cudaGetErrorString(ce)); }
    else {
        ce = cudaGraphInstantiate(&ge, g, 0);
        if (ce != cudaSuccess) printf("graph: instantiate failed: %s\n", cudaGetErrorString(ce));
        else {
            for (int i = 0; i < 3; ++i) {
                ce = cudaGraphLaunch(ge, s);
                if (ce != cudaSuccess) { printf("graph: launch failed: %s\n", cudaGetErrorString(ce)); break; }
                CK(cudaStreamSynchronize(s));
                CK(cudaMemcpy(&hout, out, 4, cudaMemcpyDeviceToHost));
                printf("graph replay %d: kernel after the chain saw %d rounds (expected 5)\n", i, hout);
            }
        }
    }
    cudaGetLastError();
    // 3. cost of ten empty follow-up launches behind a 256-CTA kernel
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaMemset(state, 0, 64));
    float ms;
    for (int variant = 0; variant < 3; ++variant) {
        for (int rep = 0; rep < 2; ++rep) {
            CK(cudaEventRecord(e0, s));
            for (int it = 0; it < 100; ++it) {
                if (variant == 0) {
                    for (int r = 0; r < 10; ++r) empty_kernel<<<256, 256, 1536, s>>>(state);
                } else if (variant == 1) {
                    for (int r = 0; r < 10; ++r) {
                        cudaLaunchConfig_t cfg = {};
                        cfg.gridDim = dim3(256); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 1536; cfg.stream = s;
                        cudaLaunchAttribute at[1];
                        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                        at[0].val.programmaticStreamSerializationAllowed = 1;
                        cfg.attrs = at; cfg.numAttrs = 1;
                        CK(cudaLaunchKernelEx(&cfg, empty_pdl_kernel, (const int*)state));
                    }
                } else {
                    empty_kernel<<<256, 256, 1536, s>>>(state);
                }
            }
            CK(cudaEventRecord(e1, s)); CK(cudaStreamSynchronize(s));
            CK(cudaEventElapsedTime(&ms, e0, e1));
        }
        printf("%s: %.2f us per group\n", variant == 0 ? "10 ordinary empty launches" : variant == 1 ? "10 PDL empty launches" : "1 empty launch", ms * 10.0f);
    }
    return 0;
}
