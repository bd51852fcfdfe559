// K3 on the tensor cores: out[j][d] = sum_n gamma[n][j] * f[d][n] / npi[j] for J <= 16 and features in their native
// (B,D,N) layout -- the HBM-bound kernel of the path, rebuilt so that the SM does almost no per-element work.
//
//   D[128 d x 32] (TMEM, FP32) += A[128 d x 32 n] * B[32 x 32 n]^T       per 32-point K block, as 4 + 4 tcgen05 MMAs
//
//   * A = the feature tile exactly as it lies in HBM (row d = 32 consecutive points = 128 bytes): it is copied
//     global -> shared with 16-byte cp.async straight into the canonical K-major SWIZZLE_128B layout, three stages
//     deep -- no registers, no transposition, and enough bytes in flight (2 CTAs x 2 stages x 16 KB per SM) to cover
//     the HBM latency.
//   * FP32 fidelity through an error-compensated TF32 split (the reference accumulates in FP32; budget 1e-4):
//       f = f_hi + f_lo, gamma = g_hi + g_lo with *_hi the value truncated to TF32's 10 mantissa bits (written back
//       so the MMA sees exactly representable operands) and *_lo the exact FP32 remainder.
//       MMA 1: A = f_hi, B = [g_hi ; g_lo]  (N = 32: columns 0-15 take f_hi g_hi, 16-31 take f_hi g_lo)
//       MMA 2: A = f_lo, B = g_hi           (N = 16, accumulated onto columns 0-15)
//     The dropped f_lo g_lo term and the rounding of the *_lo operands are ~2^-21 relative; accumulation is FP32 in
//     TMEM.  The split costs one 128-bit shared load, 8 ALU ops and two 128-bit stores per four feature values.
//   * B: the 32 x 16 gamma block of the stage (2 KB, from L2) is split and transposed into a 32-row K-major
//     SWIZZLE_128B tile (rows 0-15 g_hi^T, 16-31 g_lo^T) by scalar shared stores -- 2 values per thread per stage.
//   * One elected thread issues the MMAs; a tcgen05.commit per stage arrives on an mbarrier that the CTA waits on, in
//     order, before it reuses the stage's ring slot and the single f_lo / B tiles.  After the last stage the 128 x 32 accumulator is read once from TMEM
//     (tcgen05.ld), the two column halves are added, divided by npi and written coalesced along d.
//
// Opt-in (OGMM_FEAT_TENSOR=1) when the layout allows (J == 16, D % 128 == 0, N % 4 == 0, contiguous gamma, 16-byte
// aligned rows).  Measured on B200 it loses to the register-operand kernels: the in-place hi/lo split moves every
// feature byte through shared memory four more times (LDS + 2 STS + the second MMA read), which, not HBM, bounds it
// (0.41 ms vs 0.365 ms for the FFMA2 kernel on the bench workload).  Kept as the tcgen05 reference for this shape.
#include "common.cuh"
#include "tc_common.cuh"

namespace ogmm {

constexpr int kTcThreads = 256;
constexpr int kTcStages = 4;                      // ring slots; loads run kTcLook stages ahead, MMAs may lag two stages
constexpr int kTcLook = 2;
constexpr int kTcRows = 128;                      // feature rows per CTA = MMA M
constexpr int kTcKB = 32;                         // points per stage = one 128-byte swizzle row
constexpr int kTcJ = 16;
constexpr int kTcTileA = kTcRows * 128;           // 16 KB
constexpr int kTcTileB = 32 * 128;                // 4 KB: rows 0-15 g_hi^T, rows 16-31 g_lo^T
constexpr int kTcRawG = kTcKB * kTcJ * 4;         // 2 KB raw gamma block
constexpr int kTcStageBytes = kTcTileA;                     // ring slot: feature tile (becomes f_hi in place)
constexpr int kTcRawSlots = kTcLook + 1;                    // raw gamma blocks: consumed by the split pass, not by the MMAs
constexpr size_t kTcSmem = (size_t)kTcStages * kTcStageBytes + 2 * (kTcTileA /*f_lo*/ + kTcTileB) +
                           (size_t)kTcRawSlots * kTcRawG + 1024 /*alignment*/ + 128;      // 113,792 B: two CTAs per SM

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

__global__ void __launch_bounds__(kTcThreads, 2)
gmm_moments_feat_tc_kernel(const float* __restrict__ gamma, const float* __restrict__ feats, int64_t f_sb, int64_t f_sd,
                           int N, int J, int D, float* __restrict__ pi_out, float* __restrict__ mu_out) {
    extern __shared__ __align__(16) unsigned char tc_raw[];
    unsigned char* base = tc_raw + ((1024u - (smem_u32(tc_raw) & 1023u)) & 1023u);
    // ring of kTcStages slots {feature tile, raw gamma block}; two f_lo tiles and two B tiles (even / odd stages);
    // two barriers (even / odd stages); TMEM address; npi
    unsigned char* s_alo0 = base + (size_t)kTcStages * kTcStageBytes;
    unsigned char* s_bt0 = s_alo0 + 2 * kTcTileA;
    unsigned char* s_raw0 = s_bt0 + 2 * kTcTileB;
    unsigned char* tail = s_raw0 + kTcRawSlots * kTcRawG;
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(tail);               // [2]
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 2);
    float* s_npi = reinterpret_cast<float*>(s_tmem + 1);               // [16]

    const int b = blockIdx.y, d0 = blockIdx.x * kTcRows;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float* f = feats + (int64_t)b * f_sb + (int64_t)d0 * f_sd;
    const float* g = gamma + (int64_t)b * N * kTcJ;

    if (tid == 0) { mbar_init(s_bar, 1); mbar_init(s_bar + 1, 1); }
    if (warp == 0) tmem_alloc(s_tmem, 32);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;
    const uint32_t idesc32 = make_idesc_tf32(kTcRows, 32), idesc16 = make_idesc_tf32(kTcRows, 16);

    const int n_stages = (N + kTcKB - 1) / kTcKB;
    // stage loader: feature rows via cp.async into the swizzled A_hi tile, raw gamma block behind it
    auto issue_stage = [&](int s) {
        const int slot = s % kTcStages, n0 = s * kTcKB;
        unsigned char* st = base + (size_t)slot * kTcStageBytes;
        const uint32_t a_hi = smem_u32(st);
        const bool full = n0 + kTcKB <= N;
#pragma unroll
        for (int k = 0; k < (kTcRows * 8) / kTcThreads; ++k) {
            const int e = tid + kTcThreads * k, r = e >> 3, ch = e & 7;
            if (full || n0 + 4 * ch + 3 < N) cp_async16(a_hi + sw128_off(r, ch), f + (int64_t)r * f_sd + n0 + 4 * ch);
            else *reinterpret_cast<float4*>(st + sw128_off(r, ch)) = make_float4(0.f, 0.f, 0.f, 0.f);   // N % 4 == 0: whole chunk out
        }
        if (tid < kTcKB * kTcJ / 4) {
            const int n = tid >> 2;                                     // 4 chunks of 16 B per point
            unsigned char* rw = s_raw0 + (s % kTcRawSlots) * kTcRawG;
            if (n0 + n < N) cp_async16(smem_u32(rw) + tid * 16, g + (int64_t)n0 * kTcJ + tid * 4);
            else *reinterpret_cast<float4*>(rw + tid * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        cp_async_commit();
    };

    float gsum = 0.f;                                                   // threads 0..15: column sum of gamma[:, tid]
    for (int s = 0; s < kTcLook && s < n_stages; ++s) issue_stage(s);
    for (int s = n_stages; s < kTcLook; ++s) cp_async_commit();          // keep the group count uniform

    uint32_t parity = 0;                                                // bit x: parity to wait for on barrier x
    for (int s = 0; s < n_stages; ++s) {
        const int slot = s % kTcStages, eo = s & 1;
        unsigned char* st = base + (size_t)slot * kTcStageBytes;
        unsigned char* s_alo = s_alo0 + eo * kTcTileA;
        unsigned char* s_bt = s_bt0 + eo * kTcTileB;
        // The MMAs of stage s-2 must be done before (a) the load issued below overwrites their ring slot and (b) the
        // split pass rewrites the f_lo / B tiles they read.  They were issued a whole stage ago: no exposed latency.
        if (s >= 2) { mbar_wait(s_bar + eo, (parity >> eo) & 1u); parity ^= 1u << eo; }
        if (s + kTcLook < n_stages) issue_stage(s + kTcLook);
        else cp_async_commit();
        cp_async_wait<kTcLook>();                                       // this thread's copies of stage s have landed
        __syncthreads();                                                // ... and everybody else's

        // ---- split pass: A_hi <- trunc(f), A_lo <- f - trunc(f); B <- [g_hi^T ; g_lo^T] --------------------------
#pragma unroll
        for (int k = 0; k < (kTcRows * 8) / kTcThreads; ++k) {
            const int e = tid + kTcThreads * k, r = e >> 3, ch = e & 7;
            float4* ph = reinterpret_cast<float4*>(st + sw128_off(r, ch));
            const float4 v = *ph;
            const float4 h = make_float4(tf32_trunc(v.x), tf32_trunc(v.y), tf32_trunc(v.z), tf32_trunc(v.w));
            *ph = h;
            *reinterpret_cast<float4*>(s_alo + sw128_off(r, ch)) = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
        }
        {
            const float* raw = reinterpret_cast<const float*>(s_raw0 + (s % kTcRawSlots) * kTcRawG);
            unsigned char* bt = s_bt;
#pragma unroll
            for (int k = 0; k < (kTcKB * kTcJ) / kTcThreads; ++k) {
                const int e = tid + kTcThreads * k, n = e >> 4, j = e & 15;        // consecutive threads: consecutive j
                const float v = raw[e];
                const float h = tf32_trunc(v);
                *reinterpret_cast<float*>(bt + sw128_off(j, n >> 2) + (n & 3) * 4) = h;
                *reinterpret_cast<float*>(bt + sw128_off(16 + j, n >> 2) + (n & 3) * 4) = v - h;
            }
            if (tid < kTcJ) {
#pragma unroll 8
                for (int n = 0; n < kTcKB; ++n) gsum += raw[n * kTcJ + tid];
            }
        }
        proxy_fence_async();
        __syncthreads();

        // ---- MMAs of this stage (one thread) -----------------------------------------------------------------------------
        if (tid == 0) {
            tc_fence_after();
            const uint32_t a_hi = smem_u32(st), a_lo = smem_u32(s_alo), bq = smem_u32(s_bt);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                umma_tf32(tmem_base, make_desc_sw128(a_hi + kk * 32), make_desc_sw128(bq + kk * 32), idesc32, (s | kk) ? 1u : 0u);
                umma_tf32(tmem_base, make_desc_sw128(a_lo + kk * 32), make_desc_sw128(bq + kk * 32), idesc16, 1u);
            }
            umma_commit(s_bar + eo);
        }
    }
    // ---- npi from the gamma column sums -------------------------------------------------------------------------------------
    if (tid < kTcJ) {
        const float pi = __fdiv_rn(gsum, (float)N);
        s_npi[tid] = __fadd_rn(__fmul_rn(pi, (float)N), 1e-5f);
        if (blockIdx.x == 0 && tid < J && pi_out) pi_out[(int64_t)b * J + tid] = pi;
    }
    // ---- wait for the last commit (it covers every earlier MMA), read the accumulators ----------------------------------
    // the last commit covers every earlier MMA
    {
        const int eo = (n_stages - 1) & 1;
        if (n_stages >= 2 && ((n_stages - 2) & 1) != eo) { /* other barrier: its last completion is not needed */ }
        mbar_wait(s_bar + eo, (parity >> eo) & 1u);
    }
    tc_fence_after();
    __syncthreads();
    if (warp < 4) {
        float acc[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16), acc);
        const int d = d0 + warp * 32 + lane;
        if (d < D) {
#pragma unroll
            for (int j = 0; j < kTcJ; ++j)
                if (j < J) mu_out[((int64_t)b * J + j) * D + d] = __fdiv_rn(acc[j] + acc[16 + j], s_npi[j]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_free(tmem_base, 32);
}

}  // namespace ogmm

using namespace ogmm;

// Returns OGMM_OK after launching, or OGMM_EUNSUPPORTED when the layout does not fit this kernel (caller falls back to
// the FP32 kernel).
int ogmm_launch_moments_feat_tc(const float* gamma, int64_t g_sb, int64_t g_sn, int64_t g_sj,
                                const float* feats, int64_t f_sb, int64_t f_sn, int64_t f_sd,
                                int64_t B, int64_t N, int64_t J, int64_t D, float* pi_out, float* mu_out, cudaStream_t s) {
    const bool ok = J == kTcJ && (D % kTcRows) == 0 && (N % 4) == 0 && N >= kTcKB && f_sn == 1 && (f_sd % 4) == 0 &&
                    (f_sb % 4) == 0 && g_sj == 1 && g_sn == J && g_sb == N * J &&
                    (reinterpret_cast<uintptr_t>(feats) & 15) == 0 && (reinterpret_cast<uintptr_t>(gamma) & 15) == 0;
    if (!ok) return OGMM_EUNSUPPORTED;
    int st = cuda_status(cudaFuncSetAttribute(gmm_moments_feat_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)kTcSmem), "cudaFuncSetAttribute(gmm_moments_feat_tc_kernel)");
    if (st != OGMM_OK) return st;
    dim3 grid((unsigned)(D / kTcRows), (unsigned)B);
    gmm_moments_feat_tc_kernel<<<grid, kTcThreads, kTcSmem, s>>>(gamma, feats, f_sb, f_sd, (int)N, (int)J, (int)D, pi_out, mu_out);
    OGMM_LAUNCH_CHECK("gmm_moments_feat_tc_kernel");
    return OGMM_OK;
}
