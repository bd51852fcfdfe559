// K3, the HBM-bound kernel of the path: out[j][d] = sum_n gamma[n][j] * f[d][n] / npi[j] for J == 16 and features in
// their native (B,D,N) layout, as a persistent, warp-specialised TMA -> shared memory -> mma.sync pipeline.
//
//   producer warp (warp 8)                                  consumer warps 0..7 (WR row groups x WN point groups)
//   ----------------------                                  -----------------------------------------------------
//   wait empty[slot]                                        wait full[slot]
//   lane 0: expect_tx + WN x cp.async.bulk.tensor.3d        per 16 points: 4 x LDS.128 features (B fragments, conflict-
//           (box = 32 points x RB rows, SWIZZLE_128B)         free), 2 x LDS.128 gamma (A fragments, pre-arranged),
//   all lanes: the stage's gamma block, 4-byte cp.async       hi/lo split, 24 x mma.sync.m16n8k8.tf32
//           scattered into A-fragment order, arriving on    lane 0: arrive empty[slot]; partial sums += in FP32
//           full[slot] by itself (mbarrier.arrive.noinc)    per item: column sums -> npi, divide, store
//
//   * The feature matrix is the only large operand (B*D*N*4 bytes, read once).  TMA moves it without touching the
//     register file, so the bytes in flight per SM are set by the ring (2 CTAs x 3 slots x 32 KB), not by occupancy,
//     and nothing in the producer ever waits on memory.  Rows past D and points past N are zero-filled by the tensor
//     map's bounds (gamma: cp.async src-size 0), so ragged shapes need no special path.
//   * Work item = (cloud, block of RB = 32 WR rows) over all points.  The grid is persistent (<= 2 CTAs per SM) and
//     sized so that every CTA walks the same number of items; the producer runs ahead into the next item while the
//     consumer warps finish the previous one, so the ring never drains between items.
//   * The product is computed transposed, D[16 clusters x 8 rows] += A[16 x 8 points] * B[8 points x 8 rows]: the big
//     operand is then B, whose two k slots per lane are adjacent values of ONE feature row.  K is a summation index,
//     so the slots may hold any permutation of the stage's points as long as A and B agree: lane (g = lane / 4,
//     t = lane % 4) takes slots t, t + 4 of k-step 2h + k2 from points 16h + 4t + 2 k2 (+1) -- halves of the float4 it
//     loads with one LDS.128, no register shuffling.  Rows are an output index and free to permute as well: B column g
//     reads tile row pi(g) = (g >> 1) | ((g & 1) << 2), which puts the two rows of every quarter-warp in opposite
//     halves of the 128-byte swizzle atom -- no bank conflicts.  The gamma quad {[p][g], [p][g+8], [p+1][g], [p+1][g+8]}
//     of a k-step is one LDS.128 from the layout the producer scatters it into, shared by the warp's four row tiles.
//   * FP32 fidelity through the error-compensated TF32 split (the reference accumulates in FP32; budget 1e-4):
//     x = hi + lo, hi = x with the low 13 mantissa bits cleared (what the tensor core reads anyway), lo = x - hi
//     (exact); sum += lo*hi + hi*lo + hi*hi.  The tensor core's own FP32 accumulation truncates, so each stage's 12
//     MMAs per tile start from zero and are added to the running sums with FADD: 1.4e-6 max relative error vs FP64
//     on the bench shape (2.4e-5 when accumulating all 1024 points inside the MMA).
//   * 8 consumer warps = WR x WN: a stage is 8 sub-tiles of 32 rows x 32 points, WR = 8 (256 rows x 32 points) when
//     that yields enough items to fill the GPU, else 4 / 2 / 1 with the stage's points split across warps and the
//     partial sums meeting in shared memory once per item.
//
// Taken by ogmm_gmm_moments_feat for J == 16, N % 4 == 0, native layout, contiguous gamma, 16-byte aligned rows
// (OGMM_FEAT_NO_TMA=1 disables it); everything else runs the FP32 FFMA2 kernel in moments.cu.
#include <cuda.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace ogmm {

constexpr int kTmConsumers = 8;
constexpr int kTmThreads = 32 * (kTmConsumers + 1);
constexpr int kTmPts = 32;                        // points per stage
constexpr int kTmRowsWarp = 32;
constexpr int kTmGammaBytes = kTmPts * 16 * 4;    // 2 KB
constexpr int kTmStageTile = 8 * 32 * 128;        // 8 sub-tiles of 32 rows x 128 B
constexpr int kTmRedPitch = 17;
constexpr int kTmMaxSlots = 8;
constexpr int kTmSmemPerCta = 113 * 1024;         // two CTAs per SM: 2 x (113 KB + 1 KB reserved) <= 228 KB
// gamma column sums [8][16] + barriers + 1 KB alignment slack
constexpr int kTmFixedSmem = 8 * 16 * 4 + 2 * kTmMaxSlots * 8 + 1024;

__device__ __forceinline__ uint32_t tm_hi(float x) { return __float_as_uint(x) & 0xffffe000u; }
__device__ __forceinline__ uint32_t tm_lo(float x, uint32_t hi) { return __float_as_uint(x - __uint_as_float(hi)); }

__device__ __forceinline__ void tm_mma(float (&c)[4], const uint32_t (&a)[4], const uint32_t b0, const uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(32 * kTmConsumers) : "memory"); }

__global__ void __launch_bounds__(kTmThreads, 2)
gmm_moments_feat_tma_kernel(const __grid_constant__ CUtensorMap fmap, const float* __restrict__ gamma, int N, int D,
                            int WR, int nslots, int nrb, int nitems, float* __restrict__ pi_out, float* __restrict__ mu_out) {
    // stage = RB rows x PB points = 8 sub-tiles of 32 rows x 32 points (one TMA box per 32-point column of sub-tiles);
    // consumer warp (wr, wn) owns sub-tile (rows 32 wr.., points 32 wn..) of every stage.  Work item = (cloud, block of
    // RB rows) over all points; a CTA walks items blockIdx.x, + gridDim.x, ... and its ring never drains in between.
    const int WN = kTmConsumers / WR, RB = kTmRowsWarp * WR, PB = kTmPts * WN;
    extern __shared__ __align__(16) unsigned char tm_raw[];
    unsigned char* ring = tm_raw + ((1024u - (smem_u32(tm_raw) & 1023u)) & 1023u);
    const int box_bytes = RB * 128;                  // one box: RB rows x 32 points
    const int tile_bytes = kTmStageTile;             // WN boxes = 32 KB
    const int stage_bytes = tile_bytes + WN * kTmGammaBytes;
    float* s_red = reinterpret_cast<float*>(ring + (size_t)nslots * stage_bytes);   // [WN - 1][RB][17] parked partial sums
    float* s_gs = s_red + (WN - 1) * RB * kTmRedPitch;                               // [WN][16] gamma column sums
    uint64_t* s_full = reinterpret_cast<uint64_t*>(s_gs + kTmConsumers * 16);
    uint64_t* s_empty = s_full + kTmMaxSlots;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nstages = (N + PB - 1) / PB;

    if (tid == 0) {
        for (int s = 0; s < nslots; ++s) { mbar_init(s_full + s, 33); mbar_init(s_empty + s, kTmConsumers); }
        asm volatile("prefetch.tensormap [%0];" ::"l"(&fmap) : "memory");
    }
    __syncthreads();

    if (warp == kTmConsumers) {
        // ================================ producer ===================================================
        // Nothing here waits on memory: the feature boxes are TMA, gamma is scattered into its fragment layout by
        // 4-byte cp.async (zero-filled past N), and both complete on the slot's full barrier by themselves.
        int slot = 0, lap = 0;
        uint32_t phase = 1;                          // parity of the slot's previous release; the first lap does not wait
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            const int b = item / nrb, d0 = (item - b * nrb) * RB;
            const float* gam = gamma + (int64_t)b * N * 16;
            for (int st = 0; st < nstages; ++st) {
                if (lap) mbar_wait(s_empty + slot, phase);
                unsigned char* sl = ring + (size_t)slot * stage_bytes;
                if (lane == 0) {
                    mbar_expect_tx(s_full + slot, (uint32_t)tile_bytes);
                    for (int k = 0; k < WN; ++k)
                        tma_load_3d(smem_u32(sl + k * box_bytes), &fmap, st * PB + k * kTmPts, d0, b, s_full + slot);
                }
                for (int k = 0; k < WN; ++k) {
                    const int n = st * PB + k * kTmPts + lane;
                    const uint32_t ok = n < N ? 4u : 0u;
                    const float* src = gam + (int64_t)(n < N ? n : 0) * 16;
                    // point p = lane of the sub-block, cluster j -> quad (p / 2, j % 8) of the A-fragment layout, element
                    // 2 (p % 2) + j / 8; quads of a point pair are XOR-swizzled by 2 (pair / 2 % 4) against bank conflicts
                    const int pp = lane >> 1;
                    const uint32_t dst = smem_u32(sl + tile_bytes + k * kTmGammaBytes) + (uint32_t)(pp * 128 + (lane & 1) * 8);
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst + (uint32_t)((((j & 7) ^ (2 * ((pp >> 1) & 3))) << 4) + (j >> 3) * 4)),
                                     "l"(src + j), "r"(ok) : "memory");
                }
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(s_full + slot)) : "memory");
                if (lane == 0) mbar_arrive(s_full + slot);
                if (++slot == nslots) { slot = 0; phase ^= 1u; lap = 1; }
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
    } else {
        // ================================ consumers ==================================================
        // Transposed product: D[16 clusters x 8 rows] += A[16 x 8 points] * B[8 points x 8 rows] per m16n8k8, so that the
        // big operand is B: a lane's two k slots of a row are adjacent in the LDS.128 it already holds (no register
        // shuffling), and the gamma fragment (A) is shared by the warp's four row tiles.
        const int g = lane >> 2, t = lane & 3;
        const int pg = (g >> 1) | ((g & 1) << 2);    // tile row (mod 8) read as B column g
        const int wr = warp & (WR - 1), wn = warp / WR;
        // byte offsets inside a slot: B chunks for h = 0, 1 of row (32 wr + pg) of box wn; A quads of point pair 2 t
        const uint32_t f_off[2] = {(uint32_t)(wn * box_bytes + (kTmRowsWarp * wr + pg) * 128 + ((t ^ pg) << 4)),
                                   (uint32_t)(wn * box_bytes + (kTmRowsWarp * wr + pg) * 128 + (((4 + t) ^ pg) << 4))};
        const uint32_t g_off = (uint32_t)(tile_bytes + wn * kTmGammaBytes + ((2 * t) * 8 + (g ^ (2 * t))) * 16);
        int slot = 0;
        uint32_t phase = 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            const int b = item / nrb, d0 = (item - b * nrb) * RB;
            float acc[4][4];                         // [row tile nt][c0..c3]: clusters g, g + 8 x rows pi(2t), pi(2t + 1)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[nt][e] = 0.f;
            float gsum[2] = {0.f, 0.f};              // sum over this lane's points of gamma[.][g], gamma[.][g + 8]
            for (int st = 0; st < nstages; ++st) {
                mbar_wait(s_full + slot, phase);
                const unsigned char* sl = ring + (size_t)slot * stage_bytes;
                // the 12 MMAs per accumulator tile of a stage start from zero and are added to the running sums in FP32
                // round-to-nearest: the tensor core's own accumulation truncates (~2e-5 relative over 1024 points)
                float part[4][4];
#pragma unroll
                for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                    for (int e = 0; e < 4; ++e) part[nt][e] = 0.f;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float4 x[4];                     // features [row tile nt][row pg][points 16 h + 4 t .. + 3]
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) x[nt] = *reinterpret_cast<const float4*>(sl + f_off[h] + nt * 1024);
#pragma unroll
                    for (int k2 = 0; k2 < 2; ++k2) { // k-step 2 h + k2: slots t, t + 4 <- points p = 16 h + 4 t + 2 k2, p + 1
                        // A quad: gamma[p][g], gamma[p][g + 8], gamma[p + 1][g], gamma[p + 1][g + 8]
                        const float4 aq = *reinterpret_cast<const float4*>(sl + g_off + (8 * h + k2) * 128);
                        if (wr == 0) { gsum[0] += aq.x + aq.z; gsum[1] += aq.y + aq.w; }
                        const float av[4] = {aq.x, aq.y, aq.z, aq.w};
                        uint32_t ah[4], al[4], bh[4][2], bl[4][2];
#pragma unroll
                        for (int e = 0; e < 4; ++e) { ah[e] = tm_hi(av[e]); al[e] = tm_lo(av[e], ah[e]); }
#pragma unroll
                        for (int nt = 0; nt < 4; ++nt) {
                            const float b0 = k2 ? x[nt].z : x[nt].x, b1 = k2 ? x[nt].w : x[nt].y;
                            bh[nt][0] = tm_hi(b0); bl[nt][0] = tm_lo(b0, bh[nt][0]);
                            bh[nt][1] = tm_hi(b1); bl[nt][1] = tm_lo(b1, bh[nt][1]);
                        }
                        // term-major order: four independent accumulator tiles between two MMAs on the same one
#pragma unroll
                        for (int nt = 0; nt < 4; ++nt) tm_mma(part[nt], al, bh[nt][0], bh[nt][1]);
#pragma unroll
                        for (int nt = 0; nt < 4; ++nt) tm_mma(part[nt], ah, bl[nt][0], bl[nt][1]);
#pragma unroll
                        for (int nt = 0; nt < 4; ++nt) tm_mma(part[nt], ah, bh[nt][0], bh[nt][1]);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(s_empty + slot);    // the slot's values are in registers (consumed by the MMAs)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                    for (int e = 0; e < 4; ++e) acc[nt][e] += part[nt][e];
                if (++slot == nslots) { slot = 0; phase ^= 1u; }
            }
            // ---- item epilogue (consumer warps only; the producer is already filling the ring for the next item) ------
            if (wr == 0) {                           // column sums of this warp's points -> s_gs[wn][16]
#pragma unroll
                for (int jt = 0; jt < 2; ++jt) {
                    gsum[jt] += __shfl_xor_sync(kFull, gsum[jt], 1);
                    gsum[jt] += __shfl_xor_sync(kFull, gsum[jt], 2);
                }
                if (t == 0) { s_gs[wn * 16 + g] = gsum[0]; s_gs[wn * 16 + 8 + g] = gsum[1]; }
            }
            const int r0 = t, r1 = t | 4;            // pi(2t), pi(2t + 1): tile rows of C columns 2t, 2t + 1
            if (wn > 0) {                            // point groups 1.. park their partial sums: s_red[wn - 1][row][17]
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    const int row = (wn - 1) * RB + kTmRowsWarp * wr + 8 * nt;
                    s_red[(row + r0) * kTmRedPitch + g] = acc[nt][0];
                    s_red[(row + r1) * kTmRedPitch + g] = acc[nt][1];
                    s_red[(row + r0) * kTmRedPitch + g + 8] = acc[nt][2];
                    s_red[(row + r1) * kTmRedPitch + g + 8] = acc[nt][3];
                }
            }
            consumer_bar();
            if (wn == 0) {                           // point group 0 adds them up and writes its 32 rows x 16 clusters
                float npi[2];
#pragma unroll
                for (int jt = 0; jt < 2; ++jt) {     // npi[j] = N * pi[j] + 1e-5 in the reference's operation order
                    float tot = 0.f;
                    for (int w = 0; w < WN; ++w) tot += s_gs[w * 16 + 8 * jt + g];
                    const float pi = __fdiv_rn(tot, (float)N);
                    npi[jt] = __fadd_rn(__fmul_rn(pi, (float)N), 1e-5f);
                    if (d0 == 0 && wr == 0 && t == 0 && pi_out) pi_out[(int64_t)b * 16 + 8 * jt + g] = pi;
                }
#pragma unroll
                for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int row = kTmRowsWarp * wr + 8 * nt + ((e & 1) ? r1 : r0), j = g + 8 * (e >> 1);
                        float v = acc[nt][e];
                        for (int w = 1; w < WN; ++w) v += s_red[((w - 1) * RB + row) * kTmRedPitch + j];
                        if (d0 + row < D) mu_out[((int64_t)b * 16 + j) * D + d0 + row] = __fdiv_rn(v, npi[e >> 1]);
                    }
            }
            consumer_bar();                          // s_red / s_gs are free for the next item
        }
    }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup (the library does not link libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

}  // namespace ogmm

using namespace ogmm;

// Returns OGMM_EUNSUPPORTED (without touching the error string) when the call does not fit; the caller falls back.
int ogmm_launch_moments_feat_tma(const float* gamma, int64_t g_sb, int64_t g_sn, int64_t g_sj,
                                 const float* feats, int64_t f_sb, int64_t f_sn, int64_t f_sd,
                                 int64_t B, int64_t N, int64_t J, int64_t D, float* pi_out, float* mu_out, cudaStream_t s) {
    if (J != 16 || (N & 3) != 0 || N < 32 || f_sn != 1 || (f_sd & 3) != 0 || (f_sb & 3) != 0 || f_sd < N || g_sj != 1 ||
        g_sn != 16 || g_sb != N * 16 || (reinterpret_cast<uintptr_t>(feats) & 15) != 0 ||
        (reinterpret_cast<uintptr_t>(gamma) & 15) != 0 || D >= (1ll << 31))
        return OGMM_EUNSUPPORTED;
    if (B > 1 && f_sb < D * f_sd) return OGMM_EUNSUPPORTED;
    EncodeTiledFn encode = encode_tiled_fn();
    if (!encode) return OGMM_EUNSUPPORTED;

    // per device (one process may drive several GPUs from several threads, e.g. nn.DataParallel): the function attribute
    // belongs to the current device's context, so it is set on every call; the SM count is cached per device
    static int sm_count_by_dev[64] = {0};
    int dev = 0;
    int st = cuda_status(cudaGetDevice(&dev), "cudaGetDevice");
    if (st != OGMM_OK) return st;
    st = cuda_status(cudaFuncSetAttribute(gmm_moments_feat_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTmSmemPerCta),
                     "cudaFuncSetAttribute(gmm_moments_feat_tma_kernel)");
    if (st != OGMM_OK) return st;
    int sm_count = (dev >= 0 && dev < 64) ? sm_count_by_dev[dev] : 0;
    if (sm_count == 0) {
        st = cuda_status(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev), "cudaDeviceGetAttribute");
        if (st != OGMM_OK) return st;
        if (dev >= 0 && dev < 64) sm_count_by_dev[dev] = sm_count;
    }
    const int64_t resident = 2ll * sm_count;         // two CTAs per SM (shared memory and registers)

    // row block of an item: 256 rows (3 ring slots, no cross-warp sums) when that still gives every resident CTA an
    // item; otherwise halve it (and split the stage's points across warps instead) until it does
    int wr = (int)((D + kTmRowsWarp - 1) / kTmRowsWarp);
    wr = wr >= 8 ? 8 : (wr >= 4 ? 4 : (wr >= 2 ? 2 : 1));
    while (wr > 1 && B * ((D + wr * kTmRowsWarp - 1) / (wr * kTmRowsWarp)) < resident) wr >>= 1;
    const int wn = kTmConsumers / wr, rows_cta = wr * kTmRowsWarp;
    const int stage_bytes = kTmStageTile + wn * kTmGammaBytes;
    const int red_bytes = (wn - 1) * rows_cta * kTmRedPitch * 4;
    int nslots = (kTmSmemPerCta - kTmFixedSmem - red_bytes) / stage_bytes;      // 3 for 256-row tiles, 2 below
    nslots = nslots > kTmMaxSlots ? kTmMaxSlots : nslots;
    const size_t smem = (size_t)nslots * stage_bytes + red_bytes + kTmFixedSmem;
    CUtensorMap map;
    const cuuint64_t dims[3] = {(cuuint64_t)N, (cuuint64_t)D, (cuuint64_t)B};
    const cuuint64_t strides[2] = {(cuuint64_t)f_sd * 4, (cuuint64_t)(B > 1 ? f_sb : D * f_sd) * 4};
    const cuuint32_t box[3] = {(cuuint32_t)kTmPts, (cuuint32_t)rows_cta, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(feats), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return OGMM_EUNSUPPORTED;

    const int64_t nrb = (D + rows_cta - 1) / rows_cta, nitems = B * nrb;
    if (nitems >= (1ll << 31)) return OGMM_EUNSUPPORTED;
    // every CTA walks the same number of items (the last ones one fewer): with the ring depth fixed, equal work per CTA
    // means equal bandwidth per CTA, so they all finish together instead of leaving a half-empty last round
    const int64_t per_cta = (nitems + resident - 1) / resident;
    const unsigned grid = (unsigned)((nitems + per_cta - 1) / per_cta);
    gmm_moments_feat_tma_kernel<<<grid, kTmThreads, smem, s>>>(map, gamma, (int)N, (int)D, wr, nslots, (int)nrb, (int)nitems,
                                                               pi_out, mu_out);
    OGMM_LAUNCH_CHECK("gmm_moments_feat_tma_kernel");
    return OGMM_OK;
}
