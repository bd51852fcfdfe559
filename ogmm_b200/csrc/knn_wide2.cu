// K1w, pipelined: kNN in feature space (32 <= C <= 128) as a warp-specialised TMA -> tcgen05 -> TMEM pipeline (sm_100a).
//
// Same contract and the same FINAL arithmetic as knn_wide.cu / knn_generic_kernel (lib/utils.py:12-44: expanded-form
// distances, clamp 1e-12 or the cosine form, k smallest ascending, ties to the lowest index; the returned distances and
// their order are exact FP32).  What changes is how the O(N M C) part is organised:
//
//   * ONE Gram pass.  A small pre-kernel writes augmented copies of the operands, queries as [-2 x, 1, 1, 0..] and
//     candidates as [y, hi(|y|^2), lo(|y|^2), 0..] (8 extra columns), so the tensor core's accumulator IS the
//     approximate distance minus |q|^2,  v' = -2 x.y + |y|^2 : no per-element arithmetic is left for the CUDA cores but
//     the compare.  (hi/lo: the TF32 datapath drops 13 mantissa bits of an operand; the norm is split so that it
//     survives to ~2^-20.)
//   * Roles.  Warp 0 (one lane) streams candidate tiles with TMA (cp.async.bulk.tensor, SWIZZLE_128B boxes of 32
//     floats x BN rows, zero fill past M and past C + 8) into a 2-stage ring; warp 1 (one lane) issues
//     tcgen05.mma.kind::tf32 (M = 128 queries, N = BN) into one of TWO TMEM accumulator buffers and commits to
//     mbarriers; warps 4..11 are the selectors: two threads per query (TMEM lane = query; warps w and w + 4 share a
//     lane quadrant and split the tile's columns) read their 32 accumulator columns with one tcgen05.ld.32x32b.x32 and
//     free the buffer at once, so the MMAs of tile t + 1 run under the selection of tile t.
//   * Selection, one pass.  Per selector thread: a sorted list of its K smallest APPROXIMATE values (registers,
//     values only) and a shared-memory column of collected (value, index) pairs.  A value at or below
//     bound = K-th smallest so far + 2 E  (E = TF32 error bound of v', per query) is parked in a staging column; when a
//     lane's staging runs low the warp merges converged: values go into the sorted lists, pairs into the collected
//     columns, and a column that fills up is compacted against the current (tighter) bound.  Every true neighbour has
//     approximate value <= K-th smallest approximate + 2 E at ANY time, hence is never dropped.  At the end the
//     column is compacted once more and its few entries (K + a handful) are re-ranked with exact FP32 distances.
//   The result is bit-identical to the FP32 kernels (tests assert torch.equal).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace ogmm {

constexpr int kW2Q = 128;               // queries per CTA = MMA M = TMEM lanes
// threads: warp 0 TMA, warp 1 MMA, warps 2-3 idle, then SEL selector threads per query (SEL = 2: warps 4..11, the two
// threads of a query split every tile's columns; SEL = 1 when the operand tiles leave no room for 256 collect columns)
constexpr int kW2Aug = 8;               // extra K columns of the augmented operands
constexpr int kW2Stage = 12;            // staging slots per selector thread
constexpr int kW2Trigger = 4;           // merge when a lane holds more than this many (room for one more group of 8)
constexpr int kW2Cap = 48;              // collected pairs per selector thread

typedef unsigned long long u64x;
__device__ __forceinline__ unsigned w2_dist_bits(float d) {
    const unsigned b = __float_as_uint(d + 0.0f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float w2_bits_dist(unsigned k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }
constexpr u64x kW2Empty = (0xff800000ull << 32) | 0xffffffffull;

__device__ __forceinline__ void w2_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// wait of the single-lane roles: back off between polls so the two spinning lanes leave the issue slots to the selectors
__device__ __forceinline__ void w2_wait_idle(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) break;
        __nanosleep(512);
    }
}
__device__ __forceinline__ void w2_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void w2_tma_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// ---- pre-kernel: augmented operands --------------------------------------------------------------------------------------
// aug_q[b][n] = [-2 x (C), 1, 1, 0 x6]; aug_c[b][m] = [y (C), hi(|y|^2), lo(|y|^2), 0 x6] (normalize: norm columns 0);
// cn_max[b] = max_m |y|^2 (error bound).  One warp per row.
__global__ void __launch_bounds__(256)
knn_wide_augment_kernel(const float* __restrict__ src, int64_t s_sb, int64_t s_sn, int64_t s_sc,
                        const float* __restrict__ dst, int64_t d_sb, int64_t d_sn, int64_t d_sc,
                        int N, int M, int C, int normalize, float* __restrict__ aug_q, float* __restrict__ aug_c,
                        int* __restrict__ cn_max) {
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int CA = C + kW2Aug;
    if (row < N) {
        const float* x = src + (int64_t)b * s_sb + (int64_t)row * s_sn;
        float* o = aug_q + ((int64_t)b * N + row) * CA;
        for (int c = lane; c < CA; c += 32) o[c] = c < C ? -2.f * x[(int64_t)c * s_sc] : (c < C + 2 ? 1.f : 0.f);
    }
    if (row < M) {
        const float* y = dst + (int64_t)b * d_sb + (int64_t)row * d_sn;
        float* o = aug_c + ((int64_t)b * M + row) * CA;
        float acc = 0.f;
        for (int c = lane; c < C; c += 32) { const float v = y[(int64_t)c * d_sc]; o[c] = v; acc = fmaf(v, v, acc); }
        acc = warp_sum(acc);
        if (normalize) acc = 0.f;
        const float hi = __uint_as_float(__float_as_uint(acc) & 0xffffe000u);
        if (lane < kW2Aug) o[C + lane] = lane == 0 ? hi : (lane == 1 ? acc - hi : 0.f);
        if (lane == 0) atomicMax(cn_max + b, __float_as_int(acc));            // acc >= 0: integer order == float order
    }
}

// exact FP32 distance, the generic kernel's arithmetic
__device__ __forceinline__ float w2_exact_dist(const float* __restrict__ x, int64_t x_sc, const float* __restrict__ y, int64_t y_sc,
                                               int C, float qn, int normalize) {
    float acc = 0.f, cn = 0.f;
    for (int c = 0; c < C; ++c) {
        const float yv = y[(int64_t)c * y_sc];
        acc = fmaf(x[(int64_t)c * x_sc], -2.f * yv, acc);
        cn = __fadd_rn(cn, __fmul_rn(yv, yv));
    }
    if (normalize) return __fadd_rn(acc, 2.0f);
    return fmaxf(__fadd_rn(__fadd_rn(acc, qn), cn), 1e-12f);
}

// The same value from the staged query tile and 16-byte candidate loads: the tile holds -2 x (exact scaling), and
// fma(-2 x, y, acc) == fma(x, -2 y, acc) bit for bit; |y|^2 is summed in the same order as above.  Eight loads are in
// flight per batch instead of one dependent global load per feature.
__device__ __noinline__ float w2_exact_dist_fast(const unsigned char* __restrict__ sA, int ql, const float* __restrict__ y,
                                                    int C, float qn, int normalize) {
    float acc = 0.f, cn = 0.f;
    const float4* y4 = reinterpret_cast<const float4*>(y);
    for (int c0 = 0; c0 < C; c0 += 32) {                        // one 128-byte K-block of the tile per batch
        float4 yv[8];
        const int nch = min(8, (C - c0) >> 2);
#pragma unroll
        for (int u = 0; u < 8; ++u) if (u < nch) yv[u] = __ldg(y4 + (c0 >> 2) + u);
        const unsigned char* blk = sA + (size_t)(c0 >> 5) * kW2Q * 128;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (u < nch) {
                const float4 xa = *reinterpret_cast<const float4*>(blk + sw128_off(ql, u));
                acc = fmaf(xa.x, yv[u].x, acc); cn = __fadd_rn(cn, __fmul_rn(yv[u].x, yv[u].x));
                acc = fmaf(xa.y, yv[u].y, acc); cn = __fadd_rn(cn, __fmul_rn(yv[u].y, yv[u].y));
                acc = fmaf(xa.z, yv[u].z, acc); cn = __fadd_rn(cn, __fmul_rn(yv[u].z, yv[u].z));
                acc = fmaf(xa.w, yv[u].w, acc); cn = __fadd_rn(cn, __fmul_rn(yv[u].w, yv[u].w));
            }
        }
    }
    if (normalize) return __fadd_rn(acc, 2.0f);
    return fmaxf(__fadd_rn(__fadd_rn(acc, qn), cn), 1e-12f);
}

template <int K>
__device__ __forceinline__ void w2_key_insert(u64x (&key)[K], u64x kv) {
    if (!(kv < key[K - 1])) return;
    bool moved = false;
#pragma unroll
    for (int j = 0; j < K; ++j) {
        moved = moved || (kv < key[j]);
        const u64x t = key[j];
        key[j] = moved ? kv : t;
        kv = moved ? t : kv;
    }
}

struct Wide2Args {
    const float* src; int64_t s_sb, s_sn, s_sc;
    const float* dst; int64_t d_sb, d_sn, d_sc;
    int N, M, C, k, normalize, BN, KB;
    const int* cn_max;
    int64_t* idx_out; float* dist_out; int32_t* stats;
};

template <int K, int SEL>
__global__ void __launch_bounds__(128 + 128 * SEL, 1)
knn_wide2_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_c, Wide2Args a) {
    extern __shared__ __align__(16) unsigned char w2_raw[];
    unsigned char* sm = w2_raw + ((1024u - (smem_u32(w2_raw) & 1023u)) & 1023u);      // SWIZZLE_128B atoms: 1024-byte aligned
    constexpr int kW2Sel = 128 * SEL;
    const int BN = a.BN, KB = a.KB;
    const uint32_t a_bytes = (uint32_t)KB * kW2Q * 128, b_bytes = (uint32_t)KB * BN * 128;
    unsigned char* sA = sm;
    unsigned char* sB = sA + a_bytes;                                                  // [2][KB][BN x 128 B]
    float* s_stage_v = reinterpret_cast<float*>(sB + 2 * b_bytes);                     // [kW2Stage][SEL x 128] values ...
    int* s_stage_i = reinterpret_cast<int*>(s_stage_v + kW2Stage * kW2Sel);            // ... and indices, same stride
    float* s_col_v = reinterpret_cast<float*>(s_stage_i + kW2Stage * kW2Sel);          // [kW2Cap][SEL x 128]
    unsigned short* s_col_i = reinterpret_cast<unsigned short*>(s_col_v + kW2Cap * kW2Sel);
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_col_i + kW2Cap * kW2Sel);          // full_a, full_b[2], empty_b[2], tfull[2], tempty[2]
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 9);
    uint64_t* full_a = s_bar, *full_b = s_bar + 1, *empty_b = s_bar + 3, *tfull = s_bar + 5, *tempty = s_bar + 7;

    const int b = blockIdx.y, q0 = blockIdx.x * kW2Q, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_tiles = (a.M + BN - 1) / BN;

    if (tid == 0) {
        mbar_init(full_a, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(full_b + i, 1); mbar_init(empty_b + i, 1); mbar_init(tfull + i, 1); mbar_init(tempty + i, 4 * SEL); }
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_c) : "memory");
    }
    if (warp == 1) tmem_alloc(s_tmem, 2 * BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            w2_expect_tx(full_a, a_bytes);
            for (int kb = 0; kb < KB; ++kb) w2_tma_3d(smem_u32(sA) + kb * kW2Q * 128, &map_q, kb * 32, q0, b, full_a);
            for (int t = 0; t < n_tiles; ++t) {
                const int s = t & 1;
                w2_wait_idle(empty_b + s, ((t >> 1) & 1) ^ 1);              // first use of a slot passes at once
                w2_expect_tx(full_b + s, b_bytes);
                for (int kb = 0; kb < KB; ++kb)
                    w2_tma_3d(smem_u32(sB) + s * b_bytes + kb * BN * 128, &map_c, kb * 32, t * BN, b, full_b + s);
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(kW2Q, BN);
            const int ksteps = (a.C + kW2Aug + 7) / 8;                     // 8 TF32 columns per instruction
            w2_wait_idle(full_a, 0);
            for (int t = 0; t < n_tiles; ++t) {
                const int s = t & 1;
                w2_wait_idle(full_b + s, (t >> 1) & 1);
                w2_wait_idle(tempty + s, ((t >> 1) & 1) ^ 1);               // selectors are done with this TMEM buffer
                tc_fence_after();
                for (int ks = 0; ks < ksteps; ++ks) {
                    const int kb = ks >> 2, kk = ks & 3;
                    const uint64_t da = make_desc_sw128(smem_u32(sA) + kb * kW2Q * 128 + kk * 32);
                    const uint64_t db = make_desc_sw128(smem_u32(sB) + s * b_bytes + kb * BN * 128 + kk * 32);
                    umma_tf32(tmem_base + s * BN, da, db, idesc, ks ? 1u : 0u);
                }
                umma_commit(empty_b + s);                                   // the ring slot may be refilled
                umma_commit(tfull + s);                                     // the accumulators are complete
            }
        }
    } else if (warp >= 4) {
        // ================= selectors =================
        const int st = tid - 128;                       // 0..255: selector thread
        const int ql = st & (kW2Q - 1), half = st >> 7;
        const int q = q0 + ql;
        const bool valid = q < a.N;
        const float* sb = a.src + (int64_t)b * a.s_sb;
        const float* db = a.dst + (int64_t)b * a.d_sb;
        const float* x = sb + (int64_t)(valid ? q : 0) * a.s_sn;
        const bool y_vec = (a.d_sc == 1) && ((a.d_sn & 3) == 0) && ((a.d_sb & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.dst) & 15) == 0);
        float qn = 0.f;
        if (!a.normalize)
            for (int c = 0; c < a.C; ++c) { const float v = x[(int64_t)c * a.s_sc]; qn = __fadd_rn(qn, __fmul_rn(v, v)); }
        else qn = 2.0f;
        // error bound of v' (knn_wide.cu): 2 |dot_tf32 - dot| <= 2^-8 |x||y|, plus fp32 / norm-split slack
        const float cn_max = __int_as_float(a.cn_max[b]);
        const float xn = a.normalize ? 1.0f : qn, yn = a.normalize ? 1.0f : cn_max;
        const float err2 = 2.f * (0.00390625f * sqrtf(xn * yn) + 4e-6f * (xn + yn) + 2e-12f);

        float best[K];                                  // K smallest approximate values so far, ascending
#pragma unroll
        for (int j = 0; j < K; ++j) best[j] = INFINITY;
        float bound = valid ? 3.0e38f : -INFINITY;      // collect bound: best[K-1] + 2E (finite so that padding +inf never passes)
        int n_st = 0, n_col = 0;
        const uint32_t tmem_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        const int cols = BN / SEL, c_lo = half * cols;  // this thread's columns of every tile (a multiple of 32)

        u64x key[K];                                    // exact (distance, index) list: filled at the end, or earlier on overflow
#pragma unroll
        for (int j = 0; j < K; ++j) key[j] = kW2Empty;
        int overflow = 0;
        float* sp = s_stage_v + st;                     // next free staging slot of this thread (index slot: + kW2Stage * kW2Sel)
        auto compact = [&]() {                          // keep the collected pairs at or below the current bound
            int w = 0;
            for (int e = 0; e < n_col; ++e) {
                const float v = s_col_v[e * kW2Sel + st];
                if (v <= bound) {
                    const unsigned short ix = s_col_i[e * kW2Sel + st];
                    s_col_v[w * kW2Sel + st] = v; s_col_i[w * kW2Sel + st] = ix; ++w;
                }
            }
            n_col = w;
        };
        auto drain_exact = [&]() {                      // exact re-rank of everything collected (per thread, divergent)
            for (int e = 0; e < n_col; ++e) {
                if (s_col_v[e * kW2Sel + st] > bound) continue;
                const int m = s_col_i[e * kW2Sel + st];
                const float d = y_vec ? w2_exact_dist_fast(sA, ql, db + (int64_t)m * a.d_sn, a.C, qn, a.normalize)
                                      : w2_exact_dist(x, a.s_sc, db + (int64_t)m * a.d_sn, a.d_sc, a.C, qn, a.normalize);
                w2_key_insert<K>(key, ((u64x)w2_dist_bits(d) << 32) | (unsigned)m);
            }
            n_col = 0;
        };
        auto merge = [&]() {                            // warp-converged
            // room for every staged pair first: compact against the current bound; a column that is still too full
            // (masses of near-ties) is emptied into the exact list, so nothing is ever dropped
            if (__any_sync(kFull, n_col + n_st > kW2Cap)) {
                compact();
                if (n_col + n_st > kW2Cap) { drain_exact(); overflow = 1; }
            }
            const int most = __reduce_max_sync(kFull, n_st);
            for (int s2 = 0; s2 < most; ++s2) {
                float v = INFINITY;
                unsigned short ix = 0;
                if (s2 < n_st) { v = s_stage_v[s2 * kW2Sel + st]; ix = (unsigned short)s_stage_i[s2 * kW2Sel + st]; }
                if (__any_sync(kFull, v < best[K - 1])) {
                    float w = v;
#pragma unroll
                    for (int j = 0; j < K; ++j) { const float lo = fminf(w, best[j]); w = fmaxf(w, best[j]); best[j] = lo; }
                }
                if (v <= bound) { s_col_v[n_col * kW2Sel + st] = v; s_col_i[n_col * kW2Sel + st] = ix; ++n_col; }
            }
            n_st = 0;
            sp = s_stage_v + st;
            if (valid) bound = fminf(best[K - 1] + err2, 3.0e38f);
        };

        for (int t = 0; t < n_tiles; ++t) {
            const int s = t & 1;
            mbar_wait(tfull + s, (t >> 1) & 1);
            tc_fence_after();
            const int m0 = t * BN + c_lo;
            for (int c0 = 0; c0 < cols; c0 += 32) {
                float v[32];
                tmem_ld32(tmem_row + s * BN + c_lo + c0, v);
                if (c0 + 32 >= cols) {                  // last read of this buffer by this warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) w2_arrive(tempty + s);
                }
                const int mb = m0 + c0;
                if (mb + 32 > a.M) {                     // last, ragged tile: columns past M are TMA zero fill, not candidates
#pragma unroll
                    for (int i = 0; i < 32; ++i) if (mb + i >= a.M) v[i] = INFINITY;
                }
#pragma unroll
                for (int g = 0; g < 4; ++g) {
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int i = 8 * g + u;
                        if (v[i] <= bound) {
                            sp[0] = v[i];
                            reinterpret_cast<int*>(sp)[kW2Stage * kW2Sel] = mb + i;
                            sp += kW2Sel;
                            ++n_st;
                        }
                    }
                    if (__any_sync(kFull, n_st > kW2Trigger)) merge();
                }
            }
        }
        merge();
        drain_exact();
        if (overflow && a.stats && lane == 0) atomicAdd(a.stats, 1);

        // the second selector of each query hands its exact list to the first one (the staging / collect columns are dead)
        if constexpr (SEL == 2) {
            asm volatile("bar.sync 1, 256;" ::: "memory");
            u64x* s_keys = reinterpret_cast<u64x*>(s_col_v);                // [K][128] u64 <= the collect columns
            if (half == 1) {
#pragma unroll
                for (int j = 0; j < K; ++j) s_keys[j * kW2Q + ql] = key[j];
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (valid && half == 0) {
#pragma unroll 1
                for (int j = 0; j < K; ++j) w2_key_insert<K>(key, s_keys[j * kW2Q + ql]);
            }
        }
        if (valid && half == 0) {
            int64_t* io = a.idx_out + ((int64_t)b * a.N + q) * a.k;
            float* dout = a.dist_out ? a.dist_out + ((int64_t)b * a.N + q) * a.k : nullptr;
#pragma unroll
            for (int j = 0; j < K; ++j) {
                if (j < a.k) {
                    io[j] = (int)(unsigned)(key[j] & 0xffffffffull);
                    if (dout) dout[j] = w2_bits_dist((unsigned)(key[j] >> 32));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_free(tmem_base, 2 * BN);
}

typedef CUresult (*W2EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static W2EncodeFn w2_encode_fn() {
    static W2EncodeFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<W2EncodeFn>(p);
    }();
    return fn;
}

static size_t wide2_smem(int KB, int BN, int sel) {
    return (size_t)KB * kW2Q * 128 + 2 * (size_t)KB * BN * 128 + (size_t)(8 * kW2Stage + 6 * kW2Cap) * 128 * sel + 9 * 8 + 16 + 1024;
}

}  // namespace ogmm

using namespace ogmm;

// Pipelined tensor-core path; returns OGMM_EUNSUPPORTED (error string untouched) when the call does not fit, and the
// caller falls back to knn_wide.cu.  Scratch for the augmented operands comes from the stream-ordered allocator.
int ogmm_launch_knn_wide2(const float* src, int64_t s_sb, int64_t s_sn, int64_t s_sc,
                          const float* dst, int64_t d_sb, int64_t d_sn, int64_t d_sc,
                          int64_t B, int64_t N, int64_t M, int64_t C, int64_t k, int normalize,
                          int64_t* idx_out, float* dist_out, int32_t* stats, cudaStream_t s) {
    if (C < 32 || C > 128 || (C & 3) != 0 || k > 32 || M > 65535 || N > (1 << 30) || B > 65535 || M < 64) return OGMM_EUNSUPPORTED;
    W2EncodeFn encode = w2_encode_fn();
    if (!encode) return OGMM_EUNSUPPORTED;
    const int CA = (int)C + kW2Aug, KB = (CA + 31) / 32;
    const int BN = 64;
    const size_t limit = 226 * 1024;
    const int sel = wide2_smem(KB, BN, 2) <= limit ? 2 : 1;
    if (wide2_smem(KB, BN, sel) > limit) return OGMM_EUNSUPPORTED;
    const size_t smem = wide2_smem(KB, BN, sel);

    float *aug_q = nullptr, *aug_c = nullptr;
    int* cn_max = nullptr;
    const size_t q_bytes = (size_t)B * N * CA * 4, c_bytes = (size_t)B * M * CA * 4;
    int st = cuda_status(cudaMallocAsync(reinterpret_cast<void**>(&aug_q), q_bytes + c_bytes + 4 * (size_t)B + 256, s), "cudaMallocAsync(knn_wide scratch)");
    if (st != OGMM_OK) return st;
    aug_c = aug_q + (size_t)B * N * CA;
    cn_max = reinterpret_cast<int*>(aug_c + (size_t)B * M * CA);
    auto fail = [&](int code) { cudaFreeAsync(aug_q, s); return code; };
    st = cuda_status(cudaMemsetAsync(cn_max, 0, 4 * (size_t)B, s), "cudaMemsetAsync(cn_max)");
    if (st != OGMM_OK) return fail(st);
    {
        const int64_t rows = N > M ? N : M;
        dim3 grid((unsigned)((rows + 7) / 8), (unsigned)B);
        knn_wide_augment_kernel<<<grid, 256, 0, s>>>(src, s_sb, s_sn, s_sc, dst, d_sb, d_sn, d_sc, (int)N, (int)M, (int)C, normalize,
                                                     aug_q, aug_c, cn_max);
        st = cuda_status(cudaGetLastError(), "knn_wide_augment_kernel");
        if (st != OGMM_OK) return fail(st);
    }
    CUtensorMap map_q, map_c;
    {
        const cuuint64_t dq[3] = {(cuuint64_t)CA, (cuuint64_t)N, (cuuint64_t)B}, dc[3] = {(cuuint64_t)CA, (cuuint64_t)M, (cuuint64_t)B};
        const cuuint64_t sq[2] = {(cuuint64_t)CA * 4, (cuuint64_t)N * CA * 4}, sc[2] = {(cuuint64_t)CA * 4, (cuuint64_t)M * CA * 4};
        const cuuint32_t bq[3] = {32, (cuuint32_t)kW2Q, 1}, bc[3] = {32, (cuuint32_t)BN, 1}, es[3] = {1, 1, 1};
        CUresult r = encode(&map_q, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, aug_q, dq, sq, bq, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r == CUDA_SUCCESS)
            r = encode(&map_c, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, aug_c, dc, sc, bc, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(OGMM_EUNSUPPORTED);
    }
    Wide2Args a{src, s_sb, s_sn, s_sc, dst, d_sb, d_sn, d_sc, (int)N, (int)M, (int)C, (int)k, normalize, BN, KB, cn_max, idx_out, dist_out, stats};
    dim3 grid((unsigned)((N + kW2Q - 1) / kW2Q), (unsigned)B);
#define LAUNCH2(KK, SS)                                                                                              \
    do {                                                                                                             \
        st = cuda_status(cudaFuncSetAttribute(knn_wide2_kernel<KK, SS>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                              (int)smem), "cudaFuncSetAttribute(knn_wide2_kernel)");                 \
        if (st != OGMM_OK) return fail(st);                                                                          \
        knn_wide2_kernel<KK, SS><<<grid, 128 + 128 * SS, smem, s>>>(map_q, map_c, a);                                \
    } while (0)
#define LAUNCH(KK)                                                                                                   \
    do {                                                                                                             \
        if (sel == 2) LAUNCH2(KK, 2);                                                                                \
        else LAUNCH2(KK, 1);                                                                                         \
    } while (0)
    if (k <= 8) LAUNCH(8);
    else if (k <= 16) LAUNCH(16);
    else if (k <= 20) LAUNCH(20);
    else LAUNCH(32);
#undef LAUNCH
#undef LAUNCH2
    st = cuda_status(cudaGetLastError(), "knn_wide2_kernel");
    cudaFreeAsync(aug_q, s);
    return st;
}
