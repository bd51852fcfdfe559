// Pieces shared by the 3-D kNN kernels (knn_sweep.cu, knn_select.cu): the register-resident (distance, index) list with
// 64-bit keys, and the CTA-wide bitonic sorts of the cloud along one axis.
#pragma once
#include "common.cuh"

namespace ogmm {

constexpr int kSwThreads = 256;
constexpr int kSwStage = 12;          // trigger (4) + one step of both sides (8)
constexpr int kSwTrigger = 4;            // measured on B200: 0 -> 1.019, 2 -> 0.991, 4 -> 0.978, 8 -> 1.006, 16 -> 1.019 ms/step
constexpr int kSwQueriesPerCta = 256;  // each CTA re-sorts the cloud (cheap) and owns 256 sorted query ranks

typedef unsigned long long u64;
constexpr u64 kEmptyKey = (0x7f800000ull << 32) | 0xffffffffull;      // distance +inf, largest index

template <int K>
struct TopK64 {
    u64 key[K];
    float thr;
    int cnt;
    bool live;
    float* sd;
    unsigned short* si;                 // staged candidate indices: the sweep kernel takes M <= 4096
    int col;

    __device__ __forceinline__ void init(float* stage_d, unsigned short* stage_i, int col_, bool live_) {
#pragma unroll
        for (int j = 0; j < K; ++j) key[j] = kEmptyKey;
        live = live_;
        thr = live ? INFINITY : -INFINITY;
        cnt = 0; sd = stage_d; si = stage_i; col = col_;
    }
    // v <= thr: equal distances are decided by index at merge time
    __device__ __forceinline__ void offer(float v, int idx) {
        if (v <= thr) {
            sd[cnt * kSwThreads + col] = v;
            si[cnt * kSwThreads + col] = (unsigned short)idx;
            ++cnt;
        }
    }
    // Sorted insert: strict compare up to the insertion point, then every entry shifts down by one.
    // (A dependency-free variant -- all compares first, then independent selects -- measured 10 % slower:
    // this kernel is bound by instruction count, not by the carried chain.)
    template <int A>
    __device__ __forceinline__ void insert_from(u64 kv) {      // positions A .. K-1; the caller knows kv >= key[A-1]
        bool moved = false;
#pragma unroll
        for (int j = A; j < K; ++j) {
            moved = moved || (kv < key[j]);
            const u64 t = key[j];
            key[j] = moved ? kv : t;
            kv = moved ? t : kv;
        }
    }
    // Candidates met late in the sweep barely beat the threshold, so they land near the tail of the list: when no
    // lane of the warp lands in the first half (or three quarters), only the tail is shifted.
    __device__ __forceinline__ void insert(u64 kv) {
        if constexpr (K >= 16) {
            constexpr int Q3 = (3 * K) / 4, Q2 = K / 2;
            if (!__any_sync(kFull, kv < key[Q3 - 1])) { insert_from<Q3>(kv); return; }
            if (!__any_sync(kFull, kv < key[Q2 - 1])) { insert_from<Q2>(kv); return; }
        }
        insert_from<0>(kv);
    }
    __device__ __forceinline__ void merge() {
        const int most = __reduce_max_sync(kFull, cnt);
        for (int s = 0; s < most; ++s) {
            u64 kv = kEmptyKey;
            if (s < cnt) kv = ((u64)__float_as_uint(sd[s * kSwThreads + col]) << 32) | (unsigned)si[s * kSwThreads + col];
            if (__any_sync(kFull, kv < key[K - 1])) insert(kv);
        }
        cnt = 0;
        thr = live ? __uint_as_float((unsigned)(key[K - 1] >> 32)) : -INFINITY;     // +inf bits until the list is full
    }
    __device__ __forceinline__ void maybe_merge() {
        if (__any_sync(kFull, cnt > kSwTrigger)) merge();
    }
};

// (key, index) bitonic sort, ascending, lexicographic; n is a power of two.
__device__ __forceinline__ void bitonic_sort_pairs(float* key, int* val, int n) {
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1, lj = 31 - __clz(k >> 1); j > 0; j >>= 1, --lj) {
            for (int t = threadIdx.x; t < (n >> 1); t += kSwThreads) {
                const int i = ((t >> lj) << (lj + 1)) + (t & (j - 1));
                const int l = i + j;
                const bool up = ((i & k) == 0);
                const float a = key[i], b = key[l];
                const int ai = val[i], bi = val[l];
                const bool a_gt_b = (a > b) || (a == b && ai > bi);
                if (a_gt_b == up) { key[i] = b; key[l] = a; val[i] = bi; val[l] = ai; }
            }
            __syncthreads();
        }
    }
}

// The same sort for n = 256 * E with E elements per thread held in registers: compare-exchange distances below E
// stay inside the thread, distances below 32 E go through warp shuffles, and only the log2(8) (log2(8) + 1) / 2 = 6
// cross-warp stages per sort touch shared memory and the block barrier (the shared-memory version above pays a
// barrier in every one of its 55 stages at n = 1024, which made the sort ~30 % of the sweep kernel's warp time).
// Fully unrolled: every register index is a compile-time constant.
template <int E>
__device__ __forceinline__ void bitonic_sort_pairs_regs(float* key, int* val) {
    constexpr int n = kSwThreads * E;
    const int t = threadIdx.x;
    float kr[E];
    int vr[E];
#pragma unroll
    for (int e = 0; e < E; ++e) { kr[e] = key[t * E + e]; vr[e] = val[t * E + e]; }
    __syncthreads();                                   // the arrays become the exchange scratch ([e][thread] layout)
#pragma unroll
    for (int k = 2; k <= n; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j < E) {
                // partner inside the thread: elements e and e | j
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    if ((e & j) == 0) {
                        const int f = e | j;
                        const bool up = (k < E) ? ((e & k) == 0) : (((t * E) & k) == 0);
                        const bool gt = (kr[e] > kr[f]) || (kr[e] == kr[f] && vr[e] > vr[f]);
                        if (gt == up) {
                            const float tk = kr[e]; kr[e] = kr[f]; kr[f] = tk;
                            const int tv = vr[e]; vr[e] = vr[f]; vr[f] = tv;
                        }
                    }
                }
            } else {
                const int pt = j / E;                  // partner thread = t ^ pt, same e
                const bool lower = (t & pt) == 0;
                const bool up = ((t * E) & k) == 0;    // k > j >= E: the direction bit lies in the thread index
                const bool take_min = lower == up;
                if (pt >= 32) {
#pragma unroll
                    for (int e = 0; e < E; ++e) { key[e * kSwThreads + t] = kr[e]; val[e * kSwThreads + t] = vr[e]; }
                    __syncthreads();
                }
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    float ok;
                    int ov;
                    if (pt >= 32) { ok = key[e * kSwThreads + (t ^ pt)]; ov = val[e * kSwThreads + (t ^ pt)]; }
                    else { ok = __shfl_xor_sync(kFull, kr[e], pt); ov = __shfl_xor_sync(kFull, vr[e], pt); }
                    const bool gt = (kr[e] > ok) || (kr[e] == ok && vr[e] > ov);      // mine after the partner's
                    if (gt == take_min) { kr[e] = ok; vr[e] = ov; }
                }
                if (pt >= 32) __syncthreads();
            }
        }
    }
#pragma unroll
    for (int e = 0; e < E; ++e) { key[t * E + e] = kr[e]; val[t * E + e] = vr[e]; }
    __syncthreads();
}

// dispatch: register version for the sizes the sweep kernel meets (256 .. 4096 points), generic otherwise
__device__ __forceinline__ void sort_pairs(float* key, int* val, int n) {
    switch (n) {
        case 256:  bitonic_sort_pairs_regs<1>(key, val); break;
        case 512:  bitonic_sort_pairs_regs<2>(key, val); break;
        case 1024: bitonic_sort_pairs_regs<4>(key, val); break;
        case 2048: bitonic_sort_pairs_regs<8>(key, val); break;
        default:   bitonic_sort_pairs(key, val, n); break;
    }
}

// Axis key of a point for the CTA sort: NaN coordinates sort like +inf, so the comparator stays a total order and the
// sorted order is a permutation of the cloud whatever the input holds (every query row gets written).
__device__ __forceinline__ float sort_key(float v) { return v == v ? v : INFINITY; }

__device__ __forceinline__ float sqn3(float x, float y, float z) {
    return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
}

__host__ __device__ inline int pow2_ge(int v) { int p = 1; while (p < v) p <<= 1; return p; }

// ---- packed FP32x2 distances of a group of four candidates (shared by the selection and the exhaustive kernels) ----
typedef unsigned long long u64t;

__device__ __forceinline__ u64t pack2(float lo, float hi) {
    return ((u64t)__float_as_uint(hi) << 32) | (u64t)__float_as_uint(lo);
}
__device__ __forceinline__ u64t mul2(u64t a, u64t b) { u64t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64t fma2(u64t a, u64t b, u64t c) { u64t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64t add2(u64t a, u64t b) { u64t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float lo32(u64t v) { return __uint_as_float((unsigned)(v & 0xffffffffull)); }
__device__ __forceinline__ float hi32(u64t v) { return __uint_as_float((unsigned)(v >> 32)); }

struct QueryPack { u64t x, y, z, s; };     // (qx,qx), (qy,qy), (qz,qz), (|q|^2,|q|^2)

// Distances of the four candidates of one group, unclamped, in the reference's operation order (two candidates per
// packed instruction; each half is an independent IEEE fp32 operation, so the values equal the scalar kernels' bit for
// bit).  Record layout per group (64 bytes): (xA,xB,yA,yB) (zA,zB,wA,wB) (xC,xD,yC,yD) (zC,zD,wC,wD), coordinates
// pre-scaled by -2.  `rec_addr` is the SHARED-window byte address of the group's record (explicit ld.shared: the
// generic-pointer form re-derives the window base with S2R + LEA at every use).
__device__ __forceinline__ void lds_pair(unsigned addr, u64t& a, u64t& b) {
    asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr));
}
__device__ __forceinline__ void group_distances(unsigned rec_addr, const QueryPack& q, float (&v)[4]) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        u64t xx, yy, zz, ww;
        lds_pair(rec_addr + 32 * h, xx, yy);
        lds_pair(rec_addr + 32 * h + 16, zz, ww);
        u64t d = mul2(q.x, xx);
        d = fma2(q.y, yy, d);
        d = fma2(q.z, zz, d);
        d = add2(d, q.s);
        d = add2(d, ww);
        v[2 * h] = lo32(d);
        v[2 * h + 1] = hi32(d);
    }
}


}  // namespace ogmm
