// Register-resident 3x3 SVD (one-sided Jacobi / Hestenes), host+device.
//
// A = U diag(S) V^T with S sorted descending, U and V orthogonal.  Used by the weighted
// Procrustes kernels in place of the reference's host LAPACK call (lib/se3.py:276,
// baseline/deepgmr.py:29).  Matrices are row-major T[9].
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define OGMM_HD __host__ __device__ __forceinline__
#else
#define OGMM_HD inline
#endif

namespace ogmm {

template <typename T>
OGMM_HD void jacobi_pair(T* b, T* v, int p, int q, bool& rotated) {
    // columns p,q of the 3x3 row-major matrices b (working copy of A) and v
    T alpha = b[p] * b[p] + b[3 + p] * b[3 + p] + b[6 + p] * b[6 + p];
    T beta = b[q] * b[q] + b[3 + q] * b[3 + q] + b[6 + q] * b[6 + q];
    T gamma = b[p] * b[q] + b[3 + p] * b[3 + q] + b[6 + p] * b[6 + q];
    const T eps = sizeof(T) == 8 ? (T)1e-15 : (T)1e-7;
    if (gamma == (T)0 || fabs((double)gamma) <= (double)eps * sqrt((double)alpha * (double)beta)) return;
    rotated = true;
    T zeta = (beta - alpha) / ((T)2 * gamma);
    T t = (zeta >= (T)0 ? (T)1 : (T)-1) / ((T)fabs((double)zeta) + (T)sqrt((double)((T)1 + zeta * zeta)));
    T c = (T)1 / (T)sqrt((double)((T)1 + t * t));
    T s = c * t;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        T bp = b[3 * r + p], bq = b[3 * r + q];
        b[3 * r + p] = c * bp - s * bq;
        b[3 * r + q] = s * bp + c * bq;
        T vp = v[3 * r + p], vq = v[3 * r + q];
        v[3 * r + p] = c * vp - s * vq;
        v[3 * r + q] = s * vp + c * vq;
    }
}

template <typename T>
OGMM_HD void swap_cols(T* m, int a, int b) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        T t = m[3 * r + a];
        m[3 * r + a] = m[3 * r + b];
        m[3 * r + b] = t;
    }
}

template <typename T>
OGMM_HD void svd3(const T* A, T* U, T* S, T* V) {
    T b[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        b[i] = A[i];
        V[i] = (i % 4 == 0) ? (T)1 : (T)0;
    }
    for (int sweep = 0; sweep < 30; ++sweep) {
        bool rotated = false;
        jacobi_pair(b, V, 0, 1, rotated);
        jacobi_pair(b, V, 0, 2, rotated);
        jacobi_pair(b, V, 1, 2, rotated);
        if (!rotated) break;
    }
    T n[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) n[c] = (T)sqrt((double)(b[c] * b[c] + b[3 + c] * b[3 + c] + b[6 + c] * b[6 + c]));
    // sort columns by singular value, descending (3-element network)
    if (n[0] < n[1]) { T t = n[0]; n[0] = n[1]; n[1] = t; swap_cols(b, 0, 1); swap_cols(V, 0, 1); }
    if (n[1] < n[2]) { T t = n[1]; n[1] = n[2]; n[2] = t; swap_cols(b, 1, 2); swap_cols(V, 1, 2); }
    if (n[0] < n[1]) { T t = n[0]; n[0] = n[1]; n[1] = t; swap_cols(b, 0, 1); swap_cols(V, 0, 1); }
    S[0] = n[0]; S[1] = n[1]; S[2] = n[2];
    const T tiny = (sizeof(T) == 8 ? (T)1e-13 : (T)1e-6) * (n[0] > (T)0 ? n[0] : (T)1);
    // U columns: normalised columns of b where the singular value is significant,
    // orthonormal completion otherwise (rank-deficient input).
    T u0[3], u1[3], u2[3];
    if (n[0] > tiny) { u0[0] = b[0] / n[0]; u0[1] = b[3] / n[0]; u0[2] = b[6] / n[0]; }
    else { u0[0] = 1; u0[1] = 0; u0[2] = 0; }
    if (n[1] > tiny) { u1[0] = b[1] / n[1]; u1[1] = b[4] / n[1]; u1[2] = b[7] / n[1]; }
    else {
        // any unit vector orthogonal to u0
        int k = 0;
        T m = (T)fabs((double)u0[0]);
        if ((T)fabs((double)u0[1]) < m) { k = 1; m = (T)fabs((double)u0[1]); }
        if ((T)fabs((double)u0[2]) < m) { k = 2; }
        T e[3] = {(T)0, (T)0, (T)0};
        e[k] = (T)1;
        T d = u0[k];
        T w[3] = {e[0] - d * u0[0], e[1] - d * u0[1], e[2] - d * u0[2]};
        T wn = (T)sqrt((double)(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]));
        u1[0] = w[0] / wn; u1[1] = w[1] / wn; u1[2] = w[2] / wn;
    }
    if (n[2] > tiny) { u2[0] = b[2] / n[2]; u2[1] = b[5] / n[2]; u2[2] = b[8] / n[2]; }
    else {
        u2[0] = u0[1] * u1[2] - u0[2] * u1[1];
        u2[1] = u0[2] * u1[0] - u0[0] * u1[2];
        u2[2] = u0[0] * u1[1] - u0[1] * u1[0];
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) { U[3 * r] = u0[r]; U[3 * r + 1] = u1[r]; U[3 * r + 2] = u2[r]; }
}

template <typename T>
OGMM_HD T det3(const T* m) {
    return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
}

// R = V * diag(1,1,d) * U^T (row-major), the rotation both registration heads need.
template <typename T>
OGMM_HD void v_d_ut(const T* V, const T* U, T d, T* R) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            R[3 * i + j] = V[3 * i] * U[3 * j] + V[3 * i + 1] * U[3 * j + 1] + d * V[3 * i + 2] * U[3 * j + 2];
}

// lib/se3.py:275-285: R = V U^T; when det(R) <= 0 use V with its third column negated.
template <typename T>
OGMM_HD void rotation_from_cov_ogmm(const T* cov, T* R) {
    T U[9], S[3], V[9];
    svd3(cov, U, S, V);
    v_d_ut(V, U, (T)1, R);
    if (!(det3(R) > (T)0)) v_d_ut(V, U, (T)-1, R);
}

// baseline/deepgmr.py:29-34: R = V diag(1,1,det(V U^T)) U^T.
template <typename T>
OGMM_HD void rotation_from_cov_deepgmr(const T* m, T* R) {
    T U[9], S[3], V[9];
    svd3(m, U, S, V);
    T P[9];
    v_d_ut(V, U, (T)1, P);
    v_d_ut(V, U, det3(P), R);
}

}  // namespace ogmm
