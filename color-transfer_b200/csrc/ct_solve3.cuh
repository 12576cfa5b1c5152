// The tiny 3x3 algebra of the linear transfers, run by one warp (one lane per pair of a batch):
// raw moments -> mean / covariance -> the method's transform.  fp64 throughout.
//
//   Reinhard  linear.py:33-38   scale_c = std_r / std_t in Lab (population std, ddof 0)
//   Xiao CCS  linear.py:64-80   T = U_t S_t^-1/2 S_r^1/2 U_r^-1, applied as  @ T.T
//   MKL       linear.py:103-122 "MK" / "sqrt" via principal square roots, "cholesky"; applied as @ T
//
// np.linalg.svd of a symmetric PSD matrix == its eigen-decomposition sorted by descending
// eigenvalue; scipy.linalg.sqrtm of a symmetric PSD matrix == V sqrt(L) V^T (the principal root
// is unique).  Both come from a cyclic Jacobi eigen-solver.  The SVD leaves the sign of each
// singular vector free (LAPACK's choice follows no rule); for CCS we fix it by requiring
// dot(u_r_i, u_t_i) >= 0, see DESIGN.md "Xiao sign convention".
#pragma once

#include "ct_common.cuh"

namespace ct {
namespace solve3 {

struct Sym3 {  // symmetric 3x3: 00 01 02 11 12 22
    double a00, a01, a02, a11, a12, a22;
};
struct Mat3 {
    double m[3][3];
};

__device__ __forceinline__ Mat3 matmul(const Mat3 &a, const Mat3 &b) {
    Mat3 c;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            c.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j];
    return c;
}
__device__ __forceinline__ Mat3 full(const Sym3 &s) {
    Mat3 r;
    r.m[0][0] = s.a00; r.m[0][1] = s.a01; r.m[0][2] = s.a02;
    r.m[1][0] = s.a01; r.m[1][1] = s.a11; r.m[1][2] = s.a12;
    r.m[2][0] = s.a02; r.m[2][1] = s.a12; r.m[2][2] = s.a22;
    return r;
}
__device__ __forceinline__ Sym3 symmetric_part(const Mat3 &a) {
    Sym3 s;
    s.a00 = a.m[0][0]; s.a11 = a.m[1][1]; s.a22 = a.m[2][2];
    s.a01 = 0.5 * (a.m[0][1] + a.m[1][0]);
    s.a02 = 0.5 * (a.m[0][2] + a.m[2][0]);
    s.a12 = 0.5 * (a.m[1][2] + a.m[2][1]);
    return s;
}

// Cyclic Jacobi: A = V diag(w) V^T, columns of V are eigenvectors, w sorted descending.
__device__ inline void eigh3(const Sym3 &s, double (&w)[3], Mat3 &V) {
    double a[3][3] = {{s.a00, s.a01, s.a02}, {s.a01, s.a11, s.a12}, {s.a02, s.a12, s.a22}};
    double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int sweep = 0; sweep < 32; ++sweep) {
        const double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
        const double diag = a[0][0] * a[0][0] + a[1][1] * a[1][1] + a[2][2] * a[2][2];
        if (!(off > 1e-34 * diag)) break;   // also leaves on NaN
#pragma unroll
        for (int pq = 0; pq < 3; ++pq) {
            const int p = pq == 2 ? 1 : 0;
            const int q = pq == 0 ? 1 : 2;
            const double apq = a[p][q];
            if (apq == 0.0) continue;
            const double theta = (a[q][q] - a[p][p]) / (2.0 * apq);
            const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
            const int r = 3 - p - q;
            const double arp = a[r][p], arq = a[r][q];
            a[p][p] -= t * apq;
            a[q][q] += t * apq;
            a[p][q] = a[q][p] = 0.0;
            a[r][p] = a[p][r] = c * arp - sn * arq;
            a[r][q] = a[q][r] = sn * arp + c * arq;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const double vkp = v[k][p], vkq = v[k][q];
                v[k][p] = c * vkp - sn * vkq;
                v[k][q] = sn * vkp + c * vkq;
            }
        }
    }
    // sort descending (3-element network), permuting the eigenvector columns with it
    int idx[3] = {0, 1, 2};
    double d[3] = {a[0][0], a[1][1], a[2][2]};
#define CT_SWAP_IF(i, j)                                                     \
    if (d[i] < d[j]) {                                                       \
        double td = d[i]; d[i] = d[j]; d[j] = td;                            \
        int ti = idx[i]; idx[i] = idx[j]; idx[j] = ti;                       \
    }
    CT_SWAP_IF(0, 1) CT_SWAP_IF(1, 2) CT_SWAP_IF(0, 1)
#undef CT_SWAP_IF
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        w[j] = d[j];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            // idx[j] is dynamic: select without local-memory indexing
            const int c = idx[j];
            V.m[k][j] = c == 0 ? v[k][0] : (c == 1 ? v[k][1] : v[k][2]);
        }
    }
}

// V diag(f(w)) V^T
__device__ __forceinline__ Mat3 spectral(const Mat3 &V, const double (&f)[3]) {
    Mat3 r;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            r.m[i][j] = V.m[i][0] * f[0] * V.m[j][0] + V.m[i][1] * f[1] * V.m[j][1] +
                        V.m[i][2] * f[2] * V.m[j][2];
    return r;
}

// lower Cholesky factor; returns false when the matrix is not positive definite
__device__ inline bool cholesky3(const Sym3 &s, Mat3 &L) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) L.m[i][j] = 0.0;
    if (!(s.a00 > 0.0)) return false;
    L.m[0][0] = sqrt(s.a00);
    L.m[1][0] = s.a01 / L.m[0][0];
    L.m[2][0] = s.a02 / L.m[0][0];
    const double d1 = s.a11 - L.m[1][0] * L.m[1][0];
    if (!(d1 > 0.0)) return false;
    L.m[1][1] = sqrt(d1);
    L.m[2][1] = (s.a12 - L.m[2][0] * L.m[1][0]) / L.m[1][1];
    const double d2 = s.a22 - L.m[2][0] * L.m[2][0] - L.m[2][1] * L.m[2][1];
    if (!(d2 > 0.0)) return false;
    L.m[2][2] = sqrt(d2);
    return true;
}
__device__ inline Mat3 inverse_lower(const Mat3 &L) {
    Mat3 r;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) r.m[i][j] = 0.0;
    r.m[0][0] = 1.0 / L.m[0][0];
    r.m[1][1] = 1.0 / L.m[1][1];
    r.m[2][2] = 1.0 / L.m[2][2];
    r.m[1][0] = -L.m[1][0] * r.m[0][0] * r.m[1][1];
    r.m[2][1] = -L.m[2][1] * r.m[1][1] * r.m[2][2];
    r.m[2][0] = -(L.m[2][0] * r.m[0][0] + L.m[2][1] * r.m[1][0]) * r.m[2][2];
    return r;
}

struct Stats {  // decoded moments of one image
    double n, mean[3];
    Sym3 m2;    // sum (x-mean)(x-mean)^T
};

// sums = { n, S(x-K)[3], S(x-K)(x-K)^T[6] }  (K = shift)
__device__ inline Stats decode(const double *sums, const double (&shift)[3]) {
    Stats st;
    st.n = sums[0];
    const double inv_n = 1.0 / st.n;
    const double s0 = sums[1], s1 = sums[2], s2 = sums[3];
    st.mean[0] = shift[0] + s0 * inv_n;
    st.mean[1] = shift[1] + s1 * inv_n;
    st.mean[2] = shift[2] + s2 * inv_n;
    st.m2.a00 = sums[4] - s0 * s0 * inv_n;
    st.m2.a01 = sums[5] - s0 * s1 * inv_n;
    st.m2.a02 = sums[6] - s0 * s2 * inv_n;
    st.m2.a11 = sums[7] - s1 * s1 * inv_n;
    st.m2.a12 = sums[8] - s1 * s2 * inv_n;
    st.m2.a22 = sums[9] - s2 * s2 * inv_n;
    return st;
}
__device__ __forceinline__ Sym3 scaled(const Sym3 &s, double f) {
    return Sym3{s.a00 * f, s.a01 * f, s.a02 * f, s.a11 * f, s.a12 * f, s.a22 * f};
}

// Writes xform = { M[9], mu_t[3], mu_r[3], 0 } and returns a ct_status.
__device__ inline int solve(int method, const double *sums_t, const double *sums_r, double *xform) {
    const bool is_lab = method == CT_REINHARD;
    const double shift[3] = {is_lab ? 50.0 : 0.5, is_lab ? 0.0 : 0.5, is_lab ? 0.0 : 0.5};
    const Stats t = decode(sums_t, shift), r = decode(sums_r, shift);
    Mat3 M;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) M.m[i][j] = 0.0;
    int status = CT_OK;

    if (method == CT_REINHARD) {
        // np.std: sqrt(sum (x-mean)^2 / n)
        M.m[0][0] = sqrt(r.m2.a00 / r.n) / sqrt(t.m2.a00 / t.n);
        M.m[1][1] = sqrt(r.m2.a11 / r.n) / sqrt(t.m2.a11 / t.n);
        M.m[2][2] = sqrt(r.m2.a22 / r.n) / sqrt(t.m2.a22 / t.n);
    } else {
        // np.cov: divide by n - 1
        const Sym3 ct_ = scaled(t.m2, 1.0 / (t.n - 1.0)), cr = scaled(r.m2, 1.0 / (r.n - 1.0));
        if (method == CT_MKL_CHOLESKY) {
            Mat3 A, B;
            if (!cholesky3(ct_, A) || !cholesky3(cr, B)) status = CT_E_NOT_PD;
            M = matmul(B, inverse_lower(A));
        } else {
            double wt[3];
            Mat3 Vt;
            eigh3(ct_, wt, Vt);
            if (method == CT_CCS) {
                double wr[3];
                Mat3 Vr;
                eigh3(cr, wr, Vr);
                // T = U_t diag(sqrt(s_r / s_t)) U_r^T with u_r_i oriented along u_t_i;
                // the remap applies T.T, so M = T.T = U_r diag(.) U_t^T.
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const double d = Vr.m[0][k] * Vt.m[0][k] + Vr.m[1][k] * Vt.m[1][k] + Vr.m[2][k] * Vt.m[2][k];
                    const double g = (d < 0 ? -1.0 : 1.0) * (1.0 / sqrt(wt[k])) * sqrt(wr[k]);
#pragma unroll
                    for (int i = 0; i < 3; ++i)
#pragma unroll
                        for (int j = 0; j < 3; ++j) M.m[i][j] += Vr.m[i][k] * g * Vt.m[j][k];
                }
                if (!(wt[2] > 0.0)) status = CT_E_SINGULAR;
            } else {
                const double rt[3] = {sqrt(wt[0]), sqrt(wt[1]), sqrt(wt[2])};
                const double irt[3] = {1.0 / rt[0], 1.0 / rt[1], 1.0 / rt[2]};
                const Mat3 A = spectral(Vt, rt), Ainv = spectral(Vt, irt);
                if (!(wt[2] > 0.0)) status = CT_E_SINGULAR;
                if (method == CT_MKL_SQRT) {
                    double wr[3];
                    Mat3 Vr;
                    eigh3(cr, wr, Vr);
                    const double rr[3] = {sqrt(wr[0]), sqrt(wr[1]), sqrt(wr[2])};
                    M = matmul(spectral(Vr, rr), Ainv);
                } else {  // CT_MKL_MK
                    const Sym3 C = symmetric_part(matmul(matmul(A, full(cr)), A));
                    double wc[3];
                    Mat3 Vc;
                    eigh3(C, wc, Vc);
                    const double rc[3] = {sqrt(wc[0]), sqrt(wc[1]), sqrt(wc[2])};
                    M = matmul(matmul(Ainv, spectral(Vc, rc)), Ainv);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) xform[3 * i + j] = M.m[i][j];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        xform[9 + c] = t.mean[c];
        xform[12 + c] = r.mean[c];
    }
    xform[15] = 0.0;
    return status;
}

}  // namespace solve3
}  // namespace ct
