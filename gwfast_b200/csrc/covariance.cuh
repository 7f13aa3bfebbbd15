// Batched covariance / conditioning of Fisher matrices (gwfast/fisherTools.py:32-196 CovMatr, :199-211
// compute_inversion_error, :216-279 CheckFisher) -- the immediate consumer of the Fisher kernel.
//
// The reference loops over events in Python and inverts each matrix with mpmath (diagonal normalisation
// ws F ws, Cholesky, (c^T c), symmetrisation, un-normalisation; eigenvalue/SVD route for matrices that are not
// positive definite).  Here one thread owns one event; the nP x nP problem lives in that thread's local memory
// in double-double arithmetic (~106 bits), so the result is at least as accurate as the reference's 53-bit
// mpmath arithmetic, and events are independent (coalesced loads/stores on the event-fastest (nP,nP,N) layout).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace gwf {

constexpr int kCovMaxP = 16;   // largest parameter count handled (gwfast: 9..15 (+2 with priors on extra rows))

// ------------------------------------------------------------------ double-double arithmetic (error-free transformations)
struct dd {
    double hi, lo;
};
__host__ __device__ __forceinline__ dd dd_make(double a) { return dd{a, 0.0}; }
__host__ __device__ __forceinline__ dd two_sum(double a, double b) {
    const double s = a + b, bb = s - a;
    return dd{s, (a - (s - bb)) + (b - bb)};
}
__host__ __device__ __forceinline__ dd quick_two_sum(double a, double b) {
    const double s = a + b;
    return dd{s, b - (s - a)};
}
__host__ __device__ __forceinline__ dd two_prod(double a, double b) {
    const double p = a * b;
    return dd{p, fma(a, b, -p)};
}
__host__ __device__ __forceinline__ dd operator+(dd a, dd b) {
    dd s = two_sum(a.hi, b.hi);
    const dd t = two_sum(a.lo, b.lo);
    s.lo += t.hi;
    s = quick_two_sum(s.hi, s.lo);
    s.lo += t.lo;
    return quick_two_sum(s.hi, s.lo);
}
__host__ __device__ __forceinline__ dd operator-(dd a) { return dd{-a.hi, -a.lo}; }
__host__ __device__ __forceinline__ dd operator-(dd a, dd b) { return a + (-b); }
__host__ __device__ __forceinline__ dd operator*(dd a, dd b) {
    dd p = two_prod(a.hi, b.hi);
    p.lo += a.hi * b.lo + a.lo * b.hi;
    return quick_two_sum(p.hi, p.lo);
}
__host__ __device__ __forceinline__ dd operator*(dd a, double b) {
    dd p = two_prod(a.hi, b);
    p.lo += a.lo * b;
    return quick_two_sum(p.hi, p.lo);
}
__host__ __device__ __forceinline__ dd operator/(dd a, dd b) {
    const double q1 = a.hi / b.hi;
    dd r = a - b * q1;
    const double q2 = r.hi / b.hi;
    r = r - b * q2;
    const double q3 = r.hi / b.hi;
    const dd q = quick_two_sum(q1, q2);
    return q + dd_make(q3);
}
__host__ __device__ __forceinline__ dd dd_sqrt(dd a) {
    if (a.hi <= 0.0) return dd{a.hi == 0.0 ? 0.0 : NAN, 0.0};
    const double x = 1.0 / sqrt(a.hi), ax = a.hi * x;
    const dd e = a - two_prod(ax, ax);
    return two_sum(ax, e.hi * (x * 0.5));
}
__host__ __device__ __forceinline__ dd dd_abs(dd a) { return a.hi < 0.0 ? -a : a; }

// status codes written per event
enum CovStatus {
    kCovOkCholesky = 0,     // positive definite: Cholesky route (invMethodIn='cho')
    kCovOkEigen = 1,        // not positive definite (or Cholesky broke down): symmetric eigen-decomposition route, the
                            // reference's alt_method='svd' (for a symmetric matrix the SVD inverse is V diag(1/lambda) V^T)
    kCovNaNInput = 2,       // Fisher all-NaN: NaN covariance (fisherTools.py:64-67)
    kCovFailed = 3,         // non-finite result
    kCovZeroDiag = 4,       // a zero on the diagonal: normalisation skipped (fisherTools.py:97-99)
};

// cyclic Jacobi eigen-decomposition of the symmetric n x n matrix a (destroyed; eigenvalues end on its diagonal), v = eigenvectors
// in columns.  Double-double throughout; converges quadratically, 12 sweeps is far more than needed for n <= 16.
__host__ __device__ inline void dd_jacobi(dd* a, dd* v, int n) {
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) v[i * n + j] = dd_make(i == j ? 1.0 : 0.0);
    for (int sweep = 0; sweep < 14; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < n; ++i) {
            diag += a[i * n + i].hi * a[i * n + i].hi;
            for (int j = i + 1; j < n; ++j) off += a[i * n + j].hi * a[i * n + j].hi;
        }
        if (!(off > 1e-62 * diag)) break;      // off-diagonal below double-double resolution (also exits on NaN)
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) {
                const dd apq = a[p * n + q];
                if (apq.hi == 0.0) continue;
                const dd app = a[p * n + p], aqq = a[q * n + q];
                // tan of the rotation angle: t = sgn(theta) / (|theta| + sqrt(theta^2 + 1)), theta = (aqq - app) / (2 apq)
                const dd theta = (aqq - app) / (apq * 2.0);
                const dd at = dd_abs(theta);
                dd t = dd_make(1.0) / (at + dd_sqrt(at * at + dd_make(1.0)));
                if (theta.hi < 0.0) t = -t;
                const dd c = dd_make(1.0) / dd_sqrt(t * t + dd_make(1.0)), s = t * c;
                a[p * n + p] = app - t * apq;
                a[q * n + q] = aqq + t * apq;
                a[p * n + q] = a[q * n + p] = dd_make(0.0);
                for (int k = 0; k < n; ++k) {
                    if (k != p && k != q) {
                        const dd akp = a[k * n + p], akq = a[k * n + q];
                        const dd np_ = c * akp - s * akq, nq_ = s * akp + c * akq;
                        a[k * n + p] = a[p * n + k] = np_;
                        a[k * n + q] = a[q * n + k] = nq_;
                    }
                    const dd vkp = v[k * n + p], vkq = v[k * n + q];
                    v[k * n + p] = c * vkp - s * vkq;
                    v[k * n + q] = s * vkp + c * vkq;
                }
            }
    }
}

// inversion methods (invMethodIn of CovMatr, fisherTools.py:33-49)
enum CovMethod {
    kCovCholesky = 0,   // 'cho' (also serves 'inv' and 'lu': the same inverse), eigen route if not positive definite
    kCovEigen = 1,      // 'svd': V diag(1/lambda) V^T
    kCovEigenTrunc = 2, // 'svd' with truncate=True: singular values below thresh * max are raised to thresh * max (fisherTools.py:137-142)
    kCovEigenReg = 3,   // 'svd_reg': singular values <= thresh (absolute) are excluded (fisherTools.py:164-175)
};

// One event of CovMatr.  F, C: element (i, j) at [(i * nP + j) * stride]; scratch: 3 * nP * nP dd.
// Returns the status; *inv_err = max |C F - 1| (compute_inversion_error, evaluated in double like the reference).
__host__ __device__ inline int cov_one(const double* __restrict__ F, long long stride, int nP, int method, double thresh, double* __restrict__ C,
                                       double* __restrict__ inv_err, dd* __restrict__ A, dd* __restrict__ L, dd* __restrict__ V) {
    const int n = nP;
    bool all_nan = true;
    for (int i = 0; i < n * n; ++i) all_nan = all_nan && isnan(F[(long long)i * stride]);
    if (all_nan) {
        for (int i = 0; i < n * n; ++i) C[(long long)i * stride] = NAN;
        *inv_err = NAN;
        return kCovNaNInput;
    }
    // diagonal normalisation ws F ws (fisherTools.py:82-90)
    dd ws[kCovMaxP];
    bool zero_diag = false;
    for (int i = 0; i < n; ++i) zero_diag = zero_diag || F[(long long)(i * n + i) * stride] == 0.0;
    for (int i = 0; i < n; ++i) ws[i] = zero_diag ? dd_make(1.0) : dd_make(1.0) / dd_sqrt(dd_make(F[(long long)(i * n + i) * stride]));
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) A[i * n + j] = ws[i] * dd_make(F[(long long)(i * n + j) * stride]) * ws[j];
    // Cholesky A = L L^T
    bool pd = method == kCovCholesky;
    for (int j = 0; j < n && pd; ++j) {
        dd d = A[j * n + j];
        for (int k = 0; k < j; ++k) d = d - L[j * n + k] * L[j * n + k];
        if (!(d.hi > 0.0)) { pd = false; break; }
        const dd ljj = dd_sqrt(d);
        L[j * n + j] = ljj;
        for (int i = j + 1; i < n; ++i) {
            dd s = A[i * n + j];
            for (int k = 0; k < j; ++k) s = s - L[i * n + k] * L[j * n + k];
            L[i * n + j] = s / ljj;
        }
    }
    int status = zero_diag ? kCovZeroDiag : kCovOkCholesky;
    if (pd) {
        // c = L^-1 (lower triangular, stored in V), cc = c^T c (fisherTools.py:116, 131)
        for (int j = 0; j < n; ++j) {
            for (int i = 0; i < j; ++i) V[i * n + j] = dd_make(0.0);
            V[j * n + j] = dd_make(1.0) / L[j * n + j];
            for (int i = j + 1; i < n; ++i) {
                dd s = dd_make(0.0);
                for (int k = j; k < i; ++k) s = s - L[i * n + k] * V[k * n + j];
                V[i * n + j] = s / L[i * n + i];
            }
        }
        for (int i = 0; i < n; ++i)
            for (int j = 0; j <= i; ++j) {
                dd s = dd_make(0.0);
                for (int k = i; k < n; ++k) s = s + V[k * n + i] * V[k * n + j];
                A[i * n + j] = A[j * n + i] = s;
            }
    } else {
        // symmetric eigen route: inverse = V diag(1/lambda) V^T
        for (int i = 0; i < n * n; ++i) L[i] = A[i];
        dd_jacobi(L, V, n);
        double smax = 0.0;
        for (int k = 0; k < n; ++k) smax = fmax(smax, fabs(L[k * n + k].hi));
        dd inv[kCovMaxP];
        for (int k = 0; k < n; ++k) {
            const dd lam = L[k * n + k];
            const double sv = fabs(lam.hi);
            if (method == kCovEigenTrunc && !(sv > thresh * smax)) inv[k] = dd_make((lam.hi < 0.0 ? -1.0 : 1.0) / (thresh * smax));
            else if (method == kCovEigenReg && !(sv > thresh)) inv[k] = dd_make(0.0);
            else inv[k] = dd_make(1.0) / lam;
        }
        for (int i = 0; i < n; ++i)
            for (int j = 0; j <= i; ++j) {
                dd s = dd_make(0.0);
                for (int k = 0; k < n; ++k) s = s + V[i * n + k] * V[j * n + k] * inv[k];
                A[i * n + j] = A[j * n + i] = s;
            }
        if (!zero_diag) status = kCovOkEigen;
    }
    // undo the normalisation, write, inversion error
    bool finite = true;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            const dd c = ws[i] * A[i * n + j] * ws[j];
            C[(long long)(i * n + j) * stride] = c.hi;
            finite = finite && isfinite(c.hi);
        }
    // inversion error of the float64 covariance just written; the products are accumulated in double-double (the reference
    // evaluates Cov @ Fisher in float128, fisherTools.py:56-57, 185), so what is measured is the rounding of Cov, not of the check
    double err = 0.0;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            dd s = dd_make(i == j ? -1.0 : 0.0);
            for (int k = 0; k < n; ++k) s = s + two_prod(C[(long long)(i * n + k) * stride], F[(long long)(k * n + j) * stride]);
            const double e = fabs(s.hi);
            err = (e > err || isnan(e)) ? e : err;
        }
    *inv_err = err;
    return finite ? status : kCovFailed;
}

// One event of CheckFisher: eigenvalues (ascending, like mpmath.eigh / scipy.linalg.eigh), eigenvectors in columns, and the
// condition number max|lambda| / min|lambda|.  evals[k * stride], evecs[(i * nP + k) * stride] = component i of eigenvector k.
__host__ __device__ inline void eig_one(const double* __restrict__ F, long long stride, int nP, double* __restrict__ evals, double* __restrict__ evecs,
                                        double* __restrict__ cond, dd* __restrict__ A, dd* __restrict__ V) {
    const int n = nP;
    bool all_nan = true;
    for (int i = 0; i < n * n; ++i) all_nan = all_nan && isnan(F[(long long)i * stride]);
    if (all_nan) {
        for (int k = 0; k < n; ++k) evals[(long long)k * stride] = NAN;
        if (evecs) for (int i = 0; i < n * n; ++i) evecs[(long long)i * stride] = NAN;
        *cond = NAN;
        return;
    }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) A[i * n + j] = dd_make(0.5 * (F[(long long)(i * n + j) * stride] + F[(long long)(j * n + i) * stride]));
    dd_jacobi(A, V, n);
    int order[kCovMaxP];
    for (int k = 0; k < n; ++k) order[k] = k;
    for (int a = 1; a < n; ++a) {           // insertion sort by eigenvalue
        const int o = order[a];
        int b = a - 1;
        while (b >= 0 && A[order[b] * n + order[b]].hi > A[o * n + o].hi) { order[b + 1] = order[b]; --b; }
        order[b + 1] = o;
    }
    double amax = 0.0, amin = INFINITY;
    for (int k = 0; k < n; ++k) {
        const double lam = A[order[k] * n + order[k]].hi;
        evals[(long long)k * stride] = lam;
        amax = fmax(amax, fabs(lam));
        amin = fmin(amin, fabs(lam));
        if (evecs)
            for (int i = 0; i < n; ++i) evecs[(long long)(i * n + k) * stride] = V[i * n + order[k]].hi;
    }
    *cond = amax / amin;
}

#ifdef __CUDACC__
__global__ void __launch_bounds__(64) covariance_kernel(const double* __restrict__ F, long long n, int nP, int method, double thresh, double* __restrict__ C,
                                                        double* __restrict__ inv_err, int* __restrict__ status) {
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= n) return;
    dd A[kCovMaxP * kCovMaxP], L[kCovMaxP * kCovMaxP], V[kCovMaxP * kCovMaxP];
    double err;
    const int st = cov_one(F + e, n, nP, method, thresh, C + e, &err, A, L, V);
    inv_err[e] = err;
    if (status) status[e] = st;
}
__global__ void __launch_bounds__(64) eigen_kernel(const double* __restrict__ F, long long n, int nP, double* __restrict__ evals, double* __restrict__ evecs,
                                                   double* __restrict__ cond) {
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= n) return;
    dd A[kCovMaxP * kCovMaxP], V[kCovMaxP * kCovMaxP];
    double c;
    eig_one(F + e, n, nP, evals + e, evecs ? evecs + e : nullptr, &c, A, V);
    cond[e] = c;
}
// max |C F - 1| per event (compute_inversion_error, fisherTools.py:199-211)
__global__ void __launch_bounds__(128) inversion_error_kernel(const double* __restrict__ F, const double* __restrict__ Cv, long long n, int nP,
                                                              double* __restrict__ err) {
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= n) return;
    double m = 0.0;
    for (int i = 0; i < nP; ++i)
        for (int j = 0; j < nP; ++j) {
            dd s = dd_make(i == j ? -1.0 : 0.0);
            for (int k = 0; k < nP; ++k) s = s + two_prod(Cv[(long long)(i * nP + k) * n + e], F[(long long)(k * nP + j) * n + e]);
            const double d = fabs(s.hi);
            m = (d > m || isnan(d)) ? d : m;
        }
    err[e] = m;
}
#endif

}  // namespace gwf
