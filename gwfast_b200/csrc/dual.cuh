// Forward-mode dual numbers for the per-event prologue (sm_100a, FP64).
//
// The reference differentiates GWstrain with jax.jacrev over (Mc, eta, chi1z, chi2z[, LambdaTilde,
// deltaLambda]) (gwfast/signal.py:1153-1189).  Here the f-independent part of that chain is evaluated once
// per event with Dual<N>; the f-dependent part uses the coefficient/basis expansion of model_*.cuh.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace gwf {

template <int N>
struct Dual {
    double v;
    double d[N];
    __host__ __device__ Dual() {}
    __host__ __device__ Dual(double x) : v(x) {
#pragma unroll
        for (int i = 0; i < N; ++i) d[i] = 0.0;
    }
    __host__ __device__ static Dual seed(double x, int k) {
        Dual r(x);
        r.d[k] = 1.0;
        return r;
    }
};

#define GWF_HD __host__ __device__ __forceinline__

// value accessors that also work for plain doubles
GWF_HD double val(double x) { return x; }
template <int N> GWF_HD double val(const Dual<N>& x) { return x.v; }

// ------------------------------------------------------------------ arithmetic
template <int N> GWF_HD Dual<N> operator-(const Dual<N>& a) {
    Dual<N> r; r.v = -a.v;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = -a.d[i];
    return r;
}
template <int N> GWF_HD Dual<N> operator+(const Dual<N>& a, const Dual<N>& b) {
    Dual<N> r; r.v = a.v + b.v;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i];
    return r;
}
template <int N> GWF_HD Dual<N> operator+(const Dual<N>& a, double b) { Dual<N> r = a; r.v += b; return r; }
template <int N> GWF_HD Dual<N> operator+(double b, const Dual<N>& a) { Dual<N> r = a; r.v += b; return r; }
template <int N> GWF_HD Dual<N> operator-(const Dual<N>& a, const Dual<N>& b) {
    Dual<N> r; r.v = a.v - b.v;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i];
    return r;
}
template <int N> GWF_HD Dual<N> operator-(const Dual<N>& a, double b) { Dual<N> r = a; r.v -= b; return r; }
template <int N> GWF_HD Dual<N> operator-(double b, const Dual<N>& a) { Dual<N> r = -a; r.v += b; return r; }
template <int N> GWF_HD Dual<N> operator*(const Dual<N>& a, const Dual<N>& b) {
    Dual<N> r; r.v = a.v * b.v;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
    return r;
}
template <int N> GWF_HD Dual<N> operator*(const Dual<N>& a, double b) {
    Dual<N> r; r.v = a.v * b;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b;
    return r;
}
template <int N> GWF_HD Dual<N> operator*(double b, const Dual<N>& a) { return a * b; }
template <int N> GWF_HD Dual<N> operator/(const Dual<N>& a, const Dual<N>& b) {
    Dual<N> r; const double inv = 1.0 / b.v; r.v = a.v / b.v;       // value by true division: bit-identical to the scalar path
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * inv;
    return r;
}
template <int N> GWF_HD Dual<N> operator/(const Dual<N>& a, double b) {
    Dual<N> r; const double inv = 1.0 / b; r.v = a.v / b;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * inv;
    return r;
}
template <int N> GWF_HD Dual<N> operator/(double a, const Dual<N>& b) {
    Dual<N> r; const double inv = 1.0 / b.v; r.v = a / b.v; const double s = -r.v * inv;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = s * b.d[i];
    return r;
}
template <int N> GWF_HD Dual<N>& operator+=(Dual<N>& a, const Dual<N>& b) { a = a + b; return a; }
template <int N> GWF_HD Dual<N>& operator-=(Dual<N>& a, const Dual<N>& b) { a = a - b; return a; }

// ------------------------------------------------------------------ elementary functions
template <int N> GWF_HD Dual<N> chain(const Dual<N>& a, double f, double df) {
    Dual<N> r; r.v = f;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = df * a.d[i];
    return r;
}
GWF_HD double dsqrt(double x) { return sqrt(x); }
// d sqrt(x) = dx / (2 sqrt(x)); where x = 0 AND dx = 0 exactly the tangent is 0 (not inf * 0 = NaN): what jax returns for a constant
// zero under a sqrt (the cotangent never meets the infinite factor), and what the oracle's dual does (oracle/dual.py:_sqrt) --
// e.g. IMRPhenomNSBH at Lambda = 0: compactness 1/2, tidal radius 0 with zero tangent (waveforms.py:3170-3174)
template <int N> GWF_HD Dual<N> dsqrt(const Dual<N>& a) {
    const double s = sqrt(a.v);
    Dual<N> r = chain(a, s, 0.5 / s);
    if (s == 0.0) {
#pragma unroll
        for (int i = 0; i < N; ++i)
            if (a.d[i] == 0.0) r.d[i] = 0.0;
    }
    return r;
}
GWF_HD double dlog(double x) { return log(x); }
template <int N> GWF_HD Dual<N> dlog(const Dual<N>& a) { return chain(a, log(a.v), 1.0 / a.v); }
GWF_HD double dexp(double x) { return exp(x); }
template <int N> GWF_HD Dual<N> dexp(const Dual<N>& a) { const double e = exp(a.v); return chain(a, e, e); }
GWF_HD double datan(double x) { return atan(x); }
template <int N> GWF_HD Dual<N> datan(const Dual<N>& a) { return chain(a, atan(a.v), 1.0 / (1.0 + a.v * a.v)); }
GWF_HD double dpow(double x, double p) { return pow(x, p); }
template <int N> GWF_HD Dual<N> dpow(const Dual<N>& a, double p) {
    const double f = pow(a.v, p);
    return chain(a, f, p * f / a.v);
}
GWF_HD double dfabs(double x) { return fabs(x); }
template <int N> GWF_HD Dual<N> dfabs(const Dual<N>& a) { return a.v < 0.0 ? -a : a; }
// select(cond, a, b): forward-mode semantics of np.where (tangent of the selected branch only)
template <class T> GWF_HD T select(bool c, const T& a, const T& b) { return c ? a : b; }

}  // namespace gwf
