// IMRPhenomNSBH (gwfast/waveforms.py:2752-3374), B200-style:
//   * phase (waveforms.py:2802-3033) = the IMRPhenomD coefficient record of model_phenomd.cuh built with the NSBH remnant
//     (final spin of arXiv:1903.11622 instead of the BBH fit, QNM tables, spin-induced quadrupole of the NS in the 2PN/3PN terms),
//     the closed-form NRTidalv2 Pade tidal term of model_nrtidal.cuh without taper, and the time shift t0 taken at the LAST sample
//     of the event's frequency grid (waveforms.py:2994-2996) -- one value per grid group, folded out of the record;
//   * amplitude (waveforms.py:3035-3259) = IMRPhenomC-style pieces aPN w- + aPM w- + aRD w+ whose ~20 per-event coefficients
//     (PN series, ringdown Lorentzian, tidal-disruption windows) are computed once per event in dual arithmetic by the prologue;
//     per sample the shape is evaluated in Dual<NT> seeded with d x/d p_j = x lam_j (the amplitude is NOT a linear
//     coefficient/basis expansion: three tanh windows and a Lorentzian whose centres and widths depend on the parameters);
//   * xi_tide (arXiv:1509.00512 eq. (8)): the reference interpolates a 200^3 table in (compactness, q, chi_BH) tri-linearly
//     (waveforms.py:3111, 3286-3373; gwfastUtils.py:1024-1206).  The table is not read from a file here: every node is the largest
//     positive real root of an order-10 polynomial, found on the device (xitide_node: root isolation through the chain of
//     derivatives, bisection, Newton polish) by xitide_table_kernel the first time a device runs the model (64 MB of HBM).
#pragma once
#include "model_nrtidal.cuh"

namespace gwf {

// ------------------------------------------------------------------------------------------------ xi_tide table
constexpr int kXiRes = 200;                     // waveforms.py:2800
constexpr double kXiCompMin = 0.1, kXiCompMax = 0.5, kXiQMin = 1.0, kXiQMax = 100.0, kXiChiMin = -1.0, kXiChiMax = 1.0;   // :3308-3310

// numpy.linspace node: arange(n)*step + start (two roundings, no fused multiply-add), last node = stop exactly
GWF_HD double xi_node_coord(int i, double start, double stop) {
    if (i == kXiRes - 1) return stop;
    const double step = (stop - start) / (kXiRes - 1);
#ifdef __CUDA_ARCH__
    return __dadd_rn(__dmul_rn((double)i, step), start);
#else
    volatile double t = (double)i * step;
    return t + start;
#endif
}

// largest positive real root s of  s^10 - 3 mu s^8 + 2 chi mu^(3/2) s^7 - 3 q s^4 + 6 q mu s^2 - 3 q mu^2 chi^2  (mu = q C), squared
// (waveforms.py:3312-3325: numpy.roots, real positive roots, max of the squares).  The real roots of a polynomial are separated by
// the real roots of its derivative: starting from the linear 9th derivative, the roots of each derivative in (0, B) bracket the
// monotone stretches of the one below, on which a sign change is bisected -- no root can be missed, whatever the complex ones do.
GWF_HD double xitide_node(double Comp, double q, double chi) {
    const double mu = q * Comp;
    double c[11];                                   // ascending powers
    c[10] = 1.0; c[9] = 0.0; c[8] = -3.0 * mu; c[7] = 2.0 * chi * (mu * sqrt(mu)); c[6] = 0.0; c[5] = 0.0; c[4] = -3.0 * q; c[3] = 0.0;
    c[2] = 6.0 * q * mu; c[1] = 0.0; c[0] = -3.0 * q * mu * chi * mu * chi;
    // Fujiwara bound on the moduli of the roots: 2 max_k |c_{n-k}|^(1/k)
    double B = 0.0;
    for (int k = 1; k <= 10; ++k) {
        const double a = fabs(c[10 - k]);
        if (a > 0.0) {
            const double b = pow(a, 1.0 / k);
            B = b > B ? b : B;
        }
    }
    B = 2.0 * B + 1e-3;
    double crit[12], next[12];
    double dbl = 0.0;                               // largest critical point that numpy would report as a (nearly) real double root
    int nc = 0;
    for (int lev = 9; lev >= 0; --lev) {
        // coefficients of the lev-th derivative
        double d[11];
        const int deg = 10 - lev;
        for (int i = 0; i <= deg; ++i) {
            double f = c[i + lev];
            for (int t = 0; t < lev; ++t) f *= (double)(i + lev - t);
            d[i] = f;
        }
        auto ev = [&](double x) {
            double v = d[deg];
            for (int i = deg - 1; i >= 0; --i) v = fma(v, x, d[i]);
            return v;
        };
        if (lev == 0) {
            // numpy.roots returns an (almost) double root as a pair c +- i sqrt(2 p(c)/p''(c)); the reference keeps roots with
            // |imag| < 1e-5 (waveforms.py:3321).  At chi = 1 the polynomial HAS a double root at s^2 = mu, so this is not a corner case
            for (int s = 0; s < nc; ++s) {
                const double x = crit[s], pv = ev(x);
                double p2 = 0.0;
                for (int i = 10; i >= 2; --i) p2 = fma(p2, x, c[i] * (double)(i * (i - 1)));
                if (pv != 0.0 && p2 != 0.0 && pv / p2 > 0.0 && pv / p2 < 0.5e-10) dbl = x;
            }
        }
        int nn = 0;
        double a = 0.0, fa = ev(0.0);
        for (int s = 0; s <= nc; ++s) {
            const double b = s < nc ? crit[s] : B;
            const double fb = ev(b);
            if (fa == 0.0 && a > 0.0) {
                if (nn == 0 || next[nn - 1] != a) next[nn++] = a;
            } else if ((fa < 0.0) != (fb < 0.0) && fb != 0.0) {
                double lo = a, hi = b;
                const bool up = fb > 0.0;
                for (int it = 0; it < 200; ++it) {
                    const double mid = 0.5 * (lo + hi);
                    if (!(mid > lo && mid < hi)) break;
                    const double fm = ev(mid);
                    if ((fm > 0.0) == up && fm != 0.0) hi = mid; else lo = mid;
                }
                next[nn++] = 0.5 * (lo + hi);
            }
            a = b;
            fa = fb;
        }
        if (fa == 0.0) next[nn++] = a;
        nc = nn;
        for (int s = 0; s < nn; ++s) crit[s] = next[s];
    }
    if (nc == 0 || dbl > crit[nc - 1]) return dbl * dbl;
    double s = crit[nc - 1];
    // Newton polish on the polynomial itself
    for (int it = 0; it < 3; ++it) {
        double p = c[10], dp = 0.0;
        for (int i = 9; i >= 0; --i) {
            dp = fma(dp, s, p);
            p = fma(p, s, c[i]);
        }
        if (dp != 0.0) {
            const double sn = s - p / dp;
            if (sn > 0.0 && fabs(sn - s) < 1e-6 * s) s = sn;
        }
    }
    return s * s;
}

// node value: from the device table, or (host emulation, tests) evaluated on the spot
GWF_HD double xitide_value(const QnmTables& t, int i, int j, int k) {
#ifdef __CUDA_ARCH__
    return t.xitide[((size_t)i * kXiRes + j) * kXiRes + k];
#else
    if (t.xitide) return t.xitide[((size_t)i * kXiRes + j) * kXiRes + k];
    return xitide_node(xi_node_coord(i, kXiCompMin, kXiCompMax), xi_node_coord(j, kXiQMin, kXiQMax), xi_node_coord(k, kXiChiMin, kXiChiMax));
#endif
}

// lower node index of the cell: searchsorted(grid, x) - 1 clamped to [0, n-2] (gwfastUtils.py:1196-1198), i.e. grid[i] < x <= grid[i+1]
GWF_HD int xi_cell(double x, double start, double stop) {
    int i = (int)floor((x - start) / ((stop - start) / (kXiRes - 1)));
    i = i < 0 ? 0 : (i > kXiRes - 2 ? kXiRes - 2 : i);
    while (i > 0 && !(xi_node_coord(i, start, stop) < x)) --i;
    while (i < kXiRes - 2 && xi_node_coord(i + 1, start, stop) < x) ++i;
    return i;
}

// tri-linear interpolation with the reference's out-of-bounds behaviour: its `out_of_bounds + x < grid[0]` parses as
// `(out_of_bounds + x) < grid[0]` (gwfastUtils.py:1203-1204), which leaves only chi > 1 flagged (NaN); everything else is
// extrapolated linearly from the edge cell.  Tangents: the gradient of the tri-linear form in the cell, as jax differentiates it.
template <class T> GWF_HD T xitide_interp(const QnmTables& t, const T& Comp, const T& q, const T& chi) {
    const int i0 = xi_cell(val(Comp), kXiCompMin, kXiCompMax), i1 = xi_cell(val(q), kXiQMin, kXiQMax), i2 = xi_cell(val(chi), kXiChiMin, kXiChiMax);
    const double a0 = xi_node_coord(i0, kXiCompMin, kXiCompMax), b0 = xi_node_coord(i0 + 1, kXiCompMin, kXiCompMax);
    const double a1 = xi_node_coord(i1, kXiQMin, kXiQMax), b1 = xi_node_coord(i1 + 1, kXiQMin, kXiQMax);
    const double a2 = xi_node_coord(i2, kXiChiMin, kXiChiMax), b2 = xi_node_coord(i2 + 1, kXiChiMin, kXiChiMax);
    const T y0 = (Comp - a0) / (b0 - a0), y1 = (q - a1) / (b1 - a1), y2 = (chi - a2) / (b2 - a2);
    T acc(0.0);
#pragma unroll
    for (int e = 0; e < 8; ++e) {          // itertools.product order: last index fastest (gwfastUtils.py:1169-1175)
        const int e0 = (e >> 2) & 1, e1 = (e >> 1) & 1, e2 = e & 1;
        const T w = ((e0 ? y0 : 1.0 - y0) * (e1 ? y1 : 1.0 - y1)) * (e2 ? y2 : 1.0 - y2);
        acc = acc + xitide_value(t, i0 + e0, i1 + e1, i2 + e2) * w;
    }
    if (val(chi) > kXiChiMax) acc = acc * nan("");
    return acc;
}

// ------------------------------------------------------------------------------------------------ remnant (shared by phase and amplitude)
template <class T> GWF_HD T nsbh_rem_model(const T& eta, const T& chi1, const T& L, const double* k) {   // arXiv:1903.11622 Tab. I, waveforms.py:2853-2861
    const T p1 = ((k[0] * chi1 + k[1]) + (k[2] * chi1 + k[3]) * eta) * eta;
    const T p2 = ((k[4] * chi1 + k[5]) + (k[6] * chi1 + k[7]) * eta) * eta;
    const T p3 = ((k[8] * chi1 + k[9]) + (k[10] * chi1 + k[11]) * eta) * eta;
    const T den = 1. + L * p3 * p3;
    T m = (1. + L * p1 + L * L * p2) / (den * den);
    if (val(chi1) < 0. && val(eta) < 0.188) m = T(1.0);
    if (val(chi1) < -0.5) m = T(1.0);
    if (val(m) > 1.) m = T(1.0);
    return m;
}
template <class T> GWF_HD void nsbh_masses(const T& eta, T& sq, T& m1, T& m2, T& q) {
    sq = seta_of(eta);
    m1 = 0.5 * (1.0 + sq);
    m2 = 0.5 * (1.0 - sq);
    q = 0.5 * (1.0 + sq - 2.0 * eta) / eta;
}
// final spin of the remnant, waveforms.py:2853-2872 (= 3124-3143)
template <class T> GWF_HD T nsbh_final_spin(const T& eta, const T& chi1, const T& L, T& Shat_out) {
    const double kSp[12] = {-5.44187381e-03, 7.91165608e-03, 2.33362046e-02, 2.47764497e-02, -8.56844797e-07, -2.81727682e-06,
                            6.61290966e-06,  4.28979016e-05, -3.04174272e-02, 2.54889050e-01, 1.47549350e-01, -4.27905832e-01};
    T sq, m1, m2, q;
    nsbh_masses(eta, sq, m1, m2, q);
    const T model = nsbh_rem_model(eta, chi1, L, kSp);
    const T e2 = eta * eta, e3 = e2 * eta;
    const T S1 = chi1 * m1 * m1;
    const T Sh = S1 / (m1 * m1 + m2 * m2);
    Shat_out = Sh;
    // arXiv:1611.00332 eq. (16) as typed at waveforms.py:2868
    const T Lorb = (2. * sqrt(3.) * eta + 5.24 * 3.8326341618708577 * e2 + 1.3 * (-9.487364155598392) * e3) / (1. + 2.88 * 2.5134875145648374 * eta) +
                   ((-0.194) * 1.0009563702914628 * Sh * (4.409160174224525 * eta + 0.5118334706832706 * e2 + (64. - 16. * 4.409160174224525 - 4. * 0.5118334706832706) * e3) +
                    0.0851 * 0.7877509372255369 * Sh * Sh * (8.77367320110712 * eta + (-32.060648277652994) * e2 + (64. - 16. * 8.77367320110712 - 4. * (-32.060648277652994)) * e3) +
                    0.00954 * 0.6540138407185817 * Sh * Sh * Sh * (22.830033250479833 * eta + (-153.83722669033995) * e2 + (64. - 16. * 22.830033250479833 - 4. * (-153.83722669033995)) * e3)) /
                       (1. + (-0.579) * 0.8396665722805308 * Sh * (1.8804718791591157 + (-4.770246856212403) * eta + 0. * e2 + (64. - 64. * 1.8804718791591157 - 16. * (-4.770246856212403) - 4. * 0.) * e3)) +
                   0.3223660562764661 * sq * e2 * (1. + 9.332575956437443 * eta) * chi1 + 2.3170397514509933 * Sh * sq * e3 * (1. + (-3.2624649875884852) * eta) * chi1 +
                   (-0.059808322561702126) * e3 * (chi1 * chi1);
    return (Lorb + S1) * model;
}
// spin-induced quadrupole of the NS as IMRPhenomNSBH types it (waveforms.py:2832; TaylorF2's fit, arXiv:1303.1528)
template <class T> GWF_HD T nsbh_quad_mon(const T& L) {
    if (val(L) < 1e-5) return T(1.0);
    const T lg = dlog(L);
    return dexp(0.194 + 0.0936 * lg + 0.0474 * lg * lg - 0.00421 * lg * lg * lg + 0.000123 * lg * lg * lg * lg);
}

// ------------------------------------------------------------------------------------------------ amplitude coefficients
enum NsbhCoef {
    NB_XDN, NB_XD2, NB_XD3, NB_XD4, NB_XD5, NB_XD6, NB_XD7,   // xdot series, waveforms.py:3062-3069
    NB_AN, NB_A2, NB_A3, NB_A4, NB_A5, NB_A5I, NB_A6,          // time-domain amplitude series, :3071-3079
    NB_GPM,                                                    // gamma_correction * g1 (coefficient of x^(5/6))
    NB_ERD,                                                    // epsilon_tide * del1 (coefficient of the Lorentzian x^(-7/6))
    NB_SIG, NB_FRING,                                          // Lorentzian width and centre
    NB_X0PN, NB_X0PM, NB_X0RD,                                 // window centres (dimensionless frequency)
    NB_DW,                                                     // window width d0 + sigma_tide
    kNsbhCoef
};

template <int N> GWF_HD Dual<N> dtanh(const Dual<N>& a) { const double t = tanh(a.v); return chain(a, t, 1.0 - t * t); }
GWF_HD double dtanh(double a) { return tanh(a); }

// all f-independent amplitude quantities, waveforms.py:3047-3207.  s = M GMsun/c^3 only scales the window centres back and
// forth in the reference (f0_tilde = x0/s, then f0_tilde*s): they are kept dimensionless here.
template <class T>
GWF_HD void nsbh_amp_coeffs(T* c, const T& eta, const T& chi1, const T& chi2, const T& L, const QnmTables& tab) {
    const T e2 = eta * eta;
    T sq, m1, m2, q;
    nsbh_masses(eta, sq, m1, m2, q);
    const T chieff = m1 * chi1 + m2 * chi2, chisum = 2. * chieff, chiprod = chieff * chieff;
    // IMRPhenomC SPA coefficients, LALSimIMRPhenomC_internals.c as typed at waveforms.py:3062-3079
    c[NB_XDN] = 64. * eta / 5.;
    c[NB_XD2] = -7.43 / 3.36 - 11. * eta / 4.;
    c[NB_XD3] = 4. * kPi - 11.3 * chieff / 1.2 + 19. * eta * chisum / 6.;
    c[NB_XD4] = 3.4103 / 1.8144 + 5. * chiprod + eta * (13.661 / 2.016 - chiprod / 8.) + 5.9 * e2 / 1.8;
    c[NB_XD5] = -kPi * (41.59 / 6.72 + 189. * eta / 8.) - chieff * (31.571 / 1.008 - 116.5 * eta / 2.4) + chisum * (21.863 * eta / 1.008 - 79. * e2 / 6.) -
                3. * chieff * chiprod / 4. + 9. * eta * chieff * chiprod / 4.;
    c[NB_XD6] = 164.47322263 / 1.39708800 - 17.12 * kEuler / 1.05 + 16. * kPi * kPi / 3. - 8.56 * (4.0 * kLn2) / 1.05 + eta * (45.1 * kPi * kPi / 4.8 - 561.98689 / 2.17728) +
                5.41 * e2 / 8.96 - 5.605 * eta * e2 / 2.592 - 80. * kPi * chieff / 3. + eta * chisum * (20. * kPi / 3. - 113.5 * chieff / 3.6) +
                chiprod * (64.153 / 1.008 - 45.7 * eta / 3.6) - chiprod * (7.87 * eta / 1.44 - 30.37 * e2 / 1.44);
    c[NB_XD7] = -kPi * (4.415 / 4.032 - 358.675 * eta / 6.048 - 91.495 * e2 / 1.512) - chieff * (252.9407 / 2.7216 - 845.827 * eta / 6.048 + 415.51 * e2 / 8.64) +
                chisum * (158.0239 * eta / 5.4432 - 451.597 * e2 / 6.048 + 20.45 * e2 * eta / 4.32 + 107. * eta * chiprod / 6. - 5. * e2 * chiprod / 24.) +
                12. * kPi * chiprod - chiprod * chieff * (150.5 / 2.4 + eta / 8.) + chieff * chiprod * (10.1 * eta / 2.4 + 3. * e2 / 8.);
    c[NB_AN] = 8. * eta * sqrt(kPi / 5.);
    c[NB_A2] = (-107. + 55. * eta) / 42.;
    c[NB_A3] = 2. * kPi - 4. * chieff / 3. + 2. * eta * chisum / 3.;
    c[NB_A4] = -2.173 / 1.512 - eta * (10.69 / 2.16 - 2. * chiprod) + 2.047 * e2 / 1.512;
    c[NB_A5] = -10.7 * kPi / 2.1 + eta * (3.4 * kPi / 2.1);
    c[NB_A5I] = -24. * eta;
    c[NB_A6] = 270.27409 / 6.46800 - 8.56 * kEuler / 1.05 + 2. * kPi * kPi / 3. + eta * (4.1 * kPi * kPi / 9.6 - 27.8185 / 3.3264) - 20.261 * e2 / 2.772 +
               11.4635 * eta * e2 / 9.9792 - 4.28 * (4.0 * kLn2) / 1.05;
    // waveforms.py:3081-3091
    T g1 = 4.149e+00 * chieff + -4.070e+00 * chiprod + -8.752e+01 * eta * chieff + -4.897e+01 * eta + 6.665e+02 * e2;
    if (val(g1) < 0.) g1 = T(0.0);
    T del1 = -5.472e-02 * chieff + 2.094e-02 * chiprod + 3.554e-01 * eta * chieff + 1.151e-01 * eta + 9.640e-01 * e2;
    T del2 = -1.235e+00 * chieff + 3.423e-01 * chiprod + 6.062e+00 * eta * chieff + 5.949e+00 * eta + -1.069e+01 * e2;
    if (val(del1) < 0.) del1 = T(0.0);
    if (val(del2) < 1.0e-4) del2 = T(1.0e-4);
    const double d0 = 0.015;
    // NS compactness, arXiv:1608.02582 eq. (78); waveforms.py:3097-3103
    const double a0C = 0.360, a1C = -0.0355, a2C = 0.000705;
    T Comp;
    if (val(L) > 1.) {
        const T lg = dlog(L);
        Comp = a0C + a1C * lg + a2C * lg * lg;
    } else {
        Comp = 0.5 + (3. * a0C - a1C - 1.5) * L * L + (-2. * a0C + a1C + 1.) * L * L * L;
    }
    const T xiT = xitide_interp(tab, Comp, q, chi1);
    // Kerr ISCO, waveforms.py:3113-3116
    const T c1s = chi1 * chi1;
    const T Z1 = 1.0 + dpow(1.0 - c1s, 1. / 3.) * (dpow(1.0 + chi1, 1. / 3.) + dpow(1.0 - chi1, 1. / 3.));
    const T Z2 = dsqrt(3.0 * c1s + Z1 * Z1);
    const T rr = dsqrt((3.0 - Z1) * (3.0 + Z1 + 2.0 * Z2));
    const T rISCO = val(chi1) > 0. ? 3.0 + Z2 - rr : 3.0 + Z2 + rr;
    const T tmpM = 0.296 * xiT * (1.0 - 2.0 * Comp) - 0.171 * q * Comp * rISCO;
    const T Mtorus = val(tmpM) > 0. ? tmpM : T(0.0);
    T Shat;
    const T chif = nsbh_final_spin(eta, chi1, L, Shat);
    // remnant mass, waveforms.py:3145-3159
    const double kM[12] = {-1.83417425e-03, 2.39226041e-03, 4.29407902e-03, 9.79775571e-03, 2.33868869e-07,  -8.28090025e-07,
                           -1.64315549e-06, 8.08340931e-06, -2.00726981e-02, 1.31986011e-01, 6.50754064e-02, -1.42749961e-01};
    const T modelM = nsbh_rem_model(eta, chi1, L, kM);
    const T e3 = e2 * eta;
    const T Erad = (((1. + -2.0 / 3.0 * sqrt(2.)) * eta + 0.5609904135313374 * e2 + (-0.84667563764404) * e3 + 3.145145224278187 * e2 * e2) *
                    (1. + 0.346 * (-0.2091189048177395) * Shat * (1.8083565298668276 + 15.738082204419655 * eta + (16. - 16. * 1.8083565298668276 - 4. * 15.738082204419655) * e2) +
                     0.211 * (-0.19709136361080587) * Shat * Shat * (4.271313308472851 + 0. * eta + (16. - 16. * 4.271313308472851 - 4. * 0.) * e2) +
                     0.128 * (-0.1588185739358418) * Shat * Shat * Shat * (31.08987570280556 + (-243.6299258830685) * eta + (16. - 16. * 31.08987570280556 - 4. * (-243.6299258830685)) * e2))) /
                       (1. + (-0.212) * 2.9852925538232014 * Shat * (1.5673498395263061 + (-0.5808669012986468) * eta + (16. - 16. * 1.5673498395263061 - 4. * (-0.5808669012986468)) * e2)) +
                   (-0.09803730445895877) * sq * e2 * (1. + (-3.2283713377939134) * eta) * chi1 + (-0.01978238971523653) * Shat * sq * eta * (1. + (-4.91667749015812) * eta) * chi1 +
                   0.01118530335431078 * e3 * c1s;
    const T finalMass = (1. - Erad) * modelM;
    // (2,2) quasi-normal mode of the remnant, waveforms.py:3161-3165: omega = 1 + k (c1 + c2 k + c3 k^2 + c4 k^3 + c5 k^4), c_n = r_n exp(i phi_n)
    const T kap = dsqrt(dlog(2. - chif) / log(3.));
    const double rr_[5] = {1.5578, 1.9510, 2.0997, 1.4109, 0.4106}, ph_[5] = {2.9031, 5.9210, 2.7606, 5.9143, 2.7952};
    T wre(0.0), wim(0.0), kp(1.0);
    for (int n = 0; n < 5; ++n) {
        wre = wre + (rr_[n] * cos(ph_[n])) * kp;
        wim = wim + (rr_[n] * sin(ph_[n])) * kp;
        kp = kp * kap;
    }
    wre = 1.0 + kap * wre;
    wim = kap * wim;
    const T fring = 0.5 * wre / kPi / finalMass;
    const T rtide = xiT * (1.0 - 2.0 * Comp) / (q * Comp);
    const T qfac = 0.5 * wre / wim;
    const T ftide = dfabs(1.0 / (kPi * (chi1 + dsqrt(rtide * rtide * rtide))) * (1.0 + 1.0 / q));
    const T frt = 0.99 * 0.98 * fring;
    const bool lam_big = val(L) > 1.0, disrupt = val(ftide) < val(fring), torus = val(Mtorus) > 0.;
    const T gamma_c = lam_big ? T(1.25) : 1.0 + 0.5 * L - 0.25 * L * L;
    const T fr1 = ftide / frt - 1.;
    const T del2p = lam_big ? 1.62496 * 0.25 * (1. + dtanh(4.0 * (fr1 - 0.0188092) / 0.338737)) : del2 - 2. * (del2 - 0.81248) * L + (del2 - 0.81248) * L * L;
    const T sigma = del2p * fring / qfac;
    const T seta_ = dsqrt(eta);
    // merger type, waveforms.py:3180-3190
    const T eps_tide = disrupt ? T(0.0) : 2. * 0.25 * (1. + dtanh(4.0 * ((fr1 * fr1 - 0.571505 * Comp - 0.00508451 * chi1) + 0.0796251) / 0.0801192));
    const T eins_raw = 1.29971 - 1.61724 * (Mtorus + 0.424912 * Comp + 0.363604 * seta_ - 0.0605591 * chi1);
    const T eps_ins = disrupt ? (val(eins_raw) > 1. ? T(1.0) : eins_raw) : (torus ? eins_raw : T(1.0));
    const T st_poly = 0.137722 - 0.293237 * (Mtorus - 0.132754 * Comp + 0.576669 * seta_ - 0.0603749 * chi1 - 0.0601185 * c1s - 0.0729134 * c1s * chi1);
    const T st_tanh = 0.5 * (1. - dtanh(4.0 * ((fr1 * fr1 - 0.657424 * Comp - 0.0259977 * chi1) + 0.206465) / 0.226844));
    const T sig_tide = disrupt ? (torus ? st_poly : 0.5 * (st_poly + st_tanh)) : st_tanh;
    const T x_nd = lam_big ? frt : (1.0 - 0.02 * L + 0.01 * L * L) * 0.98 * fring;       // non-disruptive centre
    const T x0PN = disrupt ? (torus ? ftide : (1.0 - 1.0 / q) * frt + eps_ins * ftide / q) : x_nd;
    const T x0PM = disrupt ? (torus ? ftide : (1.0 - 1.0 / q) * frt + ftide / q) : x_nd;
    const T x0RD = disrupt ? T(0.0) : x_nd;
    c[NB_GPM] = gamma_c * g1;
    c[NB_ERD] = eps_tide * del1;
    c[NB_SIG] = sigma;
    c[NB_FRING] = fring;
    c[NB_X0PN] = eps_ins * x0PN;
    c[NB_X0PM] = x0PM;
    c[NB_X0RD] = x0RD;
    c[NB_DW] = d0 + sig_tide;
}

// the frequency-dependent quantities a sample needs, each with its logarithmic x-derivative (x d/dx)
struct NsbhPowers {
    double x, v, lnv, x56, xm76;      // x, (pi x)^(1/3), ln v, x^(5/6), x^(-7/6)
};
template <class T> struct NsbhLift;
template <> struct NsbhLift<double> {
    const double* lam;
    GWF_HD double operator()(double v, double) const { return v; }
};
template <int N> struct NsbhLift<Dual<N>> {
    const double* lam;
    GWF_HD Dual<N> operator()(double v, double xdv) const {
        Dual<N> r; r.v = v;
#pragma unroll
        for (int j = 0; j < N; ++j) r.d[j] = xdv * lam[j];
        return r;
    }
};

// amplitudeIMR(x), waveforms.py:3236-3253
template <class T>
GWF_HD T nsbh_amp_shape(const T* c, const NsbhPowers& p, const double* lam) {
    NsbhLift<T> lift{lam};
    const T x = lift(p.x, p.x), v = lift(p.v, p.v * (1. / 3.)), lnv = lift(p.lnv, 1. / 3.), px = kPi * x;
    const T v2 = v * v, px2 = px * px, v5 = v2 * v2 * v;
    const T xdot = c[NB_XDN] * (v5 * v5) *
                   (1. + c[NB_XD2] * v2 + c[NB_XD3] * px + c[NB_XD4] * px * v + c[NB_XD5] * v2 * px + (c[NB_XD6] + (-856. / 105.) * 2. * lnv) * px2 + c[NB_XD7] * v * px2);
    const T ampfac = dsqrt(dfabs(kPi / (1.5 * v * xdot)));
    const T pre = ampfac * c[NB_AN] * v2;
    const T re = pre * (1. + c[NB_A2] * v2 + c[NB_A3] * px + c[NB_A4] * v * px + c[NB_A5] * v2 * px + (c[NB_A6] + (-428. / 105.) * 2. * lnv) * px2);
    const T im = pre * (c[NB_A5I] * v2 * px + (4.28 * kPi / 1.05) * px2);
    const T aPN = dsqrt(re * re + im * im);
    const T aPM = c[NB_GPM] * lift(p.x56, p.x56 * (5. / 6.));
    const T u = x - c[NB_FRING], s2 = c[NB_SIG] * c[NB_SIG];
    const T aRD = c[NB_ERD] * (s2 / (u * u + s2 * 0.25)) * lift(p.xm76, p.xm76 * (-7. / 6.));
    const T iw = 4. / c[NB_DW];
    const T wPN = 0.5 * (1. - dtanh((x - c[NB_X0PN]) * iw));
    const T wPM = 0.5 * (1. - dtanh((x - c[NB_X0PM]) * iw));
    const T wRD = 0.5 * (1. + dtanh((x - c[NB_X0RD]) * iw));
    return aPN * wPN + aPM * wPM + aRD * wRD;
}

// amplitudeIMR(x) and its tangents without dual arithmetic: the gradient of the shape with respect to its 22 coefficients and to
// ln x is written out by hand (reverse mode), then contracted with the coefficient tangents of the record:
//   dF_j = sum_k dF/dc_k * dc_k/dp_j + (x dF/dx) lam_j        (about 300 FP64 operations per sample instead of ~2000 in Dual<6>)
template <int NT>
GWF_HD void nsbh_amp_grad(const double (*c)[1 + NT], const NsbhPowers& p, const double* lam, double& F, double* dF) {
    const double x = p.x, v = p.v, lnv = p.lnv, px = kPi * x;
    const double v2 = v * v, px2 = px * px, v5 = v2 * v2 * v;
    const double k6 = (-856. / 105.) * 2., k6a = (-428. / 105.) * 2.;
    // xdot series S1 and amplitude series S2 + i T2 on their bases, with x d/dx of every term
    const double b1[6] = {v2, px, px * v, v2 * px, px2, v * px2};
    const double e1[6] = {2. / 3., 1., 4. / 3., 5. / 3., 2., 7. / 3.};
    double S1 = fma(k6 * lnv, px2, 1.0), xS1 = k6 * px2 * fma(2.0, lnv, 1. / 3.);
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const double t = c[NB_XD2 + k][0] * b1[k];
        S1 += t;
        xS1 = fma(e1[k], t, xS1);
    }
    const double b2[5] = {v2, px, v * px, v2 * px, px2};
    const double e2[5] = {2. / 3., 1., 4. / 3., 5. / 3., 2.};
    const int i2[5] = {NB_A2, NB_A3, NB_A4, NB_A5, NB_A6};
    double S2 = fma(k6a * lnv, px2, 1.0), xS2 = k6a * px2 * fma(2.0, lnv, 1. / 3.);
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const double t = c[i2[k]][0] * b2[k];
        S2 += t;
        xS2 = fma(e2[k], t, xS2);
    }
    const double A6I = 4.28 * kPi / 1.05;
    const double tI = v2 * px;
    const double T2 = fma(c[NB_A5I][0], tI, A6I * px2), xT2 = fma(5. / 3. * c[NB_A5I][0], tI, 2. * A6I * px2);
    const double R2 = S2 * S2 + T2 * T2;
    const double xdot = c[NB_XDN][0] * (v5 * v5) * S1;
    const double ampfac = sqrt(fabs(kPi / (1.5 * v * xdot)));
    const double aPN = ampfac * c[NB_AN][0] * v2 * sqrt(R2);
    const double iS1 = 1.0 / S1, iR2 = 1.0 / R2;
    // windows
    const double idw = 1.0 / c[NB_DW][0], iw = 4.0 * idw;
    const double zPN = (x - c[NB_X0PN][0]) * iw, zPM = (x - c[NB_X0PM][0]) * iw, zRD = (x - c[NB_X0RD][0]) * iw;
    const double tPN = tanh(zPN), tPM = tanh(zPM), tRD = tanh(zRD);
    const double wPN = 0.5 * (1. - tPN), wPM = 0.5 * (1. - tPM), wRD = 0.5 * (1. + tRD);
    const double sPN = -0.5 * (1. - tPN * tPN), sPM = -0.5 * (1. - tPM * tPM), sRD = 0.5 * (1. - tRD * tRD);     // dw/dz
    const double aPM = c[NB_GPM][0] * p.x56;
    const double u = x - c[NB_FRING][0], sg = c[NB_SIG][0], s2 = sg * sg;
    const double D = fma(u, u, 0.25 * s2), iD = 1.0 / D;
    const double LRD = s2 * iD;
    const double aRD = c[NB_ERD][0] * LRD * p.xm76;
    F = aPN * wPN + aPM * wPM + aRD * wRD;
    // gradient, contracted with the coefficient tangents term by term (nothing but the NT sums stays live)
    const double qPN = aPN * wPN;                                   // every ln-derivative of aPN is multiplied by this
    const double hPN = aPN * sPN, hPM = aPM * sPM, hRD = aRD * sRD; // window derivatives d/dz
    const double rd = c[NB_ERD][0] * p.xm76 * wRD * 2. * sg * iD * iD;       // d aRD wRD / d(sig, fring) share this factor
    const double gx = qPN * (-0.5 * xS1 * iS1 - 7. / 6. + (S2 * xS2 + T2 * xT2) * iR2) + (5. / 6.) * aPM * wPM +
                      (-rd * sg * u * x - (7. / 6.) * aRD * wRD) + (hPN + hPM + hRD) * x * iw;
#pragma unroll
    for (int j = 0; j < NT; ++j) dF[j] = gx * lam[j];
    auto add = [&](int k, double gk) {
#pragma unroll
        for (int j = 0; j < NT; ++j) dF[j] = fma(gk, c[k][1 + j], dF[j]);
    };
    add(NB_XDN, -0.5 * qPN / c[NB_XDN][0]);
    const double q1 = -0.5 * qPN * iS1;
#pragma unroll
    for (int k = 0; k < 6; ++k) add(NB_XD2 + k, q1 * b1[k]);
    add(NB_AN, qPN / c[NB_AN][0]);
    const double qS = qPN * S2 * iR2, qT = qPN * T2 * iR2;
#pragma unroll
    for (int k = 0; k < 5; ++k) add(i2[k], qS * b2[k]);
    add(NB_A5I, qT * tI);
    add(NB_GPM, p.x56 * wPM);
    add(NB_ERD, LRD * p.xm76 * wRD);
    add(NB_SIG, rd * u * u);
    add(NB_FRING, rd * sg * u);
    add(NB_X0PN, -hPN * iw);
    add(NB_X0PM, -hPM * iw);
    add(NB_X0RD, -hRD * iw);
    add(NB_DW, -(hPN * zPN + hPM * zPM + hRD * zRD) * idw);
}

// ------------------------------------------------------------------------------------------------ record
template <int NT>
struct NSBHRec {
    PhenomDRec<NT, false> d;             // phase regions only (t0 NOT folded in), s, lam, tau, fcut; no IMRPhenomD amplitude rows
    double fcut_hz;                      // 0.2 / s, waveforms.py:3284
    double C, lnC_d[NT];                 // 2 sqrt(5/(64 pi)) M^2 GMsun_c2_Gpc GMsun_c3 / dL, waveforms.py:3256
    double sm76;                         // s^(-7/6)
    double kph[1 + NT];                  // -kappa2T c_Newt/(m1 m2), waveforms.py:3006-3022
    double t0[kMaxGroups][1 + NT];       // d Phi_MRD/dx at the last sample of the group's grid, waveforms.py:2994-2996
    double amp[kNsbhCoef][1 + NT];
};

// fmax_g[g]: upper end of group g's frequency band in Hz (0 = none): the grid of an event ends at min(fcut, fmax_g) (signal.py:717-718);
// fmax_exact: fmax_g[0] IS the last sample (stand-alone calls on a user grid)
template <int NT>
GWF_HD void nsbh_prologue(NSBHRec<NT>& r, const Intrinsic<NT>& p, double dL, const QnmTables& q, const double* fmin_g, const double* fmax_g, bool fmax_exact,
                          int ngroups, const ModelCfg& cfg, double s_host = 0.0, double fcut_host = 0.0) {
    typedef Dual<NT> D;
    const D L = p.L2;                                      // Lambda1 is discarded (waveforms.py:2826-2829)
    const D qm2 = nsbh_quad_mon(L);
    D Shat;
    const D chif = nsbh_final_spin(p.eta, p.chi1, L, Shat);
    const D erad = radiated_energy(p.eta, p.chi1, p.chi2);  // waveforms.py:2874
    // np.interp(chif.real, ...) (waveforms.py:2876-2877)
    const D fring = qnm_interp(q, q.fring, chif) / (1.0 - erad), fdamp = qnm_interp(q, q.fdamp, chif) / (1.0 - erad);
    PhenomDCore<NT> c;
    c.build_rd(p.eta, p.chi1, p.chi2, D(1.0), qm2, fring, fdamp, 1);
    c.t0 = D(0.0);
    const D M = p.Mc / dpow(p.eta, 3. / 5.);
    phenomd_fill(r.d, c, M, D(dL), fmin_g, ngroups, cfg, s_host, fcut_host, 1);
    D s = M * kGMsunC3;
    if (s_host > 0.0) s.v = s_host;
    tau_fill(r.d.tau, s, p.eta);
    r.fcut_hz = r.d.fcut_hz;
    // t0 per group and the constant of the group: t0 xref - phiRef + pi  (waveforms.py:3029-3033)
    const bool apply_cut = !(cfg.flags & kFlagNoFcut);
    for (int g = 0; g < ngroups; ++g) {
        double fmax = r.fcut_hz;
        if (fmax_g && fmax_g[g] > 0.0 && (fmax_exact || fmax_g[g] < fmax)) fmax = fmax_g[g];
        const D xpk = s * fmax;
        const D t0 = c.dphi_mrd(xpk);
        put(r.t0[g], t0);
        const D xref = (cfg.flags & kFlagHasFRef) ? s * cfg.fRef : s * fmin_g[g];
        put(r.d.pc[g], t0 * xref - c.phi_regions(xref, apply_cut) + kPi);
    }
    // tidal phase coefficient, waveforms.py:3006, 3022 (Lambda of the NS only)
    D sq, m1, m2, qq;
    nsbh_masses(p.eta, sq, m1, m2, qq);
    const D m2s = m2 * m2;
    const D k2T = (3.0 / 13.0) * ((1.0 + 12.0 * m1 / m2) * (m2s * m2s * m2) * L);
    put(r.kph, -k2T * 2.4375 / (m1 * m2));
    // amplitude
    const D Cc = 2. * sqrt(5. / (64. * kPi)) * M * kGMsunC2Gpc * M * kGMsunC3 / D(dL);
    r.C = Cc.v;
#pragma unroll
    for (int j = 0; j < NT; ++j) r.lnC_d[j] = Cc.d[j] / Cc.v;
    const double sm13 = r.d.sp.sm13;
    r.sm76 = sm13 * sm13 * sm13 * sqrt(sm13);
    D ac[kNsbhCoef];
    nsbh_amp_coeffs(ac, p.eta, p.chi1, p.chi2, L, q);
    for (int k = 0; k < kNsbhCoef; ++k) put(r.amp[k], ac[k]);
}

GWF_HD void nsbh_powers(NsbhPowers& np, const XPow& p, double sm76) {
    np.x = p.x;
    np.v = 1.4645918875615232630201425272637904 * p.x13;
    np.lnv = p.lpx3;
    np.x56 = p.x * sqrt(p.xm13);
    np.xm76 = sm76 * p.fm76;
}

}  // namespace gwf
