// Constants and f-independent PN / phenomenological building blocks, templated on the scalar type
// (double for values, Dual<N> for values + tangents).  Each block cites the reference lines it reproduces.
#pragma once
#include "dual.cuh"

namespace gwf {

// gwfast/gwfastGlobals.py:38-111 -- copied digit for digit; the constants are not mutually consistent
// (TaylorF2 uses clightGpc, the PhenomD family GMsun_over_c2_Gpc) and parity needs them as written.
constexpr double kGMsunC3 = 4.925491025543575903411922162094833998e-6;   // s
constexpr double kGMsunC2 = 1.476625061404649406193430731479084713e3;    // m
constexpr double kGpc = 3.085677581491367278913937957796471611e25;       // m
constexpr double kGMsunC2Gpc = kGMsunC2 / kGpc;
constexpr double kREarth = 6371.00;                                      // km
constexpr double kClight = 2.99792458e5;                                 // km/s
constexpr double kClightGpc = kClight / 3.0856778570831e+22;
constexpr double kDay = 3600. * 24.;
constexpr double kPi = 3.141592653589793238462643383279502884;
constexpr double kEuler = 0.5772156649015328606065120900824024310;
constexpr double kLn2 = 0.6931471805599453094172321214581765681;

// sqrt(where(eta<0.25, 1-4 eta, 0)), waveforms.py:759.  At eta>=0.25 the tangent is defined as 0 here
// (JAX gives NaN there; catalogs keep eta<0.25).
GWF_HD double seta_of(double eta) { return eta < 0.25 ? sqrt(1.0 - 4.0 * eta) : 0.0; }
template <int N> GWF_HD Dual<N> seta_of(const Dual<N>& eta) {
    if (eta.v < 0.25) return dsqrt(1.0 - 4.0 * eta);
    return Dual<N>(0.0);
}

// ---- tidal deformability maps, gwfastUtils.py:398-448
template <class T> GWF_HD void lamt_dellam_from_lam12(const T& L1, const T& L2, const T& eta, T& lt, T& dl) {
    const T e2 = eta * eta;
    const T s = seta_of(eta);
    lt = (8. / 13.) * ((1. + 7. * eta - 31. * e2) * (L1 + L2) + s * (1. + 9. * eta - 11. * e2) * (L1 - L2));
    dl = 0.5 * (s * (1. - 13272. / 1319. * eta + 8944. / 1319. * e2) * (L1 + L2) +
                (1. - 15910. / 1319. * eta + 32850. / 1319. * e2 + 3380. / 1319. * e2 * eta) * (L1 - L2));
}
template <class T> GWF_HD void lam12_from_lamt_dellam(const T& lt, const T& dl, const T& eta, T& L1, T& L2) {
    const T e2 = eta * eta;
    const T s = seta_of(eta);
    const T a = (8. / 13.) * (1. + 7. * eta - 31. * e2);
    const T b = (8. / 13.) * s * (1. + 9. * eta - 11. * e2);
    const T c = s * (1. - (13272. / 1319.) * eta + (8944. / 1319.) * e2) * 0.5;
    const T d = (1. - (15910. / 1319.) * eta + (32850. / 1319.) * e2 + (3380. / 1319.) * (e2 * eta)) * 0.5;
    const T e4 = e2 * e2;
    const T det = (306656. / 1319.) * (e4 * eta) - (5936. / 1319.) * e4;
    L1 = ((c - d) * lt + (b - a) * dl) / det;
    L2 = ((-d - c) * lt + (b + a) * dl) / det;
}

// ---- spin-induced quadrupole / octupole fits, waveforms.py:779, 1394, 1564
template <class T> GWF_HD T quad_mon(const T& Lam) {
    if (val(Lam) < 1.) return 1. + Lam * (0.427688866723244 + Lam * (-0.324336526985068 + Lam * 0.1107439432180572));
    const T lg = dlog(Lam);
    return dexp(0.1940 + 0.09163 * lg + 0.04812 * lg * lg - 4.283e-3 * lg * lg * lg + 1.245e-4 * lg * lg * lg * lg);
}
template <class T> GWF_HD T oct_mon_minus1(const T& qm) {
    const T lq = dlog(qm);
    return -1. + dexp(0.003131 + 2.071 * lq - 0.7152 * lq * lq + 0.2458 * lq * lq * lq - 0.03309 * lq * lq * lq * lq);
}

// ---- 3.5PN TaylorF2 phasing coefficients (coefficient of v^k), waveforms.py:784-812 and 1059-1077
template <class T>
struct PNPhase {
    T c2, c3, c4, c5, c6, c7, ss6;   // c5l = 3*c5, c6l = -6848/21
};
template <class T>
GWF_HD PNPhase<T> pn_phase_coeffs(const T& eta, const T& chi1, const T& chi2, const T& qm1, const T& qm2, bool spin_ho_3p5) {
    PNPhase<T> c;
    const T e2 = eta * eta;
    const T s = seta_of(eta);
    const T m1 = 0.5 * (1.0 + s), m2 = 0.5 * (1.0 - s);
    const T c12 = chi1 * chi1, c22 = chi2 * chi2, c1c2 = chi1 * chi2;
    const T xs = 0.5 * (chi1 + chi2), xa = 0.5 * (chi1 - chi2);
    const T m1s = m1 * m1, m2s = m2 * m2;
    c.c2 = 3715. / 756. + (55. * eta) / 9.;
    c.c3 = -16. * kPi + (113. * s * xa) / 3. + (113. / 3. - (76. * eta) / 3.) * xs;
    c.c4 = 5. * (3058.673 / 7.056 + 5429. / 7. * eta + 617. * e2) / 72. + 247. / 4.8 * eta * c1c2 - 721. / 4.8 * eta * c1c2 +
           (-720. / 9.6 * qm1 + 1. / 9.6) * m1s * c12 + (-720. / 9.6 * qm2 + 1. / 9.6) * m2s * c22 +
           (240. / 9.6 * qm1 - 7. / 9.6) * m1s * c12 + (240. / 9.6 * qm2 - 7. / 9.6) * m2s * c22;
    const T t5 = (732985. / 2268. - 24260. * eta / 81. - 340. * e2 / 9.) * xs + (732985. / 2268. + 140. * eta / 9.) * s * xa;
    c.c5 = 38645. * kPi / 756. - 65. * kPi * eta / 9. - t5;
    const T k1a = 4703.5 / 8.4 + 2935. / 6. * m1 - 120. * m1s, k1b = -4108.25 / 6.72 - 108.5 / 1.2 * m1 + 125.5 / 3.6 * m1s;
    const T k2a = 4703.5 / 8.4 + 2935. / 6. * m2 - 120. * m2s, k2b = -4108.25 / 6.72 - 108.5 / 1.2 * m2 + 125.5 / 3.6 * m2s;
    const T cross = (326.75 / 1.12 + 557.5 / 1.8 * eta) * eta * c1c2;
    const T ssq = cross + k1a * m1s * qm1 * c12 + k1b * m1s * c12 + k2a * m2s * qm2 * c22 + k2b * m2s * c22;
    c.ss6 = cross + (k1a + k1b) * m1s * c12 + (k2a + k2b) * m2s * c22;
    c.c6 = 11583.231236531 / 4.694215680 - 640. / 3. * kPi * kPi - 684.8 / 2.1 * kEuler +
           eta * (-15737.765635 / 3.048192 + 225.5 / 1.2 * kPi * kPi) + e2 * (76.055 / 1.728) - e2 * eta * (127.825 / 1.296) -
           (2.0 * kLn2) * (684.8 / 2.1) + kPi * chi1 * m1 * (1490. / 3. + m1 * 260.) + kPi * chi2 * m2 * (1490. / 3. + m2 * 260.) + ssq;
    const T base7 = 77096675. * kPi / 254016. + 378515. * kPi * eta / 1512. - 74045. * kPi * e2 / 756.;
    const T lin_s = -25150083775. / 3048192. + 10566655595. * eta / 762048. - 1042165. * e2 / 3024. + 5345. * e2 * eta / 36.;
    const T lin_a = -25150083775. / 3048192. + 26804935. * eta / 6048. - 1985. * e2 / 48.;
    if (spin_ho_3p5) {
        const T xs2 = xs * xs, xa2 = xa * xa;
        c.c7 = base7 + (lin_s + (14585. / 8. - 7270. * eta + 80. * e2) * xa2) * xs +
               (14585. / 24. - 475. * eta / 6. + 100. * e2 / 3.) * xs2 * xs +
               s * (lin_a * xa + (14585. / 24. - 2380. * eta) * xa2 * xa + (14585. / 8. - 215. * eta / 2.) * xa * xs2);
    } else {
        c.c7 = base7 + lin_s * xs + s * (lin_a * xa);
    }
    return c;
}

// ---- 3.5PN time to coalescence, waveforms.py:878-901 (same text at :1299, :1769, :2712), written as
//   tau = sum_k T_k * (pi x)^((k-8)/3)   (+ T_6l * log(16 v^2) v^-2),  x = M GMsun/c^3 f, v = (pi x)^(1/3)
// kTau = 8 basis terms: v^-8, v^-6, v^-5, v^-4, v^-3, v^-2, v^-2*log(16 v^2), v^-1
constexpr int kTau = 8;
template <class T> GWF_HD void tau_coeffs(const T& Ms, const T& eta, T* t) {
    const T e2 = eta * eta;
    const T fac = (5. / 256.) * Ms / eta;
    t[0] = fac;
    t[1] = fac * (743. / 252. + 11. / 3. * eta);
    t[2] = fac * (-32. / 5. * kPi);
    t[3] = fac * (3058673. / 508032. + 5429. / 504. * eta + 617. / 72. * e2);
    t[4] = fac * (-(7729. / 252. - 13. / 3. * eta) * kPi);
    t[5] = fac * (-10052469856691. / 23471078400. + 128. / 3. * kPi * kPi + 6848. / 105. * kEuler +
                  (3147553127. / 3048192. - 451. / 12. * kPi * kPi) * eta - 15211. / 1728. * e2 + 25565. / 1296. * e2 * eta);
    t[6] = fac * (3424. / 105.);
    t[7] = fac * ((-15419335. / 127008. - 75703. / 756. * eta + 14809. / 378. * e2) * kPi);
}

// ---- final spin and radiated energy, waveforms.py:1256-1297
template <class T> GWF_HD T final_spin(const T& eta, const T& chi1, const T& chi2) {
    const T sq = seta_of(eta);
    const T m1 = 0.5 * (1.0 + sq), m2 = 0.5 * (1.0 - sq);
    const T s = m1 * m1 * chi1 + m2 * m2 * chi2;
    const T e2 = eta * eta;
    const T af1 = eta * (3.4641016151377544 - 4.399247300629289 * eta + 9.397292189321194 * e2 - 13.180949901606242 * e2 * eta);
    const T af2 = eta * (s * ((1.0 / eta - 0.0850917821418767 - 5.837029316602263 * eta) + (0.1014665242971878 - 2.0967746996832157 * eta) * s));
    const T af3 = eta * (s * ((-1.3546806617824356 + 4.108962025369336 * eta) * s * s + (-0.8676969352555539 + 2.064046835273906 * eta) * s * s * s));
    return af1 + af2 + af3;
}
template <class T> GWF_HD T radiated_energy(const T& eta, const T& chi1, const T& chi2) {
    const T sq = seta_of(eta);
    const T m1 = 0.5 * (1.0 + sq), m2 = 0.5 * (1.0 - sq);
    const T s = (m1 * m1 * chi1 + m2 * m2 * chi2) / (m1 * m1 + m2 * m2);
    const T e2 = eta * eta;
    const T ens = eta * (0.055974469826360077 + 0.5809510763115132 * eta - 0.9606726679372312 * e2 + 3.352411249771192 * e2 * eta);
    return (ens * (1. + (-0.0030302335878845507 - 2.0066110851351073 * eta + 7.7050567802399215 * e2) * s)) /
           (1. + (-0.6714403054720589 - 1.4756929437702908 * eta + 7.304676214885011 * e2) * s);
}

// ---- QNM ringdown tables: piecewise-linear np.interp (waveforms.py:1034-1035); the tangent is the slope of
// the containing segment, right-continuous at the nodes.  The a-grid is NOT uniform (0.001 steps in the two
// end intervals, 0.002 inside), so the index is found by bisection.
struct QnmTables {
    const double* a;
    const double* fring;
    const double* fdamp;
    int n;
    const double* xitide;   // IMRPhenomNSBH: 200^3 table of xi_tide (model_nsbh.cuh), nullptr until that model runs on the device
};
GWF_HD int qnm_segment(const QnmTables& q, double x) {
    int lo = 0, hi = q.n - 1;               // invariant: a[lo] <= x < a[hi] (after clamping)
    if (x <= q.a[0]) return 0;
    if (x >= q.a[q.n - 1]) return q.n - 2;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (q.a[mid] <= x) lo = mid; else hi = mid;
    }
    return lo;
}
template <class T> GWF_HD T qnm_interp(const QnmTables& q, const double* tab, const T& a) {
    const double x = val(a);
    const int i = qnm_segment(q, x);
    const double slope = (tab[i + 1] - tab[i]) / (q.a[i + 1] - q.a[i]);
    const bool outside = (x < q.a[0]) || (x > q.a[q.n - 1]);
    const double xc = x < q.a[0] ? q.a[0] : (x > q.a[q.n - 1] ? q.a[q.n - 1] : x);
    // value: slope*(x - xp[i]) + fp[i] as numpy computes it; tangent: slope (0 outside the table)
    return (outside ? 0.0 : slope) * (a - xc) + (slope * (xc - q.a[i]) + tab[i]);
}

}  // namespace gwf
