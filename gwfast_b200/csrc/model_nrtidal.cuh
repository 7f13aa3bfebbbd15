// IMRPhenomD_NRTidalv2 (gwfast/waveforms.py:1339-1832) = the IMRPhenomD record of model_phenomd.cuh (with the
// spin-induced quadrupole moments entering the 2PN/3PN phase coefficients) plus
//   * the NRTidalv2 Pade tidal phase (waveforms.py:1543-1558) and the 3.5PN SS/SSS term (:1564-1568),
//   * the tidal amplitude correction (:1677-1687),
//   * the Planck taper with the reference's custom JVP (:1702-1722): the tangent w.r.t. the frequency argument is
//     DROPPED, only the tangent through f_merger is kept.
// The taper is defined here as exactly 0 for x >= 1.2 f_merger: the reference's last grid sample sits on that
// point and comes out as 0 or 1 depending on last-bit rounding.
#pragma once
#include "model_phenomd.cuh"

namespace gwf {

template <int NT>
struct NRTidalRec {
    PhenomDRec<NT> d;
    double fcut_hz;           // 1.2 f_merger / (M GMsun/c^3), waveforms.py:1794-1832
    double sm76;              // s^(-7/6)
    double kph[1 + NT];       // -kappa2T c_Newt/(m1 m2): coefficient of the Pade tidal phase
    double kam[1 + NT];       // -9 kappa2T 2 sqrt(pi/5) / amp0: coefficient of the tidal amplitude
    double ym[1 + NT];        // dimensionless merger frequency of the Planck taper
};

template <class T> GWF_HD T nrt_kappa2T(const T& eta, const T& L1, const T& L2) {   // waveforms.py:1543
    const T sq = seta_of(eta);
    const T Xa = 0.5 * (1.0 + sq), Xb = 0.5 * (1.0 - sq);
    const T Xa2 = Xa * Xa, Xb2 = Xb * Xb;
    return (3.0 / 13.0) * ((1.0 + 12.0 * Xb / Xa) * (Xa2 * Xa2 * Xa) * L1 + (1.0 + 12.0 * Xa / Xb) * (Xb2 * Xb2 * Xb) * L2);
}
template <class T> GWF_HD T nrt_fmerger(const T& eta, const T& k2T) {               // waveforms.py:1690-1700
    const T sq = seta_of(eta);
    const T q = 0.5 * (1.0 + sq - 2.0 * eta) / eta;
    const T num = 1.0 + 3.35411203e-2 * k2T + 4.31460284e-5 * k2T * k2T;
    const T den = 1.0 + 7.54224145e-2 * k2T + 2.23626859e-4 * k2T * k2T;
    return (0.3586 / dsqrt(q)) * (num / den) / (2. * kPi);
}

// parts: 1 = the phase half of the record, 2 = the amplitude half, 3 = both (see PhenomDCore::build)
template <int NT>
GWF_HD void nrtidal_prologue(NRTidalRec<NT>& r, const Intrinsic<NT>& p, double dL, const QnmTables& q, const double* fmin_g, int ngroups,
                             const ModelCfg& cfg, bool lambda_for_fcut, double s_host = 0.0, double fcut_host = 0.0, int parts = 3) {
    typedef Dual<NT> D;
    const D qm1 = quad_mon(p.L1), qm2 = quad_mon(p.L2);            // waveforms.py:1394-1395
    PhenomDCore<NT> c;
    c.build(p.eta, p.chi1, p.chi2, qm1, qm2, q, parts);
    const D M = p.Mc / dpow(p.eta, 3. / 5.);
    ModelCfg cfg_cut = cfg;
    cfg_cut.flags &= ~kFlagNoFcut;                                 // the amplitude always applies the cut (waveforms.py:1673)
    phenomd_fill(r.d, c, M, D(dL), fmin_g, ngroups, cfg, s_host, 0.0, parts);
    const D sq = seta_of(p.eta);
    const D m1 = 0.5 * (1.0 + sq), m2 = 0.5 * (1.0 - sq);
    const D k2T = nrt_kappa2T(p.eta, p.L1, p.L2);
    if (parts & 2) {
        const D amp0 = dsqrt(2.0 * p.eta / 3.0) * pow(kPi, -1. / 6.);
        put(r.kam, (-9.0 * 2. * sqrt(kPi / 5.)) * k2T / amp0);
        const D ym = nrt_fmerger(p.eta, k2T);
        put(r.ym, ym);
        D s = M * kGMsunC3;
        if (s_host > 0.0) s.v = s_host;
        ScalePow sp;
        sp.set(s.v);
        const double sm13 = sp.sm13;
        r.sm76 = sm13 * sm13 * sm13 * sqrt(sm13);
    }
    if (!(parts & 1)) return;
    put(r.kph, -k2T * 2.4375 / (m1 * m2));
    // fcut uses whatever Lambda the events dict carried when fcut() ran (0 if absent: waveforms.py:1809-1812, SURVEY A-20)
    const double k2T_cut = lambda_for_fcut ? k2T.v : 0.0;
    r.fcut_hz = fcut_host > 0.0 ? fcut_host : 1.2 * nrt_fmerger(p.eta.v, k2T_cut) / r.d.s;
    r.d.fcut_hz = r.fcut_hz;
    // 3.5PN spin-squared / spin-cubed terms, waveforms.py:1564-1568: (SS+SSS) * 3/(128 eta) * (pi x)^(2/3)
    const D c12 = p.chi1 * p.chi1, c22 = p.chi2 * p.chi2, m1s = m1 * m1, m2s = m2 * m2;
    const D o1 = oct_mon_minus1(qm1), o2 = oct_mon_minus1(qm2);
    const D ss = -400. * kPi * (qm1 - 1.) * c12 * m1s - 400. * kPi * (qm2 - 1.) * c22 * m2s;
    const D sss = 10. * ((m1s + 308. / 3. * m1) * p.chi1 + (m2s - 89. / 3. * m2) * p.chi2) * (qm1 - 1.) * m1s * c12 +
                  10. * ((m2s + 308. / 3. * m2) * p.chi2 + (m1s - 89. / 3. * m1) * p.chi1) * (qm2 - 1.) * m2s * c22 -
                  440. * o1 * m1s * m1 * c12 * p.chi1 - 440. * o2 * m2s * m2 * c22 * p.chi2;
    const double cp = cbrt(kPi);
    const D c23 = (ss + sss) * c.norm * (cp * cp);
    put(r.d.pins[1], get<NT>(r.d.pins[1]) + c23);
    put(r.d.pint[4], c23);
    put(r.d.pmrd[4], c23);
}

// Pade tidal phase shape R(p) = p^(5/3) N(p)/D(p), p = pi x, and p dR/dp; waveforms.py:1545-1558
GWF_HD void nrt_phase_shape(double p13, double& R, double& xRp) {
    const double p23 = p13 * p13, p = p23 * p13, p43 = p * p13, p53 = p * p23, p2 = p * p;
    const double N = 1.0 + (-12.615214237993088 * p23) + (19.0537346970349 * p) + (-21.166863146081035 * p43) + (90.55082156324926 * p53) +
                     (-60.25357801943598 * p2);
    const double Dn = 1.0 + (-15.11120782773667 * p23) + (22.195327350624694 * p) + (8.064109635305156 * p43);
    const double pN = (2. / 3.) * (-12.615214237993088 * p23) + (19.0537346970349 * p) + (4. / 3.) * (-21.166863146081035 * p43) +
                      (5. / 3.) * (90.55082156324926 * p53) + 2. * (-60.25357801943598 * p2);
    const double pD = (2. / 3.) * (-15.11120782773667 * p23) + (22.195327350624694 * p) + (4. / 3.) * (8.064109635305156 * p43);
    const double iN = rcp_fast(N), iD = rcp_fast(Dn);
    R = p53 * N * iD;
    xRp = R * (5. / 3. + pN * iN - pD * iD);
}

// tidal amplitude shape Q(xt) = xt^3.25 (1 + n1 xt + n289 xt^2.89)/(1 + d xt^4), xt = (pi x)^(2/3), and x dQ/dx; waveforms.py:1682-1687
GWF_HD void nrt_amp_shape(double p13, double lpx3, double& Q, double& xQp) {
    const double xt = p13 * p13, lxt = 2.0 * (lpx3 * 1.0);     // ln xt = (2/3) ln(pi x) = 2 * ln(pi x)/3
    const double x325 = xt * xt * xt * sqrt(sqrt(xt));
    const double x289 = exp(2.89 * lxt);
    const double xt2 = xt * xt, xt4 = xt2 * xt2;
    const double num = 1.0 + 4.157407407407407 * xt + 2519.111111111111 * x289;
    const double den = 1. + 13477.8073677 * xt4;
    const double iden = 1.0 / den;
    Q = x325 * num * iden;
    // xt dQ/dxt = Q (3.25 + (n1 xt + 2.89 n289 xt^2.89)/num - 4 d xt^4/den); x d/dx = (2/3) xt d/dxt
    xQp = (2. / 3.) * Q * (3.25 + (4.157407407407407 * xt + 2.89 * 2519.111111111111 * x289) / num - 4. * 13477.8073677 * xt4 * iden);
}

// Planck taper value and d/dy (y = f_merger); waveforms.py:1703-1717, with T := 0 for x >= 1.2 y (see header)
GWF_HD void nrt_taper(double x, double y, double& T, double& Ty) {
    const double a = 1.2, yp = a * y;
    T = 1.0; Ty = 0.0;
    if (x < y) return;
    if (x >= yp * (1.0 - 1e-12)) { T = 0.0; return; }
    const double u = x - y, v = x - yp, w = yp - y;
    if (u <= 0.0) return;                       // x == y: exp(+inf) -> taper 1, tangent nan_to_num -> 0
    const double ex = w / u + w / v;
    if (ex > 700.) return;                      // e -> inf: T = 1, tangent 0 (nan_to_num of inf/inf)
    const double e = exp(ex);
    const double ie1 = 1.0 / (e + 1.0);
    T = 1.0 - ie1;
    // e/(e+1)^2 = T/(e+1): no overflow of e*(...) or underflow of 1/(e+1)^2 for e up to exp(700) (the reference's inf/inf
    // there is turned into 0 by its nan_to_num, waveforms.py:1717; the true value is < 1e-290)
    Ty = ((a - 1.) / u + (a - 1.) / v + w / (u * u) + 1.2 * w / (v * v)) * T * ie1;
}

}  // namespace gwf
