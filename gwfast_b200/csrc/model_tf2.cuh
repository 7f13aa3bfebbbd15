// TaylorF2_RestrictedPN (gwfast/waveforms.py:697-953): 3.5PN phase (+5/6PN tidal, spin-induced quadrupole,
// 3.5PN spin higher-order terms), Newtonian amplitude.  Same coefficient-record / basis-expansion split as
// model_phenomd.cuh, basis in v = (pi M f)^(1/3).  The eccentric extension (waveforms.py:814-845) is not built.
#pragma once
#include "model_common.cuh"

namespace gwf {

constexpr int kTF2 = 11;   // v^-5, v^-3, v^-2, v^-1, 1, log v, v, v log v, v^2, v^5, v^7
template <int NT>
struct TF2Rec {
    double s;
    ScalePow sp;
    double lam[NT];
    double fcut_hz;
    double C, lnC_d[NT];          // A = C f^(-7/6)
    double ph[kTF2][1 + NT];
    TauRec tau;
};

template <int NT> GWF_HD void put_tf2(double* dst, const Dual<NT>& c) {
    dst[0] = c.v;
#pragma unroll
    for (int j = 0; j < NT; ++j) dst[1 + j] = c.d[j];
}

// Kerr-ISCO cut, waveforms.py:922-953 (value only: the grid is not differentiated)
GWF_HD double tf2_r_isco(double chi) {
    const double z1 = 1.0 + cbrt(1.0 - chi * chi) * (cbrt(1.0 + chi) + cbrt(1.0 - chi));
    const double z2 = sqrt(3.0 * chi * chi + z1 * z1);
    const double root = sqrt((3.0 - z1) * (3.0 + z1 + 2.0 * z2));
    return chi > 0. ? 3.0 + z2 - root : 3.0 + z2 + root;
}
GWF_HD double tf2_fcut_kerr(double Mc, double eta, double chi1, double chi2) {
    const double e2 = eta * eta, Mtot = Mc / pow(eta, 3. / 5.), sq = seta_of(eta);
    const double m1 = 0.5 * (1.0 + sq), m2 = 0.5 * (1.0 - sq);
    const double s = (m1 * m1 * chi1 + m2 * m2 * chi2) / (m1 * m1 + m2 * m2);
    const double r21 = m2 / m1;
    const double atot = (chi1 + chi2 * r21 * r21) / ((1. + r21) * (1. + r21));
    const double aeff = atot + 0.41616 * eta * (chi1 + chi2);
    const double r = tf2_r_isco(aeff);
    const double ens = eta * (0.055974469826360077 + 0.5809510763115132 * eta - 0.9606726679372312 * e2 + 3.352411249771192 * e2 * eta);
    const double etot = (ens * (1. + (-0.0030302335878845507 - 2.0066110851351073 * eta + 7.7050567802399215 * e2) * s)) /
                        (1. + (-0.6714403054720589 - 1.4756929437702908 * eta + 7.304676214885011 * e2) * s);
    const double Mfin = Mtot * (1. - etot);
    const double L = 2. / (3. * sqrt(3.)) * (1. + 2. * sqrt(3. * r - 2.));
    const double E = sqrt(1. - 2. / (3. * r));
    const double chif = atot + eta * (L - 2. * atot * (E - 1.)) + (-3.821158961 - 1.2019 * aeff - 1.20764 * aeff * aeff) * e2 +
                        (3.79245 + 1.18385 * aeff + 4.90494 * aeff * aeff) * e2 * eta;
    const double rf = tf2_r_isco(chif);
    const double om = 1. / (rf * sqrt(rf) + chif);
    return om / (kPi * Mfin * kGMsunC3);
}

template <int NT>
GWF_HD void tf2_prologue(TF2Rec<NT>& r, const Intrinsic<NT>& p, double dL, const ModelCfg& cfg, double fcut_host = 0.0) {
    typedef Dual<NT> D;
    const bool tidal = cfg.flags & kFlagTidal;
    const D M = p.Mc / dpow(p.eta, 3. / 5.);
    const D s = M * kGMsunC3;
    r.s = s.v;
    r.sp.set(s.v);
#pragma unroll
    for (int j = 0; j < NT; ++j) r.lam[j] = s.d[j] / s.v;
    r.fcut_hz = fcut_host > 0.0 ? fcut_host
                                : ((cfg.flags & kFlagKerrISCO) ? tf2_fcut_kerr(p.Mc.v, p.eta.v, p.chi1.v, p.chi2.v) : cfg.fcutPar / M.v);   // waveforms.py:918-953
    // amplitude, waveforms.py:875
    const D Cc = sqrt(5. / 24.) * pow(kPi, -2. / 3.) * kClightGpc / dL * dpow(kGMsunC3 * p.Mc, 5. / 6.);
    r.C = Cc.v;
#pragma unroll
    for (int j = 0; j < NT; ++j) r.lnC_d[j] = Cc.d[j] / Cc.v;
    // phase, waveforms.py:784-862
    D q1(1.0), q2(1.0);
    if (tidal && (cfg.flags & kFlagQuadMonTid)) { q1 = quad_mon(p.L1); q2 = quad_mon(p.L2); }
    const PNPhase<D> c = pn_phase_coeffs(p.eta, p.chi1, p.chi2, q1, q2, (cfg.flags & kFlag3p5SpinHO) != 0);
    const D n = 3. / (128. * p.eta);
    D c5 = c.c5;
    double phiR = kPi;
    if (cfg.flags & kFlagPhirefVlso) { c5 = c.c5 * (1. - 3. * log(1. / sqrt(6.))); phiR = 0.; }
    if (cfg.flags & kFlagNewtonian) {
        // NewtInspiral.Phi = 3/4 (8 pi GMsun/c^3 Mc f)^(-5/3) - pi/4 = 3/(128 eta) v^-5 - pi/4 (waveforms.py:226-227)
        for (int k = 0; k < kTF2; ++k) put_tf2(r.ph[k], D(0.0));
        put_tf2(r.ph[0], 0.75 * dpow(8.0 * kGMsunC3 * p.Mc, -5. / 3.) * dpow(s, 5. / 3.));     // times v^-5 = (pi s f)^(-5/3)
        put_tf2(r.ph[4], D(-kPi * 0.25));
        tau_fill_newtonian(r.tau, s, p.Mc);
        return;
    }
    put_tf2(r.ph[0], n);
    put_tf2(r.ph[1], n * c.c2);
    put_tf2(r.ph[2], n * c.c3);
    put_tf2(r.ph[3], n * c.c4);
    put_tf2(r.ph[4], n * c5 + (phiR - kPi * 0.25));
    put_tf2(r.ph[5], n * c.c5 * 3.);
    put_tf2(r.ph[6], n * c.c6);
    put_tf2(r.ph[7], n * (-6848. / 21.));
    put_tf2(r.ph[8], n * c.c7);
    if (tidal) {
        D lt, dl;
        lamt_dellam_from_lam12(p.L1, p.L2, p.eta, lt, dl);      // waveforms.py:853-855
        put_tf2(r.ph[9], n * ((-0.5 * 39.) * lt));
        put_tf2(r.ph[10], n * ((-3115. / 64.) * lt + (6595. / 364.) * seta_of(p.eta) * dl));
    } else {
        put_tf2(r.ph[9], D(0.0));
        put_tf2(r.ph[10], D(0.0));
    }
    tau_fill(r.tau, s, p.eta);
}

// v-powers at x
struct VPow {
    double v, vm1, lv, lpx3;
    GWF_HD void set(const ScalePow& sp, const FreqPoint& fp) {
        const double cp = 1.4645918875615232630201425272637904;      // pi^(1/3)
        const double cpm = 0.68278406325529568146702083315816;       // pi^(-1/3)
        v = cp * sp.s13 * fp.f13;
        vm1 = cpm * sp.sm13 * fp.fm13;
        lpx3 = fma(fp.lnf, 1. / 3., sp.lps3);
        lv = lpx3;          // log v = log(pi x)/3
    }
};

template <int NT>
GWF_HD void tf2_phase(const TF2Rec<NT>& r, const VPow& p, double& phi, double* phi_d) {
    const double v = p.v, v2 = v * v, vm2 = p.vm1 * p.vm1, vm3 = vm2 * p.vm1, vm5 = vm3 * vm2, v5 = v2 * v2 * v, v7 = v5 * v2, vl = v * p.lv;
    const double b[kTF2] = {vm5, vm3, vm2, p.vm1, 1., p.lv, v, vl, v2, v5, v7};
    const double bx[kTF2] = {-5. / 3. * vm5, -vm3, -2. / 3. * vm2, -1. / 3. * p.vm1, 0., 1. / 3., 1. / 3. * v, 1. / 3. * (vl + v), 2. / 3. * v2,
                             5. / 3. * v5, 7. / 3. * v7};
    double val_ = 0., dx = 0., d[NT];
#pragma unroll
    for (int j = 0; j < NT; ++j) d[j] = 0.;
#pragma unroll
    for (int k = 0; k < kTF2; ++k) {
        val_ = fma(r.ph[k][0], b[k], val_);
        dx = fma(r.ph[k][0], bx[k], dx);
#pragma unroll
        for (int j = 0; j < NT; ++j) d[j] = fma(r.ph[k][1 + j], b[k], d[j]);
    }
    phi = val_;
#pragma unroll
    for (int j = 0; j < NT; ++j) phi_d[j] = d[j] + dx * r.lam[j];
}

}  // namespace gwf
