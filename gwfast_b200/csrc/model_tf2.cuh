// TaylorF2_RestrictedPN (gwfast/waveforms.py:697-953): 3.5PN phase (+5/6PN tidal, spin-induced quadrupole,
// 3.5PN spin higher-order terms), Newtonian amplitude.  Same coefficient-record / basis-expansion split as
// model_phenomd.cuh, basis in v = (pi M f)^(1/3).  The eccentric extension (waveforms.py:814-845) is not built.
#pragma once
#include "model_common.cuh"

namespace gwf {

constexpr int kTF2 = 11;   // v^-5, v^-3, v^-2, v^-1, 1, log v, v, v log v, v^2, v^5, v^7
constexpr int kTF2Ecc = 7; // eccentric phase: v^(-34/3) * {1, v^2, v^3, v^4, v^5, v^6, v^6 log v}
template <int NT>
struct TF2Rec {
    double s;
    ScalePow sp;
    double lam[NT];
    double fcut_hz;
    double C, lnC_d[NT];          // A = C f^(-7/6)
    double ph[kTF2][1 + NT];
    TauRec tau;
    int has_ecc, pad_;
    double ec[kMaxGroups][kTF2Ecc][1 + NT];   // eccentric phase coefficients per grid group (v0ecc = v at the group's fmin unless fRef_ecc is given)
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

// eccentric phase, low-eccentricity limit to 3PN (arXiv:1605.00304; waveforms.py:814-845): coefficients of the powers of v with the
// reference velocity v0 folded in; everything multiplied by 3/(128 eta) * (-2355/1462) e0^2 v0^(19/3)
template <int NT>
GWF_HD void tf2_ecc_coeffs(double (*ec)[1 + NT], const Dual<NT>& eta, const Dual<NT>& ecc, const Dual<NT>& v0) {
    typedef Dual<NT> D;
    const D e2 = eta * eta, e3 = e2 * eta;
    const D v02 = v0 * v0, v03 = v02 * v0, v04 = v02 * v02, v05 = v04 * v0, v06 = v03 * v03;
    const D twoV = 29.9076223 / 8.1976608 + 18.766963 / 2.927736 * eta, twoV0 = 2.833 / 1.008 - 19.7 / 3.6 * eta;
    const double threeV = -28.19123 / 2.82600 * kPi, threeV0 = 37.7 / 7.2 * kPi;
    const D fourV4 = 16.237683263 / 3.330429696 + 241.33060753 / 9.71375328 * eta + 156.2608261 / 6.9383952 * e2;
    const D fourV2V02 = 84.7282939759 / 8.2632420864 - 7.18901219 / 3.68894736 * eta - 36.97091711 / 1.05398496 * e2;
    const D fourV04 = -1.193251 / 3.048192 - 66.317 / 9.072 * eta + 18.155 / 1.296 * e2;
    const D fiveV5 = -28.31492681 / 1.18395270 * kPi - 115.52066831 / 2.70617760 * kPi * eta;
    const D fiveV3V02 = -79.86575459 / 2.84860800 * kPi + 55.5367231 / 1.0173600 * kPi * eta;
    const D fiveV2V03 = 112.751736071 / 5.902315776 * kPi + 70.75145051 / 2.10796992 * kPi * eta;
    const D fiveV05 = 76.4881 / 9.0720 * kPi - 94.9457 / 2.2680 * kPi * eta;
    const D sixV6 = -436.03153867072577087 / 1.32658535116800000 + 53.6803271 / 1.9782000 * kEuler + 157.22503703 / 3.25555200 * kPi * kPi +
                    (2991.72861614477 / 6.89135247360 - 15.075413 / 1.446912 * kPi * kPi) * eta + 345.5209264991 / 4.1019955200 * e2 +
                    506.12671711 / 8.78999040 * e3 + 384.3505163 / 5.9346000 * log(2.) - 112.1397129 / 1.7584000 * log(3.);
    const D sixV4V02 = 46.001356684079 / 3.357073133568 + 253.471410141755 / 5.874877983744 * eta - 169.3852244423 / 2.3313007872 * e2 -
                       307.833827417 / 2.497822272 * e3;
    const double sixV3V03 = -106.2809371 / 2.0347200 * kPi * kPi;
    const D sixV2V04 = -3.56873002170973 / 2.49880440692736 - 260.399751935005 / 8.924301453312 * eta + 15.0484695827 / 3.5413894656 * e2 +
                       340.714213265 / 3.794345856 * e3;
    const D sixV06 = 265.31900578691 / 1.68991764480 - 33.17 / 1.26 * kEuler + 12.2833 / 1.0368 * kPi * kPi +
                     (91.55185261 / 5.48674560 - 3.977 / 1.152 * kPi * kPi) * eta - 5.732473 / 1.306368 * e2 - 30.90307 / 1.39968 * e3 +
                     87.419 / 1.890 * log(2.) - 260.01 / 5.60 * log(3.);
    const double kL = 53.6803271 / 3.9564000;
    const D pre = (3. / (128. * eta)) * (-2.355 / 1.462) * ecc * ecc * dpow(v0, 19. / 3.);
    put_tf2(ec[0], pre * (1. + twoV0 * v02 + threeV0 * v03 + fourV04 * v04 + fiveV05 * v05 + (sixV06 - 33.17 / 2.52 * dlog(16. * v02)) * v06));
    put_tf2(ec[1], pre * (twoV + fourV2V02 * v02 + fiveV2V03 * v03 + sixV2V04 * v04));
    put_tf2(ec[2], pre * (threeV + fiveV3V02 * v02 + sixV3V03 * v03));
    put_tf2(ec[3], pre * (fourV4 + sixV4V02 * v02));
    put_tf2(ec[4], pre * fiveV5);
    put_tf2(ec[5], pre * (sixV6 + kL * log(16.)));
    put_tf2(ec[6], pre * (2. * kL));
}

template <int NT>
GWF_HD void tf2_prologue(TF2Rec<NT>& r, const Intrinsic<NT>& p, double dL, const ModelCfg& cfg, double fcut_host = 0.0,
                         const double* fmin_g = nullptr, int ngroups = 0) {
    typedef Dual<NT> D;
    r.has_ecc = 0;
    r.pad_ = 0;
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
    if ((cfg.flags & kFlagEccentric) && fmin_g) {
        r.has_ecc = 1;
        for (int g = 0; g < ngroups; ++g) {
            // v0ecc = min_f v (the grid's first sample) or (pi M fRef_ecc)^(1/3), waveforms.py:817-820
            const double f0 = (cfg.flags & kFlagHasFRef) ? cfg.fRef : fmin_g[g];
            tf2_ecc_coeffs<NT>(r.ec[g], p.eta, p.ecc, dpow(kPi * s * f0, 1. / 3.));
        }
    }
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
GWF_HD void tf2_phase(const TF2Rec<NT>& r, const VPow& p, double& phi, double* phi_d, int g = 0) {
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
    if (r.has_ecc) {
        // v^(-34/3) = v^-11 v^(-1/3)
        const double vm13 = 1.0 / cbrt(v), vm11 = vm5 * vm5 * p.vm1, e0 = vm11 * vm13, v3 = v2 * v, v4 = v2 * v2, v6 = v3 * v3;
        const double eb[kTF2Ecc] = {e0, e0 * v2, e0 * v3, e0 * v4, e0 * v5, e0 * v6, e0 * v6 * p.lv};
        const double q0 = -34. / 9.;
        const double ebx[kTF2Ecc] = {q0 * eb[0], (q0 + 2. / 3.) * eb[1], (q0 + 1.) * eb[2], (q0 + 4. / 3.) * eb[3], (q0 + 5. / 3.) * eb[4],
                                     (q0 + 2.) * eb[5], (q0 + 2.) * eb[6] + 1. / 3. * eb[5]};
#pragma unroll
        for (int k = 0; k < kTF2Ecc; ++k) {
            val_ = fma(r.ec[g][k][0], eb[k], val_);
            dx = fma(r.ec[g][k][0], ebx[k], dx);
#pragma unroll
            for (int j = 0; j < NT; ++j) d[j] = fma(r.ec[g][k][1 + j], eb[k], d[j]);
        }
    }
    phi = val_;
#pragma unroll
    for (int j = 0; j < NT; ++j) phi_d[j] = d[j] + dx * r.lam[j];
}

}  // namespace gwf
