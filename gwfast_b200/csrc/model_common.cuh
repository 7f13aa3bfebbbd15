// Shared model plumbing: per-event inputs, Fisher-parametrisation seeding, waveform-point outputs, tau(f).
#pragma once
#include "pn_common.cuh"

namespace gwf {

constexpr int kMaxGroups = 4;   // distinct (fmin, fmax) frequency-grid groups in one network launch

// model ids (gwf_model.id in include/gwfast_b200.h)
enum ModelId { kTaylorF2 = 0, kPhenomD = 1, kNRTidalv2 = 2, kPhenomHM = 3, kNSBH = 4 };

// model option flags (gwf_model.flags)
enum ModelFlags {
    kFlagTidal = 1,          // TaylorF2(is_tidal=True)
    kFlag3p5SpinHO = 2,      // TaylorF2(use_3p5PN_SpinHO=True)
    kFlagPhirefVlso = 4,     // TaylorF2(phiref_vlso=True)
    kFlagQuadMonTid = 8,     // TaylorF2(use_QuadMonTid=True)
    kFlagKerrISCO = 16,      // TaylorF2(which_ISCO='Kerr')
    kFlagNoFcut = 32,        // apply_fcut=False
    kFlagHasFRef = 64,       // PhenomD-family fRef given by the user
    kFlagLambdaGiven = 128,  // the events carried Lambda1/Lambda2 when fcut() was evaluated
    kFlagEccentric = 512,    // TaylorF2(is_eccentric=True): extra parameter ecc, reference frequency fRef_ecc in ModelCfg::fRef if kFlagHasFRef
    kFlagNewtonian = 256,    // NewtInspiral (waveforms.py:205-260): leading-order phase, numerical tau_star of the base class
};

// Fisher parametrisation flags (gwf_opts.flags)
enum OptFlags {
    kOptM1M2 = 1,         // differentiate w.r.t. (m1, m2) instead of (Mc, eta); signal.py:818-822, 522-527
    kOptChiSChiA = 2,     // differentiate w.r.t. (chiS, chiA) instead of (chi1z, chi2z); signal.py:838-842
    kOptLinGrid = 4,      // spacing='lin'; signal.py:895-896
};

struct ModelCfg {
    int id;
    int flags;
    double fcutPar;   // Hz*Msun for TaylorF2 (Schw ISCO), dimensionless Mf for the PhenomD family
    double fRef;      // Hz, only if kFlagHasFRef
};

// plain per-event inputs as the events dict carries them
// reciprocal for the per-sample weights: MUFU seed + two Newton steps (error < 1 ulp; no denormal/special-case branch
// of the IEEE division -- the arguments here are PSD values and amplitude denominators, always normal and positive)
GWF_HD double rcp_fast(double x) {
#ifdef __CUDA_ARCH__
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);          // seed 2^-23 -> 2^-46 -> rounding level
#else
    return 1.0 / x;
#endif
}

struct EventIn {
    double Mc, eta, dL, theta, phi, iota, psi, tcoal, Phicoal, chi1z, chi2z, Lambda1, Lambda2;
    double fcut_host, s_host;   // optional host-computed wf_model.fcut and M*GMsun_over_c3 (0 = compute on the device)
    double ecc;                 // orbital eccentricity e0 (eccentric TaylorF2 only)
    // upper ends of the grid groups' bands in Hz (0 = none), set by the caller of the prologue: IMRPhenomNSBH takes its time shift at
    // the LAST sample of the event's grid (waveforms.py:2994).  fmax_exact: fmax_g[0] is that sample itself (user grids)
    const double* fmax_g = nullptr;
    bool fmax_exact = false;
};

// intrinsic parameters seeded for differentiation.  Slots: 0,1 = (Mc,eta) or (m1,m2); 2,3 = (chi1z,chi2z) or
// (chiS,chiA); 4,5 = (LambdaTilde, deltaLambda) for tidal models; the last slot of an odd NT (5, 7) = ecc.  The re-mapping back to what the waveform
// consumes is done in dual arithmetic exactly like GWstrain does it (signal.py:522-559).
template <int NT>
struct Intrinsic {
    Dual<NT> Mc, eta, chi1, chi2, L1, L2, ecc;
};

template <int NT>
GWF_HD Intrinsic<NT> seed_intrinsic(const EventIn& e, int opt_flags, bool tidal) {
    typedef Dual<NT> D;
    Intrinsic<NT> p;
    if (opt_flags & kOptM1M2) {
        // m1m2_from_Mceta (gwfastUtils.py:450) then Mceta_from_m1m2 (gwfastUtils.py:466) on the seeded masses
        const double sq = seta_of(e.eta);
        const double M = e.Mc / pow(e.eta, 3. / 5.);
        const D m1 = D::seed(0.5 * M * (1. + sq), 0), m2 = D::seed(0.5 * M * (1. - sq), 1);
        const D prod = m1 * m2, sum = m1 + m2;
        p.Mc = dpow(prod, 3. / 5.) / dpow(sum, 1. / 5.);
        p.eta = prod / (sum * sum);
    } else {
        p.Mc = D::seed(e.Mc, 0);
        p.eta = D::seed(e.eta, 1);
    }
    if (opt_flags & kOptChiSChiA) {
        const D cs = D::seed(0.5 * (e.chi1z + e.chi2z), 2), ca = D::seed(0.5 * (e.chi1z - e.chi2z), 3);
        p.chi1 = cs + ca;
        p.chi2 = cs - ca;
    } else {
        p.chi1 = D::seed(e.chi1z, 2);
        p.chi2 = D::seed(e.chi2z, 3);
    }
    // eccentric TaylorF2: e0 is differentiated in the last slot (ParNums['ecc'] = 11 or 13, waveforms.py:113-124)
    p.ecc = (NT == 5 || NT == 7) ? D::seed(e.ecc, NT - 1) : D(e.ecc);
    if (tidal && NT >= 6) {
        // FisherMatr: (Lambda1,Lambda2,etaOr) -> (LambdaTilde, deltaLambda) (signal.py:871); GWstrain maps them
        // back with the *seeded* eta (signal.py:554), which adds an eta-dependence on the AD path.
        double lt, dl;
        lamt_dellam_from_lam12(e.Lambda1, e.Lambda2, e.eta, lt, dl);
        const D LT = D::seed(lt, NT >= 6 ? 4 : 0), DL = D::seed(dl, NT >= 6 ? 5 : 0);
        lam12_from_lamt_dellam(LT, DL, p.eta, p.L1, p.L2);
    } else {
        p.L1 = D(e.Lambda1);
        p.L2 = D(e.Lambda2);
    }
    return p;
}

// value-only version of the same re-mapping (used by the SNR / waveform-value kernels: SNRInteg passes the
// dict entries straight to the waveform, signal.py:715-726)
struct IntrinsicV {
    double Mc, eta, chi1, chi2, L1, L2;
};

// one sample of the frequency grid with the powers every model needs.  On a geometric grid these are advanced by
// recurrence (a handful of multiplications per sample instead of exp10/cbrt/log/sqrt calls), see Grid in fisher_core.cuh.
struct FreqPoint {
    double f;      // Hz
    double f13;    // f^(1/3)
    double fm13;   // f^(-1/3)
    double fm76;   // f^(-7/6)
    double lnf;    // ln f
    double w;      // trapezoid weight (f_{k+1} - f_{k-1})/2, one-sided at the ends
    GWF_HD void from_f(double f_) {
        f = f_;
        f13 = cbrt(f_);
        fm13 = 1.0 / f13;
        fm76 = fm13 * fm13 * fm13 * sqrt(fm13);
        lnf = log(f_);
    }
};
// per-event scale s (x = s f) in the same powers
struct ScalePow {
    double s13, sm13, lps3;   // s^(1/3), s^(-1/3), ln(pi s)/3
    GWF_HD void set(double s) {
        s13 = cbrt(s);
        sm13 = 1.0 / s13;
        lps3 = log(kPi * s) * (1. / 3.);
    }
};

// waveform quantities at one frequency: amplitude, d(ln A), phase tangent (value optional)
template <int NT>
struct WfPoint {
    double A;           // amplitude (0 beyond the cut)
    double lnA_d[NT];   // d ln A / d intrinsic_j
    double phi;         // phase value (only filled when requested)
    double phi_d[NT];   // d Phi / d intrinsic_j
};

// tau(f) and its tangents (slots 0,1 only: tau depends on Mc, eta alone), basis form of pn_common.cuh.
// lpx3 = log(pi x)/3, vm1 = (pi x)^(-1/3); lam[j] = d ln(s)_j so that d(basis)/d slot j = (x d/dx basis) * lam[j]
struct TauRec {
    double t[kTau];
    double td[2][kTau];
};
GWF_HD void tau_eval(const TauRec& r, double vm1, double lpx3, const double* lam, double& tau, double* dtau) {
    const double vm2 = vm1 * vm1, vm3 = vm2 * vm1, vm4 = vm2 * vm2, vm5 = vm4 * vm1, vm6 = vm3 * vm3, vm8 = vm4 * vm4;
    const double lg = 4.0 * kLn2 + 2.0 * lpx3;   // log(16 v^2)
    const double b[kTau] = {vm8, vm6, vm5, vm4, vm3, vm2, vm2 * lg, vm1};
    // x d/dx of each basis term (v = (pi x)^(1/3) => x dv^p/dx = p/3 v^p)
    const double bx[kTau] = {-8. / 3. * vm8, -2. * vm6, -5. / 3. * vm5, -4. / 3. * vm4, -vm3, -2. / 3. * vm2,
                             -2. / 3. * vm2 * lg + 2. / 3. * vm2, -1. / 3. * vm1};
    double v = 0., d0 = 0., d1 = 0., dx = 0.;
#pragma unroll
    for (int k = 0; k < kTau; ++k) {
        v = fma(r.t[k], b[k], v);
        d0 = fma(r.td[0][k], b[k], d0);
        d1 = fma(r.td[1][k], b[k], d1);
        dx = fma(r.t[k], bx[k], dx);
    }
    tau = v;
    dtau[0] = fma(dx, lam[0], d0);
    dtau[1] = fma(dx, lam[1], d1);
}
// WaveFormModel.tau_star, the base-class default used by NewtInspiral (waveforms.py:176-186):
// 2.18567 (1.21/Mc)^(5/3) (100/f)^(8/3), written on the v^-8 basis slot: tau = t0 (pi s f)^(-8/3)
template <int NT>
GWF_HD void tau_fill_newtonian(TauRec& r, const Dual<NT>& Ms, const Dual<NT>& Mc) {
    const Dual<NT> t0 = 2.18567 * dpow(1.21 / Mc, 5. / 3.) * pow(100., 8. / 3.) * dpow(kPi * Ms, 8. / 3.);
#pragma unroll
    for (int k = 0; k < kTau; ++k) r.t[k] = r.td[0][k] = r.td[1][k] = 0.0;
    r.t[0] = t0.v;
    r.td[0][0] = t0.d[0];
    r.td[1][0] = t0.d[1];
}
template <int NT>
GWF_HD void tau_fill(TauRec& r, const Dual<NT>& Ms, const Dual<NT>& eta) {
    Dual<NT> t[kTau];
    tau_coeffs(Ms, eta, t);
#pragma unroll
    for (int k = 0; k < kTau; ++k) {
        r.t[k] = t[k].v;
        r.td[0][k] = t[k].d[0];
        r.td[1][k] = t[k].d[1];
    }
}

}  // namespace gwf
