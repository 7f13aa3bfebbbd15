// IMRPhenomHM (gwfast/waveforms.py:1838-2749): six modes (21, 22, 32, 33, 43, 44), each an IMRPhenomD amplitude and
// phase evaluated at a piecewise-linearly mapped frequency, rescaled with 1.5PN amplitude ratios and combined with
// spin-weighted spherical harmonics into h+ and hx (hphc, waveforms.py:2256-2616).
//
// Unlike the (2,2)-only models no common phase factors out of the strain, so the per-frequency evaluation is done
// on complex values.  The per-event prologue stores every f-independent quantity as a dual (value + tangents w.r.t.
// Mc, eta, chi1z, chi2z; iota enters only through the harmonics and is differentiated in closed form), and the
// per-frequency code is written once, generically in the scalar type: Dual<NT> for the Fisher rows, double for
// the SNR.  QNM frequencies come from the complex polynomial fits of _RDfreqCalc (waveforms.py:2661-2710), not
// from the tables.
#pragma once
#include "model_phenomd.cuh"

namespace gwf {

constexpr int kHMModes = 6;   // order 21, 22, 32, 33, 43, 44 (waveforms.py:2305)

// ---- extra dual functions used only here
GWF_HD double dcbrt(double x) { return cbrt(x); }
template <int N> GWF_HD Dual<N> dcbrt(const Dual<N>& a) { const double c = cbrt(a.v); return chain(a, c, c / (3.0 * a.v)); }
GWF_HD void dsincos(double x, double& s, double& c) { sincos(x, &s, &c); }
template <int N> GWF_HD void dsincos(const Dual<N>& a, Dual<N>& s, Dual<N>& c) {
    double sv, cv;
    sincos(a.v, &sv, &cv);
    s = chain(a, sv, cv);
    c = chain(a, cv, -sv);
}
template <class T> GWF_HD T lift(double v) { return T(v); }

// complex polynomial fits of the QNM frequencies in kappa = (log(2 - a)/log 3)^(1/(2 + l - m)), waveforms.py:2683-2699:
// res = sum_k c_k exp(i th_k) kappa^k
struct QnmFit { double beta; double c[7], th[7]; };
#define GWF_HM_QNM_FITS \
    {1. / 3., {0.589113, 0.18896353, 1.15012965, 6.04585476, 11.12627777, 9.34711461, 3.03838318}, {0.043525, 2.289868, 5.810057, 2.741967, 5.844130, 2.669372, 5.791518}}, \
    {0.5, {1.0, 1.557847, 1.95097051, 2.09971716, 1.41094660, 0.41063923, 0.}, {0., 2.903124, 5.920970, 2.760585, 5.914340, 2.795235, 0.}}, \
    {1. / 3., {1.022464, 0.24731213, 1.70468239, 0.94604882, 1.53189884, 2.28052668, 0.92150314}, {0.004870, 0.665292, 3.138283, 0.163247, 5.703573, 2.685231, 5.841704}}, \
    {0.5, {1.5, 2.095657, 2.46964352, 2.66552551, 1.75836443, 0.49905688, 0.}, {0., 2.964973, 5.996734, 2.817591, 5.932693, 2.781658, 0.}}, \
    {1. / 3., {1.5, 0.205046, 3.10333396, 4.23612166, 3.02890198, 0.90843949, 0.}, {0., 0.595328, 3.016200, 6.038842, 2.826239, 5.915164, 0.}}, \
    {0.5, {2.0, 2.658908, 2.97825567, 3.21842350, 2.12764967, 0.60338186, 0.}, {0., 3.002787, 6.050955, 2.877514, 5.989669, 2.830031, 0.}}
__device__ __constant__ QnmFit kQnmFitsDev[kHMModes] = {GWF_HM_QNM_FITS};
static const QnmFit kQnmFitsHost[kHMModes] = {GWF_HM_QNM_FITS};
GWF_HD const QnmFit& qnm_fit(int m) {
#ifdef __CUDA_ARCH__
    return kQnmFitsDev[m];
#else
    return kQnmFitsHost[m];
#endif
}
GWF_HD int hm_ell(int m) { return m < 2 ? 2 : (m < 4 ? 3 : 4); }
GWF_HD int hm_mm(int m) { return m == 0 ? 1 : (m == 1 ? 2 : (m == 2 ? 2 : (m == 3 ? 3 : (m == 4 ? 3 : 4)))); }
// complShiftm[mm], waveforms.py:1872
GWF_HD double hm_shift(int mm) { return mm == 1 ? 0.5 * kPi : (mm == 3 ? -0.5 * kPi : (mm == 4 ? kPi : 0.0)); }

template <class T> GWF_HD void qnm_freqs(int m, const T& aeff, const T& finMass, T& fring, T& fdamp) {
    const QnmFit& q = qnm_fit(m);
    const T alpha = dlog(2. - aeff) / log(3.);
    const T kappa = dpow(alpha, q.beta);
    T re(q.c[6] * cos(q.th[6])), im(q.c[6] * sin(q.th[6]));
    for (int k = 5; k >= 0; --k) {                 // Horner in kappa
        re = re * kappa + q.c[k] * cos(q.th[k]);
        im = im * kappa + q.c[k] * sin(q.th[k]);
    }
    fring = re / (2. * kPi * finMass);
    fdamp = im / (2. * kPi * finMass);
}

// per-mode f-independent quantities
template <class T>
struct HMMode {
    T fi_amp, fi_phi, fr;             // map break points (amp: 0.014/Rho, phase: 0.018/Rho, fring_lm)
    T am_amp, bm_amp, br_amp;         // amplitude map, middle and ringdown regimes (inspiral: 2/m, 0; ringdown slope 1)
    T am_phi, bm_phi, rho;            // phase map (ringdown: slope Rho, 0)
    T inv_am_phi, inv_rho;            // reciprocals of the two map slopes (the mapped phase is divided by them)
    T c1, c2;                         // C1MRDHM, C2MRDHM
    T at_c, at_a, at_w;               // alpha4 Rho/eta, alpha5 fring, fdamp Rho Tau
    T kB, kC;                         // -PhDBconst + PhDBAterm, -PhDCconst + tmpphaseC
};

template <int NT>
struct HMRec {
    typedef Dual<NT> D;
    double fcut_hz;
    double x_mrd, x_peak;             // values of fMRDJoin and fpeak (region tests)
    D s;                              // M GMsun/c^3: x = s f
    D eta, seta, chis, chia;          // for the 1.5PN amplitude ratios
    D Camp;                           // M^2 GMsun_c2_Gpc GMsun_c3/dL * amp0   (no 2 sqrt(5/64pi): waveforms.py:2172-2174)
    D t0;
    D xref[kMaxGroups], phi0[kMaxGroups];
    D pins[kPIns];                    // inspiral phase coefficients (same basis as PhenomDRec, -t0 NOT folded in)
    D pint[4];                        // C1Int, beta1/eta + C2Int, -beta3/(3 eta), beta2/eta
    D pmrd[3];                        // -alpha2/eta, 4/3 alpha3/eta, alpha1/eta
    D ains[kAIns], aint[kAInt], amrd[4];
    HMMode<D> mode[kHMModes];
    TauRec tau;
    double lam[NT];                   // d ln s
};

// accessor: the record stores duals; the value-only path reads their values
template <class T> struct Ld;
template <> struct Ld<double> { template <int N> static GWF_HD double get(const Dual<N>& c) { return c.v; } };
template <int N> struct Ld<Dual<N>> { static GWF_HD const Dual<N>& get(const Dual<N>& c) { return c; } };

// ---- PhenomD phase at mapped frequency y for one mode (completePhase, waveforms.py:2023-2027)
template <class T, int NT>
GWF_HD T hm_complete_phase(const HMRec<NT>& r, const T& y, const T& c1, const T& c2, const T& at_c, const T& at_a, const T& at_w, bool apply_cut) {
    typedef Ld<T> L;
    const double yv = val(y);
    if (yv < kPhiJoinIns) {
        const T y13 = dcbrt(y), y23 = y13 * y13, lg = dlog(kPi * y) * (1. / 3.);
        const T ym13 = 1.0 / y13, ym23 = ym13 * ym13, ym1 = ym23 * ym13;
        return L::get(r.pins[0]) + L::get(r.pins[1]) * y23 + L::get(r.pins[2]) * y13 + L::get(r.pins[3]) * (y13 * lg) + L::get(r.pins[4]) * lg +
               L::get(r.pins[5]) * ym13 + L::get(r.pins[6]) * ym23 + L::get(r.pins[7]) * ym1 + L::get(r.pins[8]) * (ym1 * ym23) + L::get(r.pins[9]) * y +
               L::get(r.pins[10]) * (y * y13) + L::get(r.pins[11]) * (y * y23) + L::get(r.pins[12]) * (y * y);
    }
    if (yv < r.x_mrd) {
        const T ym1 = 1.0 / y;
        return L::get(r.pint[0]) + L::get(r.pint[1]) * y + L::get(r.pint[2]) * (ym1 * ym1 * ym1) + L::get(r.pint[3]) * dlog(y);
    }
    if (!apply_cut || yv < kMfCut) {
        const T sy = dsqrt(y);
        return L::get(r.pmrd[0]) / y + L::get(r.pmrd[1]) * (sy * dsqrt(sy)) + L::get(r.pmrd[2]) * y + at_c * datan((y - at_a) / at_w) + c1 + c2 * y;
    }
    return T(0.0);
}

// ---- PhenomD amplitude shape at mapped frequency y (the where of completeAmpl, waveforms.py:2176-2180)
template <class T, int NT>
GWF_HD T hm_amp_shape(const HMRec<NT>& r, const T& y, const T& y13, bool apply_cut) {
    typedef Ld<T> L;
    const double yv = val(y);
    if (yv < kAmpJoinIns) {
        const T y23 = y13 * y13, y2 = y * y;
        return L::get(r.ains[0]) + L::get(r.ains[1]) * y23 + L::get(r.ains[2]) * y + L::get(r.ains[3]) * (y * y13) + L::get(r.ains[4]) * (y * y23) +
               L::get(r.ains[5]) * y2 + L::get(r.ains[6]) * (y2 * y13) + L::get(r.ains[7]) * (y2 * y23) + L::get(r.ains[8]) * (y2 * y);
    }
    if (yv < r.x_peak) {
        const T u = y - kAmpJoinIns;
        return L::get(r.aint[0]) + u * (L::get(r.aint[1]) + u * (L::get(r.aint[2]) + u * (L::get(r.aint[3]) + u * L::get(r.aint[4]))));
    }
    if (!apply_cut || yv < kMfCut) {
        const T u = y - L::get(r.amrd[0]), w = L::get(r.amrd[2]);
        return dexp(-u * L::get(r.amrd[1])) * L::get(r.amrd[3]) / (u * u + w * w);
    }
    return T(0.0);
}

// |H_lm(v)| of OnePointFiveSpinPN (waveforms.py:2182-2212 / 2503-2515), without the common pi sqrt(2 eta/3) v^-3.5
template <class T>
GWF_HD T hm_absH(int m, const T& v, const T& eta, const T& seta, const T& chis, const T& chia) {
    const T v2 = v * v;
    switch (m) {
        case 0: {   // (2,1): complex
            const T v3 = v2 * v;
            const T re = (sqrt(2.0) / 3.0) * (v * seta - v2 * 1.5 * (chia + seta * chis) + v3 * seta * ((335.0 / 672.0) + (eta * 117.0 / 56.0)) +
                                              v3 * v * (chia * (3427.0 / 1344. - eta * 2101.0 / 336.) + seta * chis * (3427.0 / 1344 - eta * 965. / 336.) + seta * (-kPi)));
            const T im = (sqrt(2.0) / 3.0) * (v3 * v * seta * (-0.5 - 2 * 0.69314718056));
            return dsqrt(re * re + im * im);
        }
        case 1: return T(1.0);
        case 2: return dfabs((1.0 / 3.0) * sqrt(5.0 / 7.0) * (v2 * (1.0 - 3.0 * eta)));
        case 3: return dfabs(0.75 * sqrt(5.0 / 7.0) * (v * seta));
        case 4: return dfabs(0.75 * sqrt(3.0 / 35.0) * (v2 * v) * seta * (1.0 - 2.0 * eta));
        default: return dfabs((4.0 / 9.0) * sqrt(10.0 / 7.0) * v2 * (1.0 - 3.0 * eta));
    }
}

// harmonic weights of hphc (waveforms.py:2517-2524, 2613-2614): hp = sum z_m Wp_m, hc = i sum z_m Wc_m, and d/d iota
struct HMWeights {
    double wp[kHMModes], wc[kHMModes], dwp[kHMModes], dwc[kHMModes];
    GWF_HD static void ylm(int m, double th, double& Y, double& Ym) {
        const double c = cos(th), s = sin(th), ch = cos(0.5 * th), sh = sin(0.5 * th);
        const double ch2 = ch * ch, ch4 = ch2 * ch2, sh2 = sh * sh, sh4 = sh2 * sh2;
        switch (m) {
            case 0: Y = sqrt(5.0 / (16.0 * kPi)) * s * (1.0 + c); Ym = sqrt(5.0 / (16.0 * kPi)) * s * (1.0 - c); break;
            case 1: Y = sqrt(5.0 / (64.0 * kPi)) * (1.0 + c) * (1.0 + c); Ym = sqrt(5.0 / (64.0 * kPi)) * (1.0 - c) * (1.0 - c); break;
            case 2: Y = sqrt(7.0 / kPi) * ch4 * (-2.0 + 3.0 * c) * 0.5; Ym = sqrt(7.0 / (4.0 * kPi)) * (2.0 + 3.0 * c) * sh4; break;
            case 3: Y = -sqrt(21.0 / (2.0 * kPi)) * (ch4 * ch) * sh; Ym = sqrt(21.0 / (2.0 * kPi)) * ch * (sh4 * sh); break;
            case 4: Y = -3.0 * sqrt(7.0 / (2.0 * kPi)) * (ch4 * ch) * (-1.0 + 2.0 * c) * sh; Ym = 3.0 * sqrt(7.0 / (2.0 * kPi)) * ch * (1.0 + 2.0 * c) * (sh4 * sh); break;
            default: Y = 3.0 * sqrt(7.0 / kPi) * (ch4 * ch2) * sh2; Ym = 3.0 * sqrt(7.0 / kPi) * ch2 * (sh4 * sh2); break;
        }
    }
    GWF_HD void set(double iota) {
        // the harmonics are trigonometric polynomials of iota/2; their iota-derivative is taken by a 4th-order central
        // difference in extended precision-free form: use the analytic product rule instead (exact)
        for (int m = 0; m < kHMModes; ++m) {
            double Y, Ym;
            ylm(m, iota, Y, Ym);
            const double sg = (hm_ell(m) & 1) ? -1.0 : 1.0;
            wp[m] = 0.5 * (Y + sg * Ym);
            wc[m] = 0.5 * (Y - sg * Ym);
            double dY, dYm;
            dylm(m, iota, dY, dYm);
            dwp[m] = 0.5 * (dY + sg * dYm);
            dwc[m] = 0.5 * (dY - sg * dYm);
        }
    }
    GWF_HD static void dylm(int m, double th, double& dY, double& dYm) {
        const double c = cos(th), s = sin(th), ch = cos(0.5 * th), sh = sin(0.5 * th);
        const double ch2 = ch * ch, ch3 = ch2 * ch, ch4 = ch2 * ch2, ch5 = ch4 * ch, sh2 = sh * sh, sh3 = sh2 * sh, sh4 = sh2 * sh2, sh5 = sh4 * sh;
        switch (m) {
            case 0: {
                const double k = sqrt(5.0 / (16.0 * kPi));
                dY = k * (c * (1.0 + c) - s * s); dYm = k * (c * (1.0 - c) + s * s); break;
            }
            case 1: {
                const double k = sqrt(5.0 / (64.0 * kPi));
                dY = -2.0 * k * (1.0 + c) * s; dYm = 2.0 * k * (1.0 - c) * s; break;
            }
            case 2: {
                // Y = k ch^4 (-2 + 3c)/2 ; d(ch^4) = -2 ch^3 sh ; dc = -s
                const double k = sqrt(7.0 / kPi);
                dY = k * 0.5 * (-2.0 * ch3 * sh * (-2.0 + 3.0 * c) - 3.0 * s * ch4);
                const double k2 = sqrt(7.0 / (4.0 * kPi));
                dYm = k2 * (-3.0 * s * sh4 + (2.0 + 3.0 * c) * 2.0 * sh3 * ch);
                break;
            }
            case 3: {
                const double k = sqrt(21.0 / (2.0 * kPi));
                dY = -k * (-2.5 * ch4 * sh2 + 0.5 * ch5 * ch);                       // d(ch^5 sh) = -5/2 ch^4 sh^2 + 1/2 ch^6
                dYm = k * (-0.5 * sh5 * sh + 2.5 * ch2 * sh4);                        // d(ch sh^5) = -1/2 sh^6 + 5/2 ch^2 sh^4
                break;
            }
            case 4: {
                const double k = 3.0 * sqrt(7.0 / (2.0 * kPi));
                // Y = -k ch^5 sh (-1 + 2c)
                dY = -k * ((-2.5 * ch4 * sh2 + 0.5 * ch5 * ch) * (-1.0 + 2.0 * c) - 2.0 * s * ch5 * sh);
                // Ym = k ch sh^5 (1 + 2c)
                dYm = k * ((-0.5 * sh5 * sh + 2.5 * ch2 * sh4) * (1.0 + 2.0 * c) - 2.0 * s * ch * sh5);
                break;
            }
            default: {
                const double k = 3.0 * sqrt(7.0 / kPi);
                // Y = k ch^6 sh^2 : d = -3 ch^5 sh^3 + ch^7 sh ;  Ym = k ch^2 sh^6 : d = -ch sh^7 + 3 ch^3 sh^5
                dY = k * (-3.0 * ch5 * sh3 + ch5 * ch2 * sh);
                dYm = k * (-ch * sh5 * sh2 + 3.0 * ch3 * sh5);
                break;
            }
        }
    }
};

// ------------------------------------------------------------------------------------------------ prologue
template <int NT>
GWF_HD void phenomhm_prologue(HMRec<NT>& r, const Intrinsic<NT>& p, double dL, const double* fmin_g, int ngroups, const ModelCfg& cfg,
                              double s_host = 0.0, double fcut_host = 0.0) {
    typedef Dual<NT> D;
    const bool apply_cut = !(cfg.flags & kFlagNoFcut);
    const D eta = p.eta;
    const D M = p.Mc / dpow(eta, 3. / 5.);
    D s = M * kGMsunC3;
    if (s_host > 0.0) s.v = s_host;       // the host's M*GMsun_over_c3: x = s f then rounds like the reference's fgrid
    r.s = s;
#pragma unroll
    for (int j = 0; j < NT; ++j) r.lam[j] = s.d[j] / s.v;
    r.fcut_hz = fcut_host > 0.0 ? fcut_host : kMfCut / s.v;    // waveforms.py:2737-2749
    r.eta = eta;
    r.seta = seta_of(eta);
    r.chis = 0.5 * (p.chi1 + p.chi2);
    r.chia = 0.5 * (p.chi1 - p.chi2);
    const D aeff = final_spin(eta, p.chi1, p.chi2), finMass = 1. - radiated_energy(eta, p.chi1, p.chi2);
    D fring, fdamp;
    qnm_freqs(1, aeff, finMass, fring, fdamp);                 // (2,2), waveforms.py:2336-2345
    PhenomDCore<NT> c;
    c.build_rd(eta, p.chi1, p.chi2, D(1.0), D(1.0), fring, fdamp);
    r.x_mrd = c.fMRDJoin.v;
    r.x_peak = c.fpeak_amp.v;
    const D amp0 = dsqrt(2.0 * eta / 3.0) * pow(kPi, -1. / 6.);
    r.Camp = M * kGMsunC2Gpc * M * kGMsunC3 / D(dL) * amp0;
    const D t0 = c.dphi_mrd(c.fpeak_amp);                      // waveforms.py:2527
    r.t0 = t0;
    const double cp = cbrt(kPi), cp2 = cp * cp;
    const D n = c.norm, ie = 1.0 / eta;
    r.pins[0] = c.pn.c5 * n; r.pins[1] = c.pn.c7 * n * cp2; r.pins[2] = c.pn.c6 * n * cp; r.pins[3] = (-6848. / 21.) * n * cp;
    r.pins[4] = 3. * c.pn.c5 * n; r.pins[5] = c.pn.c4 * n / cp; r.pins[6] = c.pn.c3 * n / cp2; r.pins[7] = c.pn.c2 * n / kPi;
    r.pins[8] = n / (kPi * cp2); r.pins[9] = c.fit[SIG1] * ie; r.pins[10] = c.fit[SIG2] * 0.75 * ie; r.pins[11] = c.fit[SIG3] * 0.6 * ie;
    r.pins[12] = c.fit[SIG4] * 0.5 * ie;
    r.pint[0] = c.C1Int; r.pint[1] = c.fit[BET1] * ie + c.C2Int; r.pint[2] = -c.fit[BET3] * ie / 3.; r.pint[3] = c.fit[BET2] * ie;
    r.pmrd[0] = -c.fit[ALP2] * ie; r.pmrd[1] = (4.0 / 3.0) * c.fit[ALP3] * ie; r.pmrd[2] = c.fit[ALP1] * ie;
    r.ains[0] = D(1.0);
    for (int k = 1; k < kAIns; ++k) r.ains[k] = c.A[k + 1];
    for (int k = 0; k < kAInt; ++k) r.aint[k] = c.e[k];
    const D fd3 = fdamp * c.fit[GAM3];
    r.amrd[0] = fring; r.amrd[1] = c.fit[GAM2] / fd3; r.amrd[2] = fd3; r.amrd[3] = fd3 * c.fit[GAM1];
    // reference phase of the (2,2) mode: completePhase(fRef, C1MRD, C2MRD, 1, 1)/2, waveforms.py:2529-2530
    const D at_c22 = c.fit[ALP4] * ie, at_a22 = c.fit[ALP5] * fring;
    for (int g = 0; g < ngroups; ++g) {
        const D xref = (cfg.flags & kFlagHasFRef) ? s * cfg.fRef : s * fmin_g[g];
        r.xref[g] = xref;
        r.phi0[g] = 0.5 * hm_complete_phase<D, NT>(r, xref, c.C1MRD, c.C2MRD, at_c22, at_a22, fdamp, apply_cut);
    }
    // per-mode quantities, waveforms.py:2534-2605
    const D fm = c.fMRDJoin;
    const D PhiIntTempVal = c.phi_int_raw(fm) / eta + c.C1Int + c.C2Int * fm;
    const D DPhiIntTempVal = c.C2Int + c.dphi_int(fm);
    for (int m = 0; m < kHMModes; ++m) {
        HMMode<D>& o = r.mode[m];
        const double mm = hm_mm(m);
        D frlm, fdlm;
        qnm_freqs(m, aeff, finMass, frlm, fdlm);
        const D Rho = fring / frlm, Tau = fdlm / fdamp;
        const D u = fm - c.fit[ALP5] * fring;
        const D wtr = fdamp * Tau * Rho;
        const D DPhiMRDVal = (c.fit[ALP1] + c.fit[ALP2] / (fm * fm) + c.fit[ALP3] / dpow(fm, 0.25) + c.fit[ALP4] / (fdamp * Tau * (1. + u * u / (wtr * wtr)))) / eta;
        const D PhiMRJoinTemp = -(c.fit[ALP2] / fm) + (4.0 / 3.0) * (c.fit[ALP3] * dpow(fm, 0.75)) + c.fit[ALP1] * fm + c.fit[ALP4] * Rho * datan(u / wtr);
        o.c2 = DPhiIntTempVal - DPhiMRDVal;
        o.c1 = PhiIntTempVal - PhiMRJoinTemp / eta - o.c2 * fm;
        o.at_c = c.fit[ALP4] * Rho * ie;
        o.at_a = c.fit[ALP5] * fring;
        o.at_w = wtr;
        o.rho = Rho;
        o.fr = frlm;
        const double ai = 2. / mm;
        // amplitude map, waveforms.py:2547-2563
        o.fi_amp = kAmpJoinIns / Rho;
        {
            const D Trd = frlm - frlm + fring, Ti = 2. * o.fi_amp / mm;
            o.am_amp = (Trd - Ti) / (frlm - o.fi_amp);
            o.bm_amp = Ti - o.fi_amp * o.am_amp;
            o.br_amp = -frlm + fring;
        }
        // phase map
        o.fi_phi = kPhiJoinIns / Rho;
        {
            const D Trd = frlm * Rho, Ti = 2. * o.fi_phi / mm;
            o.am_phi = (Trd - Ti) / (frlm - o.fi_phi);
            o.bm_phi = Ti - o.fi_phi * o.am_phi;
        }
        o.inv_am_phi = 1.0 / o.am_phi;
        o.inv_rho = 1.0 / Rho;
        // continuity constants, waveforms.py:2586-2596
        auto cP = [&](const D& y) { return hm_complete_phase<D, NT>(r, y, o.c1, o.c2, o.at_c, o.at_a, o.at_w, apply_cut); };
        const D PhDBconst = cP(o.am_phi * o.fi_phi + o.bm_phi) / o.am_phi;
        const D PhDCconst = cP(Rho * frlm) / Rho;
        const D PhDBAterm = cP(ai * o.fi_phi) / ai;
        const D tmpphaseC = -PhDBconst + PhDBAterm + cP(o.am_phi * frlm + o.bm_phi) / o.am_phi;
        o.kB = -PhDBconst + PhDBAterm;
        o.kC = -PhDCconst + tmpphaseC;
    }
    tau_fill(r.tau, s, eta);
}

// ------------------------------------------------------------------------------------------------ per frequency
// amplitude A_lm and phase Phi_lm of the six modes (IMRPhenomHM.Ampl / .Phi, waveforms.py:1874-2254, as vectorised in hphc);
// T = Dual<NT> (tangents w.r.t. the intrinsic slots) or double.  Every mode is handed to `sink(m, A, Phi)` as soon as it is
// known, so that callers which only need sums over the modes keep nothing per mode (a rolled loop over per-mode arrays
// lives in local memory).  The PhenomD phase is evaluated through ONE call site per mode -- the three frequency regimes only
// choose its argument, offset and scale -- which keeps the kernel's code (and its instruction-cache footprint) small.
template <class T, int NT, class Sink>
GWF_HD void phenomhm_foreach_mode(const HMRec<NT>& r, int g, double f, bool apply_cut, Sink& sink) {
    typedef Ld<T> L;
    const T x = L::get(r.s) * f;
    const double xv = val(x);
    const T eta = L::get(r.eta), seta = L::get(r.seta), chis = L::get(r.chis), chia = L::get(r.chia);
    const T x13 = dcbrt(x);
    const T lin = L::get(r.t0) * (x - L::get(r.xref[g]));
#pragma unroll 1
    for (int m = 0; m < kHMModes; ++m) {
        const HMMode<Dual<NT>>& o = r.mode[m];
        const double mm = hm_mm(m), ai = 2. / mm;
        // amplitude: completeAmpl(fS) * (beta1/beta2) * HMamp1/HMamp2, waveforms.py:2569-2582
        T y;
        if (xv < o.fi_amp.v) y = x * ai;
        else if (xv < o.fr.v) y = x * L::get(o.am_amp) + L::get(o.bm_amp);
        else y = x + L::get(o.br_amp);
        const T y13 = dcbrt(y);
        const T shape = hm_amp_shape<T, NT>(r, y, y13, apply_cut);
        const double c2pm = cbrt(2. * kPi / mm);
        const T v1 = c2pm * x13, v2 = v1 * cbrt(ai), vS = c2pm * y13;
        const T h2 = hm_absH(m, v2, eta, seta, chis, chia);
        T A(0.0);
        if (val(h2) != 0.0 && val(shape) != 0.0) {            // nan_to_num of 0/0 (waveforms.py:2582)
            const T h1 = hm_absH(m, v1, eta, seta, chis, chia), hS = hm_absH(m, vS, eta, seta, chis, chia);
            const T ym76 = 1.0 / (y * dsqrt(y13));
            A = L::get(r.Camp) * ym76 * shape * (h1 * hS / h2);
        }
        // phase, waveforms.py:2600-2607: offset + completePhase(mapped frequency) * scale
        T yp, off, scale;
        if (xv < o.fi_phi.v) { yp = x * ai; off = T(0.0); scale = T(1.0 / ai); }
        else if (xv < o.fr.v) { yp = x * L::get(o.am_phi) + L::get(o.bm_phi); off = L::get(o.kB); scale = L::get(o.inv_am_phi); }
        else { yp = x * L::get(o.rho); off = L::get(o.kC); scale = L::get(o.inv_rho); }
        const T cph = hm_complete_phase<T, NT>(r, yp, L::get(o.c1), L::get(o.c2), L::get(o.at_c), L::get(o.at_a), L::get(o.at_w), apply_cut);
        sink(m, A, off + cph * scale - lin - mm * L::get(r.phi0[g]) + hm_shift((int)mm));
    }
}

template <class T> struct HMStoreSink {
    T* amp;
    T* phase;
    GWF_HD void operator()(int m, const T& A, const T& ph) { amp[m] = A; phase[m] = ph; }
};
template <class T, int NT>
GWF_HD void phenomhm_amp_phase(const HMRec<NT>& r, int g, double f, bool apply_cut, T* amp, T* phase) {
    HMStoreSink<T> sink = {amp, phase};
    phenomhm_foreach_mode<T, NT>(r, g, f, apply_cut, sink);
}

}  // namespace gwf
