// IMRPhenomHM (gwfast/waveforms.py:1838-2749): six modes (21, 22, 32, 33, 43, 44), each an IMRPhenomD amplitude and
// phase evaluated at a piecewise-linearly mapped frequency, rescaled with 1.5PN amplitude ratios and combined with
// spin-weighted spherical harmonics into h+ and hx (hphc, waveforms.py:2256-2616).
//
// Unlike the (2,2)-only models no common phase factors out of the strain, so the phase VALUE of every mode is needed at every
// frequency.  The work is split like IMRPhenomD's (model_phenomd.cuh):
//   * the per-event prologue evaluates every f-independent quantity in dual arithmetic (value + tangents w.r.t. Mc, eta, chi1z,
//     chi2z; iota enters only through the harmonics and is differentiated in closed form) and leaves a COEFFICIENT RECORD:
//     the PhenomD phase/amplitude coefficient rows shared by all modes, and per mode the frequency maps y = a x + b of the three
//     regimes (with the phase offset and scale that go with them), the merger-ringdown constants and the amplitude prefactor;
//   * the per-frequency code has no dual arithmetic: for a mode, y and d y/d p_j follow from the map rows, the PhenomD region
//     that holds y is a sum over basis functions b_k(y) against the coefficient rows (value, y d/dy, and the NT tangents of the
//     coefficients), and the chain rule adds (y d/dy) * (d ln y/d p_j).  The 1.5PN amplitude ratios beta(x)/beta(2x/m) HM(y) of
//     the power-law modes collapse to |c g(eta)| (pi y)^(p/3) (the x-dependence cancels), the (2,1) mode keeps its complex
//     polynomial in v.
// QNM frequencies come from the complex polynomial fits of _RDfreqCalc (waveforms.py:2661-2710), not from the tables.
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

// ---- prologue-side: every f-independent PhenomD quantity as a dual, and the functions of the mapped frequency the prologue
// itself has to evaluate (continuity constants, reference phase)
template <int NT>
struct HMCore {
    typedef Dual<NT> D;
    double x_mrd, x_peak;             // values of fMRDJoin and fpeak (region tests)
    D pins[kPIns];                    // inspiral phase coefficients (same basis as PhenomDRec, -t0 NOT folded in)
    D pint[4];                        // C1Int, beta1/eta + C2Int, -beta3/(3 eta), beta2/eta
    D pmrd[3];                        // -alpha2/eta, 4/3 alpha3/eta, alpha1/eta
    D ains[kAIns], aint[kAInt], amrd[4];
};

// PhenomD phase at mapped frequency y for one mode (completePhase, waveforms.py:2023-2027)
template <int NT>
GWF_HD Dual<NT> hm_complete_phase(const HMCore<NT>& r, const Dual<NT>& y, const Dual<NT>& c1, const Dual<NT>& c2, const Dual<NT>& at_c,
                                  const Dual<NT>& at_a, const Dual<NT>& at_w, bool apply_cut) {
    typedef Dual<NT> T;
    const double yv = y.v;
    if (yv < kPhiJoinIns) {
        const T y13 = dcbrt(y), y23 = y13 * y13, lg = dlog(kPi * y) * (1. / 3.);
        const T ym13 = 1.0 / y13, ym23 = ym13 * ym13, ym1 = ym23 * ym13;
        return r.pins[0] + r.pins[1] * y23 + r.pins[2] * y13 + r.pins[3] * (y13 * lg) + r.pins[4] * lg + r.pins[5] * ym13 + r.pins[6] * ym23 +
               r.pins[7] * ym1 + r.pins[8] * (ym1 * ym23) + r.pins[9] * y + r.pins[10] * (y * y13) + r.pins[11] * (y * y23) + r.pins[12] * (y * y);
    }
    if (yv < r.x_mrd) {
        const T ym1 = 1.0 / y;
        return r.pint[0] + r.pint[1] * y + r.pint[2] * (ym1 * ym1 * ym1) + r.pint[3] * dlog(y);
    }
    if (!apply_cut || yv < kMfCut) {
        const T sy = dsqrt(y);
        return r.pmrd[0] / y + r.pmrd[1] * (sy * dsqrt(sy)) + r.pmrd[2] * y + at_c * datan((y - at_a) / at_w) + c1 + c2 * y;
    }
    return T(0.0);
}

// ---- coefficient record
// power p of v in |H_lm(v)| of the power-law modes (waveforms.py:2503-2515): 22: 0, 32: 2, 33: 1, 43: 3, 44: 2; (2,1) is a polynomial
GWF_HD constexpr int hm_vpow(int m) { return m == 1 ? 0 : (m == 2 ? 2 : (m == 3 ? 1 : (m == 4 ? 3 : (m == 5 ? 2 : -1)))); }

// per mode; every row is (value, NT tangents), padded to RowLen<NT>
template <int NT>
struct HMModeRec {
    static constexpr int RL = RowLen<NT>::v;
    double fi_amp, fi_phi, fr;         // regime break points of the two maps (0.014/Rho, 0.018/Rho, fring_lm): values only
    double k13[2], kln[2], k34[2];     // regimes whose map is a pure scaling y = k x: k^(1/3) and ln k of the amplitude / phase map of
                                       // regime 0 (k = 2/m), and of the phase map of regime 2 (k = Rho): [0] = 2/m, [1] = Rho; k34 = k^(3/4)
    double G, lnG_d[NT];               // power-law modes: A = G y^(p/3 - 7/6) ampIMR(y), G = Camp |c g(eta)| pi^(p/3)
    alignas(16) double amap[3][2][RL]; // amplitude map of the three regimes: a, b          (y = a x + b)
    alignas(16) double pmap[3][4][RL]; // phase map: a, b, offset, scale                     (Phi = offset + scale * completePhase(y))
    alignas(16) double mrd[5][RL];     // C1MRDHM, C2MRDHM, alpha4 Rho/eta, alpha5 fring, fdamp Rho Tau
};

template <int NT>
struct HMRec {
    static constexpr int RL = RowLen<NT>::v;
    double s;                          // x = s f, s = M GMsun/c^3
    ScalePow sp;
    double ln_s;
    double lam[NT];                    // d ln s
    double fcut_hz;
    double x_mrd, x_peak;
    double Camp, lnCamp_d[NT];         // M^2 GMsun_c2_Gpc GMsun_c3/dL * amp0   (no 2 sqrt(5/64pi): waveforms.py:2172-2174)
    double t0[1 + NT];
    double pc0[kMaxGroups][1 + NT];    // t0 * xRef per grid group
    double phi0[kMaxGroups][1 + NT];   // completePhase_22(xRef)/2
    double h21[5][1 + NT];             // (2,1) amplitude polynomial: r1, r2, r3, r4 (real part, powers v..v^4) and i4 (imaginary, v^4)
    double amrd[4][1 + NT];            // fring, gamma2/(fdamp gamma3), fdamp gamma3, fdamp gamma3 gamma1
    alignas(16) double pins[kPIns][RL];
    alignas(16) double pint[4][RL];
    alignas(16) double pmrd[3][RL];
    alignas(16) double ains[kAIns][RL];
    alignas(16) double aint[kAInt][RL];
    HMModeRec<NT> mode[kHMModes];
    TauRec tau;
};
static_assert(sizeof(HMRec<4>) % sizeof(double) == 0, "records are copied as doubles");

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
template <int NT> GWF_HD void put_row(double* dst, const Dual<NT>& c) {
    put(dst, c);
    for (int k = 1 + NT; k < RowLen<NT>::v; ++k) dst[k] = 0.0;
}

template <int NT>
GWF_HD void phenomhm_prologue(HMRec<NT>& r, const Intrinsic<NT>& p, double dL, const double* fmin_g, int ngroups, const ModelCfg& cfg,
                              double s_host = 0.0, double fcut_host = 0.0) {
    typedef Dual<NT> D;
    const bool apply_cut = !(cfg.flags & kFlagNoFcut);
    const D eta = p.eta;
    const D M = p.Mc / dpow(eta, 3. / 5.);
    D s = M * kGMsunC3;
    if (s_host > 0.0) s.v = s_host;       // the host's M*GMsun_over_c3: x = s f then rounds like the reference's fgrid
    r.s = s.v;
    r.sp.set(s.v);
    r.ln_s = log(s.v);
#pragma unroll
    for (int j = 0; j < NT; ++j) r.lam[j] = s.d[j] / s.v;
    r.fcut_hz = fcut_host > 0.0 ? fcut_host : kMfCut / s.v;    // waveforms.py:2737-2749
    const D seta = seta_of(eta);
    const D chis = 0.5 * (p.chi1 + p.chi2), chia = 0.5 * (p.chi1 - p.chi2);
    const D aeff = final_spin(eta, p.chi1, p.chi2), finMass = 1. - radiated_energy(eta, p.chi1, p.chi2);
    D fring, fdamp;
    qnm_freqs(1, aeff, finMass, fring, fdamp);                 // (2,2), waveforms.py:2336-2345
    PhenomDCore<NT> c;
    c.build_rd(eta, p.chi1, p.chi2, D(1.0), D(1.0), fring, fdamp);
    HMCore<NT> h;
    h.x_mrd = r.x_mrd = c.fMRDJoin.v;
    h.x_peak = r.x_peak = c.fpeak_amp.v;
    const D amp0 = dsqrt(2.0 * eta / 3.0) * pow(kPi, -1. / 6.);
    const D Camp = M * kGMsunC2Gpc * M * kGMsunC3 / D(dL) * amp0;
    r.Camp = Camp.v;
#pragma unroll
    for (int j = 0; j < NT; ++j) r.lnCamp_d[j] = Camp.d[j] / Camp.v;
    const D t0 = c.dphi_mrd(c.fpeak_amp);                      // waveforms.py:2527
    put(r.t0, t0);
    const double cp = cbrt(kPi), cp2 = cp * cp;
    const D n = c.norm, ie = 1.0 / eta;
    h.pins[0] = c.pn.c5 * n; h.pins[1] = c.pn.c7 * n * cp2; h.pins[2] = c.pn.c6 * n * cp; h.pins[3] = (-6848. / 21.) * n * cp;
    h.pins[4] = 3. * c.pn.c5 * n; h.pins[5] = c.pn.c4 * n / cp; h.pins[6] = c.pn.c3 * n / cp2; h.pins[7] = c.pn.c2 * n / kPi;
    h.pins[8] = n / (kPi * cp2); h.pins[9] = c.fit[SIG1] * ie; h.pins[10] = c.fit[SIG2] * 0.75 * ie; h.pins[11] = c.fit[SIG3] * 0.6 * ie;
    h.pins[12] = c.fit[SIG4] * 0.5 * ie;
    h.pint[0] = c.C1Int; h.pint[1] = c.fit[BET1] * ie + c.C2Int; h.pint[2] = -c.fit[BET3] * ie / 3.; h.pint[3] = c.fit[BET2] * ie;
    h.pmrd[0] = -c.fit[ALP2] * ie; h.pmrd[1] = (4.0 / 3.0) * c.fit[ALP3] * ie; h.pmrd[2] = c.fit[ALP1] * ie;
    h.ains[0] = D(1.0);
    for (int k = 1; k < kAIns; ++k) h.ains[k] = c.A[k + 1];
    for (int k = 0; k < kAInt; ++k) h.aint[k] = c.e[k];
    const D fd3 = fdamp * c.fit[GAM3];
    h.amrd[0] = fring; h.amrd[1] = c.fit[GAM2] / fd3; h.amrd[2] = fd3; h.amrd[3] = fd3 * c.fit[GAM1];
    for (int k = 0; k < kPIns; ++k) put_row(r.pins[k], h.pins[k]);
    for (int k = 0; k < 4; ++k) put_row(r.pint[k], h.pint[k]);
    for (int k = 0; k < 3; ++k) put_row(r.pmrd[k], h.pmrd[k]);
    for (int k = 0; k < kAIns; ++k) put_row(r.ains[k], h.ains[k]);
    // the intermediate amplitude is stored in powers of (y - 0.014) like IMRPhenomD's
    for (int k = 0; k < kAInt; ++k) put_row(r.aint[k], h.aint[k]);
    for (int k = 0; k < 4; ++k) put(r.amrd[k], h.amrd[k]);
    // (2,1) amplitude polynomial H_21(v) = sqrt(2)/3 v [r1 + r2 v + r3 v^2 + (r4 + i i4) v^3], waveforms.py:2506
    put(r.h21[0], seta);
    put(r.h21[1], -1.5 * (chia + seta * chis));
    put(r.h21[2], seta * ((335.0 / 672.0) + (eta * 117.0 / 56.0)));
    put(r.h21[3], chia * (3427.0 / 1344. - eta * 2101.0 / 336.) + seta * chis * (3427.0 / 1344 - eta * 965. / 336.) + seta * (-kPi));
    put(r.h21[4], seta * (-0.5 - 2 * 0.69314718056));
    // reference phase of the (2,2) mode: completePhase(fRef, C1MRD, C2MRD, 1, 1)/2, waveforms.py:2529-2530
    const D at_c22 = c.fit[ALP4] * ie, at_a22 = c.fit[ALP5] * fring;
    for (int g = 0; g < ngroups; ++g) {
        const D xref = (cfg.flags & kFlagHasFRef) ? s * cfg.fRef : s * fmin_g[g];
        put(r.pc0[g], t0 * xref);
        put(r.phi0[g], 0.5 * hm_complete_phase<NT>(h, xref, c.C1MRD, c.C2MRD, at_c22, at_a22, fdamp, apply_cut));
    }
    // per-mode quantities, waveforms.py:2534-2605
    const D fm = c.fMRDJoin;
    const D PhiIntTempVal = c.phi_int_raw(fm) / eta + c.C1Int + c.C2Int * fm;
    const D DPhiIntTempVal = c.C2Int + c.dphi_int(fm);
    for (int m = 0; m < kHMModes; ++m) {
        HMModeRec<NT>& o = r.mode[m];
        const double mm = hm_mm(m);
        D frlm, fdlm;
        qnm_freqs(m, aeff, finMass, frlm, fdlm);
        const D Rho = fring / frlm, Tau = fdlm / fdamp;
        const D u = fm - c.fit[ALP5] * fring;
        const D wtr = fdamp * Tau * Rho;
        const D DPhiMRDVal = (c.fit[ALP1] + c.fit[ALP2] / (fm * fm) + c.fit[ALP3] / dpow(fm, 0.25) + c.fit[ALP4] / (fdamp * Tau * (1. + u * u / (wtr * wtr)))) / eta;
        const D PhiMRJoinTemp = -(c.fit[ALP2] / fm) + (4.0 / 3.0) * (c.fit[ALP3] * dpow(fm, 0.75)) + c.fit[ALP1] * fm + c.fit[ALP4] * Rho * datan(u / wtr);
        const D c2 = DPhiIntTempVal - DPhiMRDVal;
        const D c1 = PhiIntTempVal - PhiMRJoinTemp / eta - c2 * fm;
        const D at_c = c.fit[ALP4] * Rho * ie, at_a = c.fit[ALP5] * fring, at_w = wtr;
        put_row(o.mrd[0], c1); put_row(o.mrd[1], c2); put_row(o.mrd[2], at_c); put_row(o.mrd[3], at_a); put_row(o.mrd[4], at_w);
        o.fr = frlm.v;
        const double ai = 2. / mm;
        // amplitude map, waveforms.py:2547-2563: regime 0: (2/m) x; 1: am x + bm; 2: x + (fring - fring_lm)
        const D fi_amp = kAmpJoinIns / Rho;
        o.fi_amp = fi_amp.v;
        {
            const D Trd = frlm - frlm + fring, Ti = 2. * fi_amp / mm;
            const D am = (Trd - Ti) / (frlm - fi_amp);
            put_row(o.amap[0][0], D(ai)); put_row(o.amap[0][1], D(0.0));
            put_row(o.amap[1][0], am); put_row(o.amap[1][1], Ti - fi_amp * am);
            put_row(o.amap[2][0], D(1.0)); put_row(o.amap[2][1], -frlm + fring);
        }
        // phase map: regime 0: (2/m) x, Phi = cP(y) m/2; 1: am x + bm, Phi = kB + cP(y)/am; 2: Rho x, Phi = kC + cP(y)/Rho
        const D fi_phi = kPhiJoinIns / Rho;
        o.fi_phi = fi_phi.v;
        const D Trd = frlm * Rho, Ti = 2. * fi_phi / mm;
        const D am_phi = (Trd - Ti) / (frlm - fi_phi);
        const D bm_phi = Ti - fi_phi * am_phi;
        // continuity constants, waveforms.py:2586-2596
        auto cP = [&](const D& y) { return hm_complete_phase<NT>(h, y, c1, c2, at_c, at_a, at_w, apply_cut); };
        const D PhDBconst = cP(am_phi * fi_phi + bm_phi) / am_phi;
        const D PhDCconst = cP(Rho * frlm) / Rho;
        const D PhDBAterm = cP(ai * fi_phi) / ai;
        const D tmpphaseC = -PhDBconst + PhDBAterm + cP(am_phi * frlm + bm_phi) / am_phi;
        put_row(o.pmap[0][0], D(ai)); put_row(o.pmap[0][1], D(0.0)); put_row(o.pmap[0][2], D(0.0)); put_row(o.pmap[0][3], D(1.0 / ai));
        put_row(o.pmap[1][0], am_phi); put_row(o.pmap[1][1], bm_phi); put_row(o.pmap[1][2], -PhDBconst + PhDBAterm); put_row(o.pmap[1][3], 1.0 / am_phi);
        put_row(o.pmap[2][0], Rho); put_row(o.pmap[2][1], D(0.0)); put_row(o.pmap[2][2], -PhDCconst + tmpphaseC); put_row(o.pmap[2][3], 1.0 / Rho);
        o.k13[0] = cbrt(ai); o.kln[0] = log(ai);
        o.k13[1] = cbrt(Rho.v); o.kln[1] = log(Rho.v);
        o.k34[0] = sqrt(ai) * sqrt(sqrt(ai)); o.k34[1] = sqrt(Rho.v) * sqrt(sqrt(Rho.v));
        // amplitude prefactor of the power-law modes: beta(x)/beta(2x/m) HM(y) = |c g| (pi y)^(p/3) (waveforms.py:2503-2515, 2569-2582)
        D g(1.0);
        if (m == 2) g = (1.0 / 3.0) * sqrt(5.0 / 7.0) * (1.0 - 3.0 * eta);
        else if (m == 3) g = 0.75 * sqrt(5.0 / 7.0) * seta;
        else if (m == 4) g = 0.75 * sqrt(3.0 / 35.0) * seta * (1.0 - 2.0 * eta);
        else if (m == 5) g = (4.0 / 9.0) * sqrt(10.0 / 7.0) * (1.0 - 3.0 * eta);
        g = dfabs(g);
        const int pw = hm_vpow(m);
        const D G = Camp * g * (pw > 0 ? pow(kPi, pw / 3.0) : 1.0);
        o.G = m == 0 ? Camp.v : G.v;
#pragma unroll
        for (int j = 0; j < NT; ++j) o.lnG_d[j] = (m == 0 || G.v == 0.0) ? r.lnCamp_d[j] : G.d[j] / G.v;
    }
    tau_fill(r.tau, s, eta);
}

// ------------------------------------------------------------------------------------------------ per frequency
// powers of x shared by the six modes
struct HMX {
    double x, x13, lnx, x34, lpx3;
    GWF_HD void set(double s, const ScalePow& sp, double ln_s, const FreqPoint& fp) {
        x = s * fp.f;
        x13 = sp.s13 * fp.f13;
        lnx = ln_s + fp.lnf;
        lpx3 = fma(fp.lnf, 1. / 3., sp.lps3);
        const double sx = sqrt(x);
        x34 = sx * sqrt(sx);
    }
};

#ifndef GWF_HM_UNROLL
#define GWF_HM_UNROLL 1
#endif
constexpr int kHmModeUnroll = GWF_HM_UNROLL;      // the mode loop stays rolled (experiment knob: see DESIGN.md, tried and rejected)

// Amplitude A_lm and phase Phi_lm of the six modes (IMRPhenomHM.Ampl / .Phi, waveforms.py:1874-2254, as vectorised in hphc) at one
// frequency, with (TAN) d ln A_lm and d Phi_lm w.r.t. the NT intrinsic slots.  Every mode is handed to `sink(m, A, Phi, dlnA, dPhi)`
// as soon as it is known, so callers that only need sums over the modes keep nothing per mode.  The mode loop is rolled: the six
// modes run the same code on different rows of the record (the kernel's instruction footprint stays that of one mode).
template <int NT, bool TAN, class Sink>
GWF_HD void phenomhm_foreach_mode(const HMRec<NT>& r, int g, const FreqPoint& fp, bool apply_cut, Sink& sink, int m_begin = 0, int m_end = kHMModes) {
    HMX X;
    X.set(r.s, r.sp, r.ln_s, fp);
    const double x = X.x;
    // part of the phase common to the modes: -t0 (x - xRef), waveforms.py:2609
    const double L0 = fma(-r.t0[0], x, r.pc0[g][0]);
    double dL0[NT];
    if (TAN) {
#pragma unroll
        for (int j = 0; j < NT; ++j) dL0[j] = fma(-fma(r.t0[0], r.lam[j], r.t0[1 + j]), x, r.pc0[g][1 + j]);
    }
#pragma unroll(kHmModeUnroll)
    for (int m = m_begin; m < m_end; ++m) {
        const HMModeRec<NT>& o = r.mode[m];
        const int mm = hm_mm(m);
        double A = 0.0, dlnA[NT];
#pragma unroll
        for (int j = 0; j < NT; ++j) dlnA[j] = 0.0;
        // ---------------- amplitude: completeAmpl(y) * (beta1/beta2) * HMamp1/HMamp2, waveforms.py:2569-2582
        {
            const int ra = (x >= o.fi_amp ? 1 : 0) + (x >= o.fr ? 1 : 0);
            const double* Ma = o.amap[ra][0];
            const double* Mb = o.amap[ra][1];
            const double y = fma(Ma[0], x, Mb[0]);
            const double iy = rcp_fast(y);
            double dlny[NT];
            if (TAN) {
#pragma unroll
                for (int j = 0; j < NT; ++j) dlny[j] = fma(fma(Ma[0], r.lam[j], Ma[1 + j]), x, Mb[1 + j]) * iy;
            }
            const double y13 = ra == 0 ? X.x13 * o.k13[0] : cbrt(y);
            // ampIMR(y): value v, y dv/dy (dx) and the coefficient tangents d[j]
            double v = 0., dx = 0., d[NT];
#pragma unroll
            for (int j = 0; j < NT; ++j) d[j] = 0.;
            bool on = true;
            if (y < kAmpJoinIns) {
                const double y23 = y13 * y13, y43 = y * y13, y53 = y * y23, y2 = y * y;
                const double b[kAIns] = {1., y23, y, y43, y53, y2, y2 * y13, y2 * y23, y2 * y};
                const double bx[kAIns] = {0., 2. / 3. * y23, y, 4. / 3. * y43, 5. / 3. * y53, 2. * y2, 7. / 3. * b[6], 8. / 3. * b[7], 3. * b[8]};
                expand<kAIns, NT>(r.ains, b, bx, v, d, dx);
            } else if (y < r.x_peak) {
                const double u = y - kAmpJoinIns, u2 = u * u;
                const double b[kAInt] = {1., u, u2, u2 * u, u2 * u2};
                const double bx[kAInt] = {0., y, 2. * y * u, 3. * y * u2, 4. * y * u2 * u};
                expand<kAInt, NT>(r.aint, b, bx, v, d, dx);
            } else if (!apply_cut || y < kMfCut) {
                // exp(-(y-fr) g) * P / ((y-fr)^2 + w^2);  amrd = {fr, g, w, P}
                const double u = y - r.amrd[0][0], gg = r.amrd[1][0], w = r.amrd[2][0], P = r.amrd[3][0];
                const double iden = rcp_fast(u * u + w * w), iP = rcp_fast(P);
                v = exp(-u * gg) * P * iden;
                const double ku = -gg - 2. * u * iden;
                dx = v * ku * y;
                if (TAN) {
#pragma unroll
                    for (int j = 0; j < NT; ++j)
                        d[j] = v * (-ku * r.amrd[0][1 + j] - u * r.amrd[1][1 + j] + r.amrd[3][1 + j] * iP - 2. * w * iden * r.amrd[2][1 + j]);
                }
            } else on = false;
            if (on && v != 0.0) {
                const double ym76 = iy * rsqrt(y13);
                const double iv = rcp_fast(v);
                if (m == 0) {
                    // (2,1): |H(v1)| |H(vS)| / |H(v2)| with v1 = (2 pi x)^(1/3), v2 = (4 pi x)^(1/3), vS = (2 pi y)^(1/3)
                    const double c2p = 1.84527014864402841909680387958898802678;      // (2 pi)^(1/3)
                    const double c4p = 2.324894703019252951141799628955674133719;      // (4 pi)^(1/3)
                    const double vv[3] = {c2p * X.x13, c4p * X.x13, c2p * y13};
                    double habs[3], dlnh[3][NT];
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        const double vq = vv[q], v2 = vq * vq, v3 = v2 * vq;
                        const double Pq = fma(r.h21[3][0], v3, fma(r.h21[2][0], v2, fma(r.h21[1][0], vq, r.h21[0][0])));
                        const double Qq = r.h21[4][0] * v3;
                        const double n2 = Pq * Pq + Qq * Qq;
                        habs[q] = (1.4142135623730951 / 3.0) * vq * sqrt(n2);
                        if (TAN && n2 != 0.0) {
                            const double in2 = rcp_fast(n2);
                            const double vPv = fma(3.0 * r.h21[3][0], v3, fma(2.0 * r.h21[2][0], v2, r.h21[1][0] * vq));   // v dP/dv
                            const double vQv = 3.0 * Qq;
                            // d ln v: lam/3 for v1, v2; dlny/3 for vS
#pragma unroll
                            for (int j = 0; j < NT; ++j) {
                                const double dlv = (q == 2 ? dlny[j] : r.lam[j]) * (1. / 3.);
                                const double dP = fma(r.h21[3][1 + j], v3, fma(r.h21[2][1 + j], v2, fma(r.h21[1][1 + j], vq, r.h21[0][1 + j]))) + vPv * dlv;
                                const double dQ = fma(r.h21[4][1 + j], v3, vQv * dlv);
                                dlnh[q][j] = dlv + (Pq * dP + Qq * dQ) * in2;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < NT; ++j) dlnh[q][j] = 0.0;
                        }
                    }
                    if (habs[1] != 0.0) {                     // nan_to_num of 0/0 (waveforms.py:2582)
                        A = r.Camp * ym76 * v * (habs[0] * habs[2] * rcp_fast(habs[1]));
                        if (TAN) {
#pragma unroll
                            for (int j = 0; j < NT; ++j)
                                dlnA[j] = r.lnCamp_d[j] - (7. / 6.) * dlny[j] + fma(dx, dlny[j], d[j]) * iv + dlnh[0][j] + dlnh[2][j] - dlnh[1][j];
                        }
                    }
                } else {
                    const int pw = hm_vpow(m);
                    const double ypw = pw == 0 ? 1.0 : (pw == 1 ? y13 : (pw == 2 ? y13 * y13 : y));
                    A = o.G * ym76 * ypw * v;
                    if (TAN) {
                        const double ex = pw * (1. / 3.) - 7. / 6.;
#pragma unroll
                        for (int j = 0; j < NT; ++j) dlnA[j] = o.lnG_d[j] + ex * dlny[j] + fma(dx, dlny[j], d[j]) * iv;
                    }
                }
            }
        }
        // ---------------- phase, waveforms.py:2600-2609: offset + scale * completePhase(mapped frequency) - t0 (x - xRef) - m phi0 + shift
        double Phi, dPhi[NT];
        {
            const int rp = (x >= o.fi_phi ? 1 : 0) + (x >= o.fr ? 1 : 0);
            const double* Ma = o.pmap[rp][0];
            const double* Mb = o.pmap[rp][1];
            const double* Mo = o.pmap[rp][2];
            const double* Ms = o.pmap[rp][3];
            const double y = fma(Ma[0], x, Mb[0]);
            const double iy = rcp_fast(y);
            double dlny[NT];
            if (TAN) {
#pragma unroll
                for (int j = 0; j < NT; ++j) dlny[j] = fma(fma(Ma[0], r.lam[j], Ma[1 + j]), x, Mb[1 + j]) * iy;
            }
            const int ks = rp == 2 ? 1 : 0;                    // which scaling constants apply when the map is y = k x
            double v = 0., dx = 0., d[NT];
#pragma unroll
            for (int j = 0; j < NT; ++j) d[j] = 0.;
            if (y < kPhiJoinIns) {
                double y13, L;
                if (rp == 1) { y13 = cbrt(y); L = log(kPi * y) * (1. / 3.); }
                else { y13 = X.x13 * o.k13[ks]; L = fma(o.kln[ks], 1. / 3., X.lpx3); }
                const double ym13 = 1.0 / y13, y23 = y13 * y13, ym23 = ym13 * ym13, ym1 = ym23 * ym13, ym53 = ym1 * ym23;
                const double y43 = y * y13, y53 = y * y23, y2 = y * y, y13L = y13 * L;
                const double b[kPIns] = {1., y23, y13, y13L, L, ym13, ym23, ym1, ym53, y, y43, y53, y2};
                const double bx[kPIns] = {0., 2. / 3. * y23, 1. / 3. * y13, 1. / 3. * (y13L + y13), 1. / 3., -1. / 3. * ym13, -2. / 3. * ym23, -ym1,
                                          -5. / 3. * ym53, y, 4. / 3. * y43, 5. / 3. * y53, 2. * y2};
                expand<kPIns, NT>(r.pins, b, bx, v, d, dx);
            } else if (y < r.x_mrd) {
                const double ly = rp == 1 ? log(y) : X.lnx + o.kln[ks];
                const double ym3 = iy * iy * iy;
                const double b[4] = {1., y, ym3, ly};
                const double bx[4] = {0., y, -3. * ym3, 1.};
                expand<4, NT>(r.pint, b, bx, v, d, dx);
            } else if (!apply_cut || y < kMfCut) {
                double y34;
                if (rp == 1) { const double sy = sqrt(y); y34 = sy * sqrt(sy); }
                else y34 = X.x34 * o.k34[ks];
                const double b[3] = {iy, y34, y};
                const double bx[3] = {-iy, 0.75 * y34, y};
                expand<3, NT>(r.pmrd, b, bx, v, d, dx);
                // + at_c atan((y - at_a)/at_w) + c1 + c2 y
                const double* c1 = o.mrd[0];
                const double* c2 = o.mrd[1];
                const double* tc = o.mrd[2];
                const double* ta = o.mrd[3];
                const double* tw = o.mrd[4];
                const double iw = rcp_fast(tw[0]);
                const double u = (y - ta[0]) * iw;
                const double at = atan(u), wq = tc[0] * rcp_fast(1.0 + u * u) * iw;     // at_c d(atan)/du / at_w
                v += fma(tc[0], at, fma(c2[0], y, c1[0]));
                dx += (wq + c2[0]) * y;
                if (TAN) {
#pragma unroll
                    for (int j = 0; j < NT; ++j) d[j] += fma(tc[1 + j], at, fma(c2[1 + j], y, c1[1 + j])) - wq * (ta[1 + j] + u * tw[1 + j]);
                }
            }
            Phi = fma(Ms[0], v, Mo[0]) + L0 - mm * r.phi0[g][0] + hm_shift(mm);
            if (TAN) {
#pragma unroll
                for (int j = 0; j < NT; ++j)
                    dPhi[j] = fma(Ms[1 + j], v, Mo[1 + j]) + Ms[0] * fma(dx, dlny[j], d[j]) + dL0[j] - mm * r.phi0[g][1 + j];
            }
        }
        sink(m, A, Phi, dlnA, dPhi);
    }
}

struct HMStoreSink {
    double* amp;
    double* phase;
    GWF_HD void operator()(int m, double A, double ph, const double*, const double*) { amp[m] = A; phase[m] = ph; }
};
// value-only amplitudes and phases of the six modes (WaveFormModel.Ampl / .Phi on a user grid)
template <int NT>
GWF_HD void phenomhm_amp_phase(const HMRec<NT>& r, int g, const FreqPoint& fp, bool apply_cut, double* amp, double* phase) {
    HMStoreSink sink = {amp, phase};
    phenomhm_foreach_mode<NT, false>(r, g, fp, apply_cut, sink);
}

}  // namespace gwf
