// Per-(event, frequency) work of the fused Fisher/SNR path, independent of the thread mapping:
// frequency grid + trapezoid weights (gwfast/signal.py:715-723, 884-901, 929), model dispatch, detector loop.
#pragma once
#include <type_traits>
#include "detector.cuh"
#include "model_phenomd.cuh"
#include "model_tf2.cuh"
#include "model_nrtidal.cuh"
#include "model_phenomhm.cuh"
#include "model_nsbh.cuh"

namespace gwf {

// ------------------------------------------------------------------ model traits
template <int MODEL, int NT> struct ModelTraits;

template <int NT> struct ModelTraits<kTaylorF2, NT> {
    typedef TF2Rec<NT> Rec;
    static GWF_HD void eval_amp(const Rec& r, const ModelCfg&, const FreqPoint& fp, bool need_tau, double& A, double& tau) {
        A = r.C * fp.fm76;
        tau = 0.;
        if (need_tau) {
            VPow p;
            p.set(r.sp, fp);
            double dtau[2];
            tau_eval(r.tau, p.vm1, p.lpx3, r.lam, tau, dtau);
        }
    }
    static GWF_HD void prologue(Rec& r, const EventIn& e, const ModelCfg& cfg, int opt_flags, const QnmTables&, const double* fmin_g, int ng, int = 3) {
        const Intrinsic<NT> p = seed_intrinsic<NT>(e, opt_flags, (cfg.flags & kFlagTidal) != 0);
        tf2_prologue(r, p, e.dL, cfg, e.fcut_host, fmin_g, ng);
    }
    // waveform at f: amplitude, d ln A, d Phi and (if need_tau) d t_noloc
    static GWF_HD void eval(const Rec& r, const ModelCfg&, int g, const FreqPoint& fp, bool need_tau, PointWf<NT>& w) {
        VPow p;
        p.set(r.sp, fp);
        tf2_phase(r, p, w.phi, w.phi_d, g);
        w.A = r.C * fp.fm76;
#pragma unroll
        for (int j = 0; j < NT; ++j) w.lnA_d[j] = r.lnC_d[j];
        w.dtn[0] = w.dtn[1] = 0.;
        w.tau = 0.;
        if (need_tau) {
            double tau, dtau[2];
            tau_eval(r.tau, p.vm1, p.lpx3, r.lam, tau, dtau);
            w.tau = tau;
            w.dtn[0] = -dtau[0] * kInvDay;
            w.dtn[1] = -dtau[1] * kInvDay;
        }
    }
};

template <int NT> struct ModelTraits<kPhenomD, NT> {
    typedef PhenomDRec<NT> Rec;
    // amplitude (and tau) only: the SNR needs neither the phase nor any tangent
    static GWF_HD void eval_amp(const Rec& r, const ModelCfg& cfg, const FreqPoint& fp, bool need_tau, double& A, double& tau) {
        XPow p;
        p.set(r.s, r.sp, fp);
        A = r.C76 * fp.fm76 * phenomd_amp_value(r, p, !(cfg.flags & kFlagNoFcut));
        tau = 0.;
        if (need_tau) {
            double dtau[2];
            tau_eval(r.tau, p.xm13 * 0.68278406325529568146702083315816, p.lpx3, r.lam, tau, dtau);
        }
    }
    // parts: the phase (1) and amplitude (2) halves of the record can be computed independently (PhenomDCore::build)
    static GWF_HD void prologue(Rec& r, const EventIn& e, const ModelCfg& cfg, int opt_flags, const QnmTables& q, const double* fmin_g, int ng, int parts = 3) {
        const Intrinsic<NT> p = seed_intrinsic<NT>(e, opt_flags, false);
        phenomd_prologue(r, p, e.dL, q, fmin_g, ng, cfg, e.s_host, e.fcut_host, parts);
    }
    static GWF_HD void eval(const Rec& r, const ModelCfg& cfg, int g, const FreqPoint& fp, bool need_tau, PointWf<NT>& w) {
        XPow p;
        p.set(r.s, r.sp, fp);
        const bool cut = !(cfg.flags & kFlagNoFcut);
        phenomd_phase(r, g, p, cut, w.phi, w.phi_d);
        phenomd_amp(r, p, cut, w.A, w.lnA_d);
        w.dtn[0] = w.dtn[1] = 0.;
        w.tau = 0.;
        if (need_tau) {
            double tau, dtau[2];
            const double cpm1 = 0.68278406325529568146702083315816;   // pi^(-1/3)
            tau_eval(r.tau, p.xm13 * cpm1, p.lpx3, r.lam, tau, dtau);
            w.tau = tau;
            w.dtn[0] = -dtau[0] * kInvDay;
            w.dtn[1] = -dtau[1] * kInvDay;
        }
    }
};

template <int NT> struct ModelTraits<kNRTidalv2, NT> {
    typedef NRTidalRec<NT> Rec;
    static GWF_HD void eval_amp(const Rec& r, const ModelCfg&, const FreqPoint& fp, bool need_tau, double& A, double& tau) {
        const PhenomDRec<NT>& d = r.d;
        XPow p;
        p.set(d.s, d.sp, fp);
        double T, Ty;
        nrt_taper(p.x, r.ym[0], T, Ty);
        A = 0.;
        tau = 0.;
        if (T == 0.0) return;
        double Q, xQp;
        nrt_amp_shape(1.4645918875615232630201425272637904 * p.x13, p.lpx3, Q, xQp);
        A = d.C * fma(r.sm76 * fp.fm76, phenomd_amp_value(d, p, true), r.kam[0] * Q) * T;
        if (need_tau) {
            double dtau[2];
            tau_eval(d.tau, p.xm13 * 0.68278406325529568146702083315816, p.lpx3, d.lam, tau, dtau);
        }
    }
    static GWF_HD void prologue(Rec& r, const EventIn& e, const ModelCfg& cfg, int opt_flags, const QnmTables& q, const double* fmin_g, int ng, int parts = 3) {
        // NT = 6: Fisher parametrisation (Lambda re-mapped through LambdaTilde/deltaLambda); NT = 4: SNR path, dict values as they are
        const Intrinsic<NT> p = seed_intrinsic<NT>(e, opt_flags, NT >= 6);
        nrtidal_prologue(r, p, e.dL, q, fmin_g, ng, cfg, NT < 6 || (cfg.flags & kFlagLambdaGiven) != 0, e.s_host, e.fcut_host, parts);
    }
    static GWF_HD void eval(const Rec& r, const ModelCfg& cfg, int g, const FreqPoint& fp, bool need_tau, PointWf<NT>& w) {
        const PhenomDRec<NT>& d = r.d;
        XPow p;
        p.set(d.s, d.sp, fp);
        w.dtn[0] = w.dtn[1] = 0.;
        w.tau = 0.;
        w.phi = 0.;
        const double cp = 1.4645918875615232630201425272637904;      // pi^(1/3)
        const double p13 = cp * p.x13;
        // amplitude: C [x^(-7/6) ampIMR + kam Q] T, waveforms.py:1724
        double T, Ty;
        nrt_taper(p.x, r.ym[0], T, Ty);
        double v, dv[NT];
        const bool inside = phenomd_amp_core(d, p, true, v, dv);
        if (T == 0.0) {
            w.A = 0.;
            return;
        }
        double Q, xQp;
        nrt_amp_shape(p13, p.lpx3, Q, xQp);
        const double xm76 = r.sm76 * fp.fm76;
        const double B = fma(xm76, v, r.kam[0] * Q);
        w.A = d.C * B * T;
        if (w.A == 0.0) return;
        const double iB = rcp_fast(B), tT = Ty * rcp_fast(T);
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const double dB = xm76 * (dv[j] - (7. / 6.) * d.lam[j] * v) + r.kam[1 + j] * Q + r.kam[0] * xQp * d.lam[j];
            w.lnA_d[j] = d.lnC_d[j] + dB * iB + tT * r.ym[1 + j];
        }
        (void)inside;
        // phase: PhenomD regions (+ SS/SSS folded into the x^(2/3) slots) + Pade tidal term, all inside the cut (waveforms.py:1570)
        const bool cut = !(cfg.flags & kFlagNoFcut);
        phenomd_phase(d, g, p, cut, w.phi, w.phi_d);
        if (!cut || p.x < kMfCut) {
            double R, xRp;
            nrt_phase_shape(p13, R, xRp);
            w.phi += r.kph[0] * R;
#pragma unroll
            for (int j = 0; j < NT; ++j) w.phi_d[j] += r.kph[1 + j] * R + r.kph[0] * xRp * d.lam[j];
        }
        if (need_tau) {
            double tau, dtau[2];
            const double cpm1 = 0.68278406325529568146702083315816;   // pi^(-1/3)
            tau_eval(d.tau, p.xm13 * cpm1, p.lpx3, d.lam, tau, dtau);
            w.tau = tau;
            w.dtn[0] = -dtau[0] * kInvDay;
            w.dtn[1] = -dtau[1] * kInvDay;
        }
    }
};

template <int NT> struct ModelTraits<kNSBH, NT> {
    typedef NSBHRec<NT> Rec;
    static GWF_HD void eval_amp(const Rec& r, const ModelCfg& cfg, const FreqPoint& fp, bool need_tau, double& A, double& tau) {
        const PhenomDRec<NT, false>& d = r.d;
        XPow p;
        p.set(d.s, d.sp, fp);
        A = 0.;
        tau = 0.;
        if ((cfg.flags & kFlagNoFcut) || p.x < kMfCut) {
            NsbhPowers np;
            nsbh_powers(np, p, r.sm76);
            double c[kNsbhCoef];
#pragma unroll
            for (int k = 0; k < kNsbhCoef; ++k) c[k] = r.amp[k][0];
            A = r.C * nsbh_amp_shape<double>(c, np, nullptr);
        }
        if (need_tau) {
            double dtau[2];
            tau_eval(d.tau, p.xm13 * 0.68278406325529568146702083315816, p.lpx3, d.lam, tau, dtau);
        }
    }
    static GWF_HD void prologue(Rec& r, const EventIn& e, const ModelCfg& cfg, int opt_flags, const QnmTables& q, const double* fmin_g, int ng, int = 3) {
        // NT = 6: Fisher parametrisation (Lambda re-mapped through LambdaTilde/deltaLambda, signal.py:871, 554); NT = 4: dict values as they are
        const Intrinsic<NT> p = seed_intrinsic<NT>(e, opt_flags, NT >= 6);
        nsbh_prologue(r, p, e.dL, q, fmin_g, e.fmax_g, e.fmax_exact, ng, cfg, e.s_host, e.fcut_host);
    }
    static GWF_HD void eval(const Rec& r, const ModelCfg& cfg, int g, const FreqPoint& fp, bool need_tau, PointWf<NT>& w) {
        const PhenomDRec<NT, false>& d = r.d;
        XPow p;
        p.set(d.s, d.sp, fp);
        w.dtn[0] = w.dtn[1] = 0.;
        w.tau = 0.;
        const bool cut = !(cfg.flags & kFlagNoFcut);
        phenomd_phase(d, g, p, cut, w.phi, w.phi_d);
        w.A = 0.;
#pragma unroll
        for (int j = 0; j < NT; ++j) w.lnA_d[j] = 0.;
        if (!cut || p.x < kMfCut) {
            // - t0_g x and the Pade tidal term (waveforms.py:3008-3033)
            double R, xRp;
            nrt_phase_shape(1.4645918875615232630201425272637904 * p.x13, R, xRp);
            const double t0 = r.t0[g][0];
            w.phi += r.kph[0] * R - t0 * p.x;
#pragma unroll
            for (int j = 0; j < NT; ++j)
                w.phi_d[j] += r.kph[1 + j] * R + (r.kph[0] * xRp - t0 * p.x) * d.lam[j] - r.t0[g][1 + j] * p.x;
            NsbhPowers np;
            nsbh_powers(np, p, r.sm76);
            double F, dF[NT];
            nsbh_amp_grad<NT>(r.amp, np, d.lam, F, dF);
            w.A = r.C * F;
            const double ia = 1.0 / F;
#pragma unroll
            for (int j = 0; j < NT; ++j) w.lnA_d[j] = fma(dF[j], ia, r.lnC_d[j]);
        }
        if (need_tau) {
            double tau, dtau[2];
            tau_eval(d.tau, p.xm13 * 0.68278406325529568146702083315816, p.lpx3, d.lam, tau, dtau);
            w.tau = tau;
            w.dtn[0] = -dtau[0] * kInvDay;
            w.dtn[1] = -dtau[1] * kInvDay;
        }
    }
};

// ------------------------------------------------------------------ frequency grid of one (event, group)
// numpy.geomspace(fmin, fcut, res) (signal.py:721, 898): f_k = 10**(log10 fmin + k (log10 fcut - log10 fmin)/(res-1)) with
// both end points overwritten exactly, or numpy.linspace for spacing='lin'.  A caller walks the grid with a fixed stride
// (32 in the kernels): on the geometric grid the sample and its powers are advanced by multiplying with the constant
// ratio of the stride, so no transcendental is evaluated per sample (the drift over res/stride <= 32 steps is < 1e-14).
struct Grid {
    double fmin, fcut;
    double l0, step;              // geom: log10 f_k = l0 + k*step ; lin: f_k = fmin + k*step
    double hw_in, hw_lo, hw_hi;   // trapezoid half-widths: geom -> multiply f_k; lin -> absolute
    double r, r13, rm13, rm76, dln;   // per-stride ratios of f, f^(1/3), f^(-1/3), f^(-7/6) and increment of ln f
    int res, lin, stride;
    GWF_HD void set(double fmin_, double fcut_, int res_, bool lin_, int stride_) {
        fmin = fmin_; fcut = fcut_; res = res_; lin = lin_; stride = stride_;
        if (lin) {
            step = (fcut - fmin) / (res - 1);
            hw_in = step; hw_lo = hw_hi = 0.5 * step;
            l0 = 0.;
            r = r13 = rm13 = rm76 = 1.; dln = 0.;
        } else {
            l0 = log10(fmin);
            step = (log10(fcut) - l0) / (res - 1);
            const double q = exp10(step);
            hw_in = 0.5 * (q - 1.0 / q); hw_lo = 0.5 * (q - 1.0); hw_hi = 0.5 * (1.0 - 1.0 / q);
            const double e = step * stride;                  // log10 of the stride ratio
            r = exp10(e); r13 = exp10(e * (1. / 3.)); rm13 = 1.0 / r13; rm76 = exp10(e * (-7. / 6.));
            dln = e * 2.3025850929940456840179914546843642;  // ln 10
        }
    }
    GWF_HD double freq(int k) const {
        if (k == 0) return fmin;
        if (k == res - 1) return fcut;
        return lin ? fma((double)k, step, fmin) : exp10(fma((double)k, step, l0));
    }
    GWF_HD void weight(int k, FreqPoint& p) const {
        const double hw = k == 0 ? hw_lo : (k == res - 1 ? hw_hi : hw_in);
        p.w = lin ? hw : p.f * hw;      // np.trapz on this grid, signal.py:929
    }
    // first sample of a walk
    GWF_HD void start(int k, FreqPoint& p) const {
        p.from_f(freq(k));
        weight(k, p);
    }
    // move from sample k - stride to sample k
    GWF_HD void advance(int k, FreqPoint& p) const {
        if (lin) {
            p.from_f(freq(k));
        } else {
            p.f = k == res - 1 ? fcut : p.f * r;
            p.f13 *= r13; p.fm13 *= rm13; p.fm76 *= rm76; p.lnf += dln;
        }
        weight(k, p);
    }
};

// Per-event constants of the Fisher kernel that do not depend on the detector: sky/orientation (six sincos) and the grid of
// every group (log10/exp10 calls).  The prologue kernel evaluates them once per event and leaves them next to the coefficient
// records; the Fisher kernel's warps only copy them (they used to be recomputed by all 32 lanes of every warp of an event).
struct EventAux {
    EvGeom geom;
    Grid grid[kMaxGroups];
};
static_assert(sizeof(EventAux) % sizeof(double) == 0, "EventAux is copied as doubles");

// ------------------------------------------------------------------ per-event detector scratch
// EvDet for every detector, plus the (frequency independent) DetPoint of the detectors that do not follow the
// Earth rotation.  Lives in shared memory, one block per warp (kernels) or on the stack (emulation).
struct EventScratch {
    EvDet ed[kMaxDet];
    DetPoint fixed[kMaxDet];
};
// detector `di` of the network (called by lane di in the kernels)
GWF_HD void scratch_set(EventScratch& s, const NetworkDev& net, const EvGeom& geom, int di) {
    const DetDev& d = net.det[di];
    s.ed[di].set(d, geom);
    if (d.no_motion) det_point(s.ed[di], 1.0, 0.0, s.fixed[di]);            // t = 0, signal.py:444-446
    else if (!d.use_rot) det_point(s.ed[di], geom.cT, geom.sT, s.fixed[di]);   // t = tcoal, signal.py:452
}

// compact detector `i` of a network in fast form
GWF_HD void scratch_set_fast(EventScratch& s, const NetworkDev& net, const EvGeom& geom, int i) {
    const DetDev& d = net.fdet[i];
    s.ed[i].set(d, geom);
    if (d.no_motion) det_point(s.ed[i], 1.0, 0.0, s.fixed[i]);
    else if (!d.use_rot) det_point(s.ed[i], geom.cT, geom.sT, s.fixed[i]);
}

// ------------------------------------------------------------------ one frequency point of one grid group
// acc: packed lower-triangular Fisher (NP(NP+1)/2), snr2: sum of 4 w |h|^2 / Sn over the arms of the pass
template <int MODEL, int NT>
GWF_HD void amp_phase_point(const typename ModelTraits<MODEL, NT>::Rec& rec, const ModelCfg& cfg, const EvGeom& geom, const NetworkDev& net,
                         const EventScratch& sc, int g, bool group_rot, const FreqPoint& fp, double* __restrict__ acc) {
    const double f = fp.f, l2f = fp.lnf * 1.4426950408889634073599246810018921;   // log2 f
    // PSD rows of the first detectors are requested before the waveform is evaluated: the dependent loads
    // (bucket -> row) then overlap with the waveform arithmetic instead of stalling the detector loop
    constexpr int kPre = 4;
    double sn_pre[kPre];
#pragma unroll
    for (int di = 0; di < kPre; ++di)
        sn_pre[di] = (di < net.ndet && net.det[di].group == g && net.det[di].arm_begin != net.det[di].arm_end)
                         ? psd_lookup(net.psd[net.det[di].psd], f, l2f) : 1.0;
    PointWf<NT> w;
    ModelTraits<MODEL, NT>::eval(rec, cfg, g, fp, group_rot, w);
    w.f = f;
    if (w.A == 0.0) return;               // beyond the model cut / taper: no contribution
    const double wA2 = 4.0 * fp.w * w.A * w.A;
    // Earth-rotation phase common to the detectors of the group: 2 pi (tcoal - tau/86400), signal.py:449
    double sBr = 0., cBr = 1.;
    if (group_rot) sincos(2.0 * kPi * fma(-w.tau, kInvDay, geom.tcoal), &sBr, &cBr);
    for (int di = 0; di < net.ndet; ++di) {
        const DetDev& d = net.det[di];
        if (d.group != g || d.arm_begin == d.arm_end) continue;
        const double Sn = di == 0 ? sn_pre[0] : (di == 1 ? sn_pre[1] : (di == 2 ? sn_pre[2] : (di == 3 ? sn_pre[3] : psd_lookup(net.psd[d.psd], f, l2f))));
        DetPoint dp;
        if (d.use_rot) det_point(sc.ed[di], cBr, sBr, dp);
        else dp = sc.fixed[di];
        DetRows<NT> dr;
        dr.set(w, dp, d.use_rot != 0, d.no_motion != 0);
        const double wgt = wA2 / Sn;
        for (int ai = d.arm_begin; ai < d.arm_end; ++ai) arm_rows_accumulate<NT>(w, dp, dr, net.arm[ai], geom, wgt, acc);
    }
}

// Shape of a fast-form network, known at compile time: base-4 digits = arms of the compact detectors 0..3 (0 = no detector).  The
// arms of the active detectors are consecutive in net.arm[] (build_network), so detector i's first arm is a prefix sum.  With the
// shape fixed every loop bound of a sample is a constant: the whole sample becomes one basic block and the scheduler overlaps
// the next detector's basis with the current detector's Gram (measured per 1e4 events, ET+2CE: IMRPhenomD 1.161 -> 1.056 ms,
// NRTidalv2 3.011 -> 2.327 ms).  SHAPE = 0: bounds read from the network at run time.
GWF_HD constexpr int shape_arms(int shape, int i) { return (shape >> (2 * i)) & 3; }
GWF_HD constexpr int shape_ndet(int shape) { return (shape_arms(shape, 0) != 0) + (shape_arms(shape, 1) != 0) + (shape_arms(shape, 2) != 0) + (shape_arms(shape, 3) != 0); }
GWF_HD constexpr int shape_total(int shape) { return shape_arms(shape, 0) + shape_arms(shape, 1) + shape_arms(shape, 2) + shape_arms(shape, 3); }
GWF_HD constexpr int shape_first(int shape, int i) { return (i > 0 ? shape_arms(shape, 0) : 0) + (i > 1 ? shape_arms(shape, 1) : 0) + (i > 2 ? shape_arms(shape, 2) : 0); }

// compile-time loop: f(std::integral_constant<int, I>) for I = BEGIN .. END - 1
template <int I, int END, class F>
__host__ __device__ __forceinline__ void static_for(F&& f) {
    if constexpr (I < END) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, END>(f);
    }
}

#ifdef __CUDA_ARCH__
// The same point for a network in "fast" form (NetworkDev::fast): fully unrolled detector loop over net.fdet[i] / net.fpsd[i]
// with compile-time i, ROT = every active detector follows the Earth rotation (else none does).  sc.ed / sc.fixed are
// indexed by the compact detector index.
template <int MODEL, int NT, bool ROT, int SHAPE = 0>
__device__ __forceinline__ void amp_phase_point_fast(const typename ModelTraits<MODEL, NT>::Rec& rec, const ModelCfg& cfg, const EvGeom& geom,
                                                     const NetworkDev& net, const EventScratch& sc, const FreqPoint& fp, double* __restrict__ acc) {
    const double f = fp.f, l2f = fp.lnf * 1.4426950408889634073599246810018921;   // log2 f
    // the PSD row of detector i + 1 is requested before detector i is processed (and the first one before the waveform):
    // the dependent shared-memory loads overlap with arithmetic, and only two PSD values are ever live
    double sn_next = psd_lookup_fast(net.fpsd[0], f, l2f);
    PointWf<NT> w;
    ModelTraits<MODEL, NT>::eval(rec, cfg, 0, fp, ROT, w);
    w.f = f;
    if (w.A == 0.0) return;
    const double wA2 = 4.0 * fp.w * w.A * w.A;
    double sBr = 0., cBr = 1.;
    // (sincospi where it schedules better: measured per 1e4 events, TaylorF2 0.546 -> 0.536 ms, IMRPhenomD 1.058 -> 1.112 ms)
    if (ROT) {
        if (MODEL == kTaylorF2) rot_sincos(fma(-w.tau, kInvDay, geom.tcoal), &sBr, &cBr);
        else sincos(2.0 * kPi * fma(-w.tau, kInvDay, geom.tcoal), &sBr, &cBr);
    }
    if constexpr (SHAPE != 0) {
        constexpr int kN = shape_ndet(SHAPE);
        static_for<0, kN>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            const DetDev& d = net.fdet[i];
            const double sn_i = sn_next;
            if (i + 1 < kN) sn_next = psd_lookup_fast(net.fpsd[i + 1 < kN ? i + 1 : 0], f, l2f);
            DetPoint dp;
            if (ROT) det_point(sc.ed[i], cBr, sBr, dp);
            else dp = sc.fixed[i];
            DetRows<NT> dr;
            dr.set(w, dp, ROT, ROT ? false : d.no_motion != 0);       // a detector that follows the rotation is not "noMotion"
            const double wgt = wA2 * rcp_fast(sn_i);
            constexpr int na = shape_arms(SHAPE, i), first = shape_first(SHAPE, i);
            if constexpr (na == 2) {
                // a summed triangle: the unit pair (1, 0), (0, 1) (host_build.h:build_network)
                arm_rows_accumulate<NT, 1>(w, dp, dr, net.arm[first], geom, wgt, acc);
                arm_rows_accumulate<NT, 2>(w, dp, dr, net.arm[first + 1], geom, wgt, acc);
            } else {
#pragma unroll
                for (int a = 0; a < na; ++a) arm_rows_accumulate<NT>(w, dp, dr, net.arm[first + a], geom, wgt, acc);
            }
        });
    } else {
#pragma unroll
        for (int i = 0; i < kMaxFastDet; ++i) {
            if (i < net.fnd) {
                const DetDev& d = net.fdet[i];
                const double sn_i = sn_next;
                if (i + 1 < kMaxFastDet && i + 1 < net.fnd) sn_next = psd_lookup_fast(net.fpsd[i + 1 < kMaxFastDet ? i + 1 : 0], f, l2f);
                DetPoint dp;
                if (ROT) det_point(sc.ed[i], cBr, sBr, dp);
                else dp = sc.fixed[i];
                DetRows<NT> dr;
                dr.set(w, dp, ROT, d.no_motion != 0);
                const double wgt = wA2 * rcp_fast(sn_i);
                for (int ai = d.arm_begin; ai < d.arm_end; ++ai) arm_rows_accumulate<NT>(w, dp, dr, net.arm[ai], geom, wgt, acc);
            }
        }
    }
}
#endif

// value-only point for the SNR kernel: per-arm SNR^2 contributions (signal.py:725-767)
// Per-arm SNR^2 accumulators: on the device they live in shared memory as [arm][lane] (a dynamically indexed register
// array would be demoted to local memory); the host emulation keeps a plain array.
#ifdef __CUDA_ARCH__
#define GWF_SNR_SLOT(a) ((a) * 32)
#else
#define GWF_SNR_SLOT(a) (a)
#endif

template <int MODEL>
GWF_HD void amp_phase_snr_point(const typename ModelTraits<MODEL, 4>::Rec& rec, const ModelCfg& cfg, const EvGeom& geom, const NetworkDev& net,
                      const EventScratch& sc, int g, bool group_rot, const FreqPoint& fp, double* __restrict__ snr2_arm) {
    const double f = fp.f, l2f = fp.lnf * 1.4426950408889634073599246810018921;   // log2 f
    double A, tau;
    ModelTraits<MODEL, 4>::eval_amp(rec, cfg, fp, group_rot, A, tau);
    (void)g;
    if (A == 0.0) return;
    const double wA2 = 4.0 * fp.w * A * A;
    double sBr = 0., cBr = 1.;
    if (group_rot) rot_sincos(fma(-tau, kInvDay, geom.tcoal), &sBr, &cBr);
    for (int di = 0; di < net.ndet; ++di) {
        const DetDev& d = net.det[di];
        if (d.group != g) continue;
        const double Sn = psd_lookup(net.psd[d.psd], f, l2f);
        DetPoint dp;
        if (d.use_rot) det_point(sc.ed[di], cBr, sBr, dp);
        else dp = sc.fixed[di];
        const double wgt = wA2 / Sn;
        for (int ai = d.arm_begin; ai < d.arm_end; ++ai) {
            double Fp, Fc;
            arm_pattern(dp, net.arm[ai], geom, Fp, Fc);
            const double Gr = Fp * geom.K, Gi = Fc * geom.ci;
            double& slot = snr2_arm[GWF_SNR_SLOT(net.arm[ai].out)];
            slot = fma(wgt * net.arm[ai].weight, Gr * Gr + Gi * Gi, slot);
        }
    }
}

#ifdef __CUDA_ARCH__
// value-only point for a network in fast form (see amp_phase_point_fast)
template <int MODEL, bool ROT, int SHAPE = 0>
__device__ __forceinline__ void amp_phase_snr_point_fast(const typename ModelTraits<MODEL, 4>::Rec& rec, const ModelCfg& cfg, const EvGeom& geom,
                                                         const NetworkDev& net, const EventScratch& sc, const FreqPoint& fp, double* __restrict__ snr2_arm) {
    const double f = fp.f, l2f = fp.lnf * 1.4426950408889634073599246810018921;   // log2 f
    double sn[kMaxFastDet];
#pragma unroll
    for (int i = 0; i < kMaxFastDet; ++i) sn[i] = (SHAPE != 0 ? i < shape_ndet(SHAPE) : i < net.fnd) ? psd_lookup_fast(net.fpsd[i], f, l2f) : 1.0;
    double A, tau;
    ModelTraits<MODEL, 4>::eval_amp(rec, cfg, fp, ROT, A, tau);
    if (A == 0.0) return;
    const double wA2 = 4.0 * fp.w * A * A;
    double sBr = 0., cBr = 1.;
    if (ROT) rot_sincos(fma(-tau, kInvDay, geom.tcoal), &sBr, &cBr);
    if constexpr (SHAPE != 0) {
        constexpr int kN = shape_ndet(SHAPE);
#pragma unroll
        for (int i = 0; i < kN; ++i) {
            DetPoint dp;
            if (ROT) det_point(sc.ed[i], cBr, sBr, dp);
            else dp = sc.fixed[i];
            const double wgt = wA2 * rcp_fast(sn[i]);
#pragma unroll
            for (int a = 0; a < shape_arms(SHAPE, i); ++a) {
                // shape-specialised form: snr2_arm is a per-lane REGISTER array indexed by compile-time constants -- in SNR mode
                // the arms of a network are its physical arms in order, each with weight 1 and output slot = its index
                // (host_build.h:build_network), so neither arm.out nor arm.weight is read
                const ArmDev& arm = net.arm[shape_first(SHAPE, i) + a];
                double Fp, Fc;
                arm_pattern(dp, arm, geom, Fp, Fc);
                const double Gr = Fp * geom.K, Gi = Fc * geom.ci;
                double& slot = snr2_arm[shape_first(SHAPE, i) + a];
                slot = fma(wgt, fma(Gr, Gr, Gi * Gi), slot);
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < kMaxFastDet; ++i) {
            if (i < net.fnd) {
                const DetDev& d = net.fdet[i];
                DetPoint dp;
                if (ROT) det_point(sc.ed[i], cBr, sBr, dp);
                else dp = sc.fixed[i];
                const double wgt = wA2 * rcp_fast(sn[i]);
                for (int ai = d.arm_begin; ai < d.arm_end; ++ai) {
                    double Fp, Fc;
                    arm_pattern(dp, net.arm[ai], geom, Fp, Fc);
                    const double Gr = Fp * geom.K, Gi = Fc * geom.ci;
                    double& slot = snr2_arm[GWF_SNR_SLOT(net.arm[ai].out)];
                    slot = fma(wgt * net.arm[ai].weight, Gr * Gr + Gi * Gi, slot);
                }
            }
        }
    }
}
#endif

// ------------------------------------------------------------------ IMRPhenomHM point functions
// per-event extras of a model (harmonic weights for HM; nothing for the (2,2)-only models)
struct NoExtra {
    GWF_HD void set(const EventIn&) {}
};
struct HMExtra {
    HMWeights w;
    GWF_HD void set(const EventIn& e) { w.set(e.iota); }
};

// mode sums of hphc: z_m = A_m exp(-i Phi_m) is folded into h+ / hx (with their intrinsic tangents and iota derivatives) as soon
// as it is known: d z_m = z_m (d ln A_m - i d Phi_m)
template <int NT> struct HMStrainSink {
    const HMWeights& w;
    double hpr, hpi, hcr, hci;
    double hpr_d[NT], hpi_d[NT], hcr_d[NT], hci_d[NT];
    double hpr_i, hpi_i, hcr_i, hci_i;
    GWF_HD explicit HMStrainSink(const HMWeights& w_) : w(w_), hpr(0.), hpi(0.), hcr(0.), hci(0.), hpr_i(0.), hpi_i(0.), hcr_i(0.), hci_i(0.) {
#pragma unroll
        for (int j = 0; j < NT; ++j) hpr_d[j] = hpi_d[j] = hcr_d[j] = hci_d[j] = 0.;
    }
    // partial sums over a subset of the modes, exchanged between the two warps of a pair: kWords doubles per lane, laid out [word][lane]
    static constexpr int kWords = 8 + 4 * NT;
    GWF_HD void store(double* __restrict__ x, int lane) const {
        x[0 * 32 + lane] = hpr; x[1 * 32 + lane] = hpi; x[2 * 32 + lane] = hcr; x[3 * 32 + lane] = hci;
        x[4 * 32 + lane] = hpr_i; x[5 * 32 + lane] = hpi_i; x[6 * 32 + lane] = hcr_i; x[7 * 32 + lane] = hci_i;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            x[(8 + 4 * j) * 32 + lane] = hpr_d[j]; x[(9 + 4 * j) * 32 + lane] = hpi_d[j];
            x[(10 + 4 * j) * 32 + lane] = hcr_d[j]; x[(11 + 4 * j) * 32 + lane] = hci_d[j];
        }
    }
    GWF_HD void add(const double* __restrict__ x, int lane) {
        hpr += x[0 * 32 + lane]; hpi += x[1 * 32 + lane]; hcr += x[2 * 32 + lane]; hci += x[3 * 32 + lane];
        hpr_i += x[4 * 32 + lane]; hpi_i += x[5 * 32 + lane]; hcr_i += x[6 * 32 + lane]; hci_i += x[7 * 32 + lane];
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            hpr_d[j] += x[(8 + 4 * j) * 32 + lane]; hpi_d[j] += x[(9 + 4 * j) * 32 + lane];
            hcr_d[j] += x[(10 + 4 * j) * 32 + lane]; hci_d[j] += x[(11 + 4 * j) * 32 + lane];
        }
    }
    GWF_HD void operator()(int m, double A, double ph, const double* __restrict__ dlnA, const double* __restrict__ dPhi) {
        if (A == 0.0) return;
        double sn, cs;
        sincos(ph, &sn, &cs);
        const double zre = A * cs, zim = -(A * sn);
        const double wp = w.wp[m], wc = w.wc[m];
        hpr = fma(zre, wp, hpr);
        hpi = fma(zim, wp, hpi);
        hcr = fma(-zim, wc, hcr);
        hci = fma(zre, wc, hci);
        hpr_i = fma(zre, w.dwp[m], hpr_i);
        hpi_i = fma(zim, w.dwp[m], hpi_i);
        hcr_i = fma(-zim, w.dwc[m], hcr_i);
        hci_i = fma(zre, w.dwc[m], hci_i);
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const double dre = fma(zre, dlnA[j], zim * dPhi[j]), dim = fma(zim, dlnA[j], -zre * dPhi[j]);
            hpr_d[j] = fma(dre, wp, hpr_d[j]);
            hpi_d[j] = fma(dim, wp, hpi_d[j]);
            hcr_d[j] = fma(-dim, wc, hcr_d[j]);
            hci_d[j] = fma(dre, wc, hci_d[j]);
        }
    }
};
struct HMValueSink {
    const HMWeights& w;
    double hpr, hpi, hcr, hci;
    GWF_HD explicit HMValueSink(const HMWeights& w_) : w(w_), hpr(0.), hpi(0.), hcr(0.), hci(0.) {}
    GWF_HD void operator()(int m, double A, double ph, const double*, const double*) {
        if (A == 0.0) return;
        double sn, cs;
        sincos(ph, &sn, &cs);
        const double zre = A * cs, zim = -(A * sn);
        hpr = fma(zre, w.wp[m], hpr);
        hpi = fma(zim, w.wp[m], hpi);
        hcr = fma(-zim, w.wc[m], hcr);
        hci = fma(zre, w.wc[m], hci);
    }
};

// accumulator layout of the IMRPhenomHM Fisher point: packed Gram, then 4 int |h|^2/Sn, then the SNRInteg integral (cross term
// dropped), then (SD) the (h | d_i h) rows
template <int NT> struct HMAcc {
    static constexpr int NP = NT + 7, kSnr2 = NP * (NP + 1) / 2, kSnr2Integ = kSnr2 + 1, kSd = kSnr2 + 2;
};

// rows of d h / d p_i (ParNums order) of one arm at one sample, without the carrier e^{i (2 pi f tcoal 86400 - Phicoal + 2 pi f Delta t)}
// (it cancels in the Gram): h = hp Fp + hc Fc (signal.py:584-607); H = (Hr, Hi) is the strain itself in the same convention
template <int NT>
GWF_HD void hm_arm_rows(const HMStrainSink<NT>& hs, const DetPoint& dp, const DetRows<NT>& dr, const ArmDev& a, const EvGeom& geom,
                        double* __restrict__ ra, double* __restrict__ rb, double& Hr, double& Hi, double& Fp, double& Fc) {
    const double av = a.S2 * dp.aS + a.C2 * dp.aC, bv = a.C2 * dp.bC + a.S2 * dp.bS;
    const double ag = a.S2 * dp.aS_g + a.C2 * dp.aC_g, bg = a.C2 * dp.bC_g + a.S2 * dp.bS_g;
    const double ad = a.S2 * dp.aS_d + a.C2 * dp.aC_d, bd = a.C2 * dp.bC_d + a.S2 * dp.bS_d;
    Fp = av * geom.c2psi + bv * geom.s2psi; Fc = bv * geom.c2psi - av * geom.s2psi;
    const double Fpg = ag * geom.c2psi + bg * geom.s2psi, Fcg = bg * geom.c2psi - ag * geom.s2psi;
    const double Fpd = ad * geom.c2psi + bd * geom.s2psi, Fcd = bd * geom.c2psi - ad * geom.s2psi;
    Hr = hs.hpr * Fp + hs.hcr * Fc; Hi = hs.hpi * Fp + hs.hci * Fc;           // signal.py:586-607
    const double Hgr = hs.hpr * Fpg + hs.hcr * Fcg, Hgi = hs.hpi * Fpg + hs.hci * Fcg;
    const double Hdr = hs.hpr * Fpd + hs.hcr * Fcd, Hdi = hs.hpi * Fpd + hs.hci * Fcd;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        const int row = j < 2 ? j : 7 + j;
        double re = hs.hpr_d[j] * Fp + hs.hcr_d[j] * Fc - Hi * dr.psi_x[j], im = hs.hpi_d[j] * Fp + hs.hci_d[j] * Fc + Hr * dr.psi_x[j];
        if (j < 2) {
            re = fma(Hgr, dr.ang_x[j], re);
            im = fma(Hgi, dr.ang_x[j], im);
        }
        ra[row] = re;
        rb[row] = im;
    }
    ra[2] = -Hr * geom.inv_dL;                                   rb[2] = -Hi * geom.inv_dL;
    ra[3] = fma(Hgr, dr.ang_t, -Hdr) - Hi * dr.ph_t;             rb[3] = fma(Hgi, dr.ang_t, -Hdi) + Hr * dr.ph_t;
    ra[4] = fma(Hgr, dr.ang_p, -Hi * dr.ph_p);                   rb[4] = fma(Hgi, dr.ang_p, Hr * dr.ph_p);
    ra[5] = hs.hpr_i * Fp + hs.hcr_i * Fc;                       rb[5] = hs.hpi_i * Fp + hs.hci_i * Fc;     // iota: harmonics only
    ra[6] = 2.0 * (hs.hpr * Fc - hs.hcr * Fp);                   rb[6] = 2.0 * (hs.hpi * Fc - hs.hci * Fp);   // psi: Fp' = 2 Fc, Fc' = -2 Fp
    ra[7] = fma(Hgr, dr.ang_c, -Hi * dr.ph_c);                   rb[7] = fma(Hgi, dr.ang_c, Hr * dr.ph_c);
    ra[8] = Hi;                                                  rb[8] = -Hr;
}

// SD: also accumulate (h | d_i h) (return_SNR_derivatives)
template <int NT, bool SD = false>
GWF_HD void hm_point(const HMRec<NT>& rec, const ModelCfg& cfg, const EvGeom& geom, const NetworkDev& net, const EventScratch& sc,
                     const HMExtra& ex, int g, bool group_rot, const FreqPoint& fp, double* __restrict__ acc) {
    constexpr int NP = NT + 7;
    double& snr2 = acc[HMAcc<NT>::kSnr2];
    double& snr2i = acc[HMAcc<NT>::kSnr2Integ];
    const double f = fp.f, l2f = fp.lnf * 1.4426950408889634073599246810018921;
    const bool cut = !(cfg.flags & kFlagNoFcut);
    // hp = sum z_m Wp_m ; hc = i sum z_m Wc_m ; and their iota derivatives (waveforms.py:2613-2614), summed mode by mode
    HMStrainSink<NT> hs(ex.w);
    phenomhm_foreach_mode<NT, true>(rec, g, fp, cut, hs);
    if (hs.hpr == 0.0 && hs.hpi == 0.0 && hs.hcr == 0.0 && hs.hci == 0.0) return;
    const double hp2 = hs.hpr * hs.hpr + hs.hpi * hs.hpi, hc2 = hs.hcr * hs.hcr + hs.hci * hs.hci;
    PointWf<NT> w;
    w.f = f;
#pragma unroll
    for (int j = 0; j < NT; ++j) w.phi_d[j] = 0.;
    w.dtn[0] = w.dtn[1] = 0.;
    double sBr = 0., cBr = 1.;
    if (group_rot) {
        double tau, dtau[2];
        const double xm13 = rec.sp.sm13 * fp.fm13, lpx3 = fma(fp.lnf, 1. / 3., rec.sp.lps3);
        tau_eval(rec.tau, 0.68278406325529568146702083315816 * xm13, lpx3, rec.lam, tau, dtau);
        w.dtn[0] = -dtau[0] * kInvDay;
        w.dtn[1] = -dtau[1] * kInvDay;
        sincos(2.0 * kPi * fma(-tau, kInvDay, geom.tcoal), &sBr, &cBr);
    }
    const double w4 = 4.0 * fp.w;
    for (int di = 0; di < net.ndet; ++di) {
        const DetDev& d = net.det[di];
        if (d.group != g || d.arm_begin == d.arm_end) continue;
        const double Sn = psd_lookup(net.psd[d.psd], f, l2f);
        DetPoint dp;
        if (d.use_rot) det_point(sc.ed[di], cBr, sBr, dp);
        else dp = sc.fixed[di];
        DetRows<NT> dr;
        dr.set(w, dp, d.use_rot != 0, d.no_motion != 0);
        const double wgt = w4 / Sn;
        for (int ai = d.arm_begin; ai < d.arm_end; ++ai) {
            const ArmDev& a = net.arm[ai];
            double ra[NP], rb[NP], Hr, Hi, Fp, Fc;
            hm_arm_rows<NT>(hs, dp, dr, a, geom, ra, rb, Hr, Hi, Fp, Fc);
            const double wg = wgt * a.weight;
            snr2 = fma(wg, Hr * Hr + Hi * Hi, snr2);
            snr2i = fma(wg, hp2 * Fp * Fp + hc2 * Fc * Fc, snr2i);      // Ap = |hp| Fp, Ac = |hc| Fc (signal.py:457-460, 727)
            if (SD) {
                const double wr = wg * Hr, wi = wg * Hi;
#pragma unroll
                for (int i = 0; i < NP; ++i) acc[HMAcc<NT>::kSd + i] = fma(wr, ra[i], fma(wi, rb[i], acc[HMAcc<NT>::kSd + i]));
            }
#pragma unroll
            for (int i = 0; i < NP; ++i) {
                const double wa = wg * ra[i], wb = wg * rb[i];
#pragma unroll
                for (int j = 0; j <= i; ++j) acc[tri(i, j)] = fma(wa, ra[j], fma(wb, rb[j], acc[tri(i, j)]));
            }
        }
    }
}

#ifdef __CUDACC__
// ---- IMRPhenomHM with the work of one block of 32 samples SPLIT over the two warps of a pair ------------------------------------
// The full point above keeps 68 accumulators (136 registers) alive while the six modes are evaluated and spilled ~600 local loads
// and stores per warp-sample (profiles/r02b: FP64 pipe 24 %, long_scoreboard 3.1 cycles per issue).  Here both warps of a pair work on
// the SAME samples: warp `half` evaluates modes 3 half .. 3 half + 2, the partial mode sums (8 + 4 NT doubles per lane) are exchanged
// through shared memory, and each warp then forms the rows of every arm but accumulates only the packed entries it owns (p & 1 ==
// half): 34 accumulators per lane instead of 68, no mode work done twice, only the row assembly (~10 % of the sample) is duplicated.
template <int NT, bool SD> struct HMSplitAcc {
    static constexpr int NP = NT + 7, NPACK = NP * (NP + 1) / 2;
    static_assert(NPACK % 2 == 0, "entries are dealt out alternately");
    static constexpr int kSnr = NPACK / 2;             // half 0: 4 int |h|^2/Sn ; half 1: the SNRInteg integral
    static constexpr int kSd = kSnr + 1;               // (h | d_i h) of the rows i with (i & 1) == half, at kSd + (i >> 1)
    static constexpr int kAcc = kSd + (SD ? (NP + 1) / 2 : 0);
};
template <int NT, bool SD, int HALF>
__device__ __forceinline__ void hm_gram_half(double wg, const double* __restrict__ ra, const double* __restrict__ rb, double Hr, double Hi,
                                             double* __restrict__ acc) {
    constexpr int NP = NT + 7;
    if (SD) {
        const double wr = wg * Hr, wi = wg * Hi;
#pragma unroll
        for (int i = HALF; i < NP; i += 2) acc[HMSplitAcc<NT, SD>::kSd + (i >> 1)] = fma(wr, ra[i], fma(wi, rb[i], acc[HMSplitAcc<NT, SD>::kSd + (i >> 1)]));
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        const double wa = wg * ra[i], wb = wg * rb[i];
#pragma unroll
        for (int j = 0; j <= i; ++j)
            if ((tri(i, j) & 1) == HALF) acc[tri(i, j) >> 1] = fma(wa, ra[j], fma(wb, rb[j], acc[tri(i, j) >> 1]));
    }
}
// xbuf: [2 halves][kWords][32] doubles of the pair; bar_id: the pair's named barrier (64 threads).  Every lane of both warps must call
// this for every block (inactive lanes -- beyond the grid -- only take part in the exchange).
template <int NT, bool SD>
__device__ __forceinline__ void hm_point_split(const HMRec<NT>& rec, const ModelCfg& cfg, const EvGeom& geom, const NetworkDev& net, const EventScratch& sc,
                                               const HMExtra& ex, int g, bool group_rot, const FreqPoint& fp, double* __restrict__ acc, int half,
                                               double* __restrict__ xbuf, int bar_id, int lane, bool active) {
    constexpr int NP = NT + 7;
    typedef HMSplitAcc<NT, SD> L;
    constexpr int kW = HMStrainSink<NT>::kWords;
    HMStrainSink<NT> hs(ex.w);
    PointWf<NT> w;
    double sBr = 0., cBr = 1.;
    // PSD rows of the first detectors are requested before the modes are evaluated: tables that found no room in shared memory are
    // read through L1/L2 (two dependent loads), and that latency then hides behind the mode arithmetic instead of stalling the detector loop
    constexpr int kPre = 4;
    double sn_pre[kPre];
    const double f = fp.f, l2f = fp.lnf * 1.4426950408889634073599246810018921;
#pragma unroll
    for (int di = 0; di < kPre; ++di)
        sn_pre[di] = (active && di < net.ndet && net.det[di].group == g && net.det[di].arm_begin != net.det[di].arm_end)
                         ? psd_lookup(net.psd[net.det[di].psd], f, l2f) : 1.0;
    if (active) {
        phenomhm_foreach_mode<NT, true>(rec, g, fp, !(cfg.flags & kFlagNoFcut), hs, 3 * half, 3 * half + 3);
        w.f = fp.f;
#pragma unroll
        for (int j = 0; j < NT; ++j) w.phi_d[j] = 0.;
        w.dtn[0] = w.dtn[1] = 0.;
        if (group_rot) {                      // every read of the staged record happens before the exchange barrier
            double tau, dtau[2];
            const double xm13 = rec.sp.sm13 * fp.fm13, lpx3 = fma(fp.lnf, 1. / 3., rec.sp.lps3);
            tau_eval(rec.tau, 0.68278406325529568146702083315816 * xm13, lpx3, rec.lam, tau, dtau);
            w.dtn[0] = -dtau[0] * kInvDay;
            w.dtn[1] = -dtau[1] * kInvDay;
            sincos(2.0 * kPi * fma(-tau, kInvDay, geom.tcoal), &sBr, &cBr);
        }
    }
    hs.store(xbuf + half * (kW * 32), lane);
    asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
    hs.add(xbuf + (1 - half) * (kW * 32), lane);
    asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");     // the partner has read this warp's words: the slot may be rewritten
    if (!active || (hs.hpr == 0.0 && hs.hpi == 0.0 && hs.hcr == 0.0 && hs.hci == 0.0)) return;
    const double hp2 = hs.hpr * hs.hpr + hs.hpi * hs.hpi, hc2 = hs.hcr * hs.hcr + hs.hci * hs.hci;
    const double w4 = 4.0 * fp.w;
    for (int di = 0; di < net.ndet; ++di) {
        const DetDev& d = net.det[di];
        if (d.group != g || d.arm_begin == d.arm_end) continue;
        const double Sn = di == 0 ? sn_pre[0] : (di == 1 ? sn_pre[1] : (di == 2 ? sn_pre[2] : (di == 3 ? sn_pre[3] : psd_lookup(net.psd[d.psd], f, l2f))));
        DetPoint dp;
        if (d.use_rot) det_point(sc.ed[di], cBr, sBr, dp);
        else dp = sc.fixed[di];
        DetRows<NT> dr;
        dr.set(w, dp, d.use_rot != 0, d.no_motion != 0);
        const double wgt = w4 * rcp_fast(Sn);
        for (int ai = d.arm_begin; ai < d.arm_end; ++ai) {
            const ArmDev& a = net.arm[ai];
            double ra[NP], rb[NP], Hr, Hi, Fp, Fc;
            hm_arm_rows<NT>(hs, dp, dr, a, geom, ra, rb, Hr, Hi, Fp, Fc);
            const double wg = wgt * a.weight;
            if (half == 0) {
                acc[L::kSnr] = fma(wg, Hr * Hr + Hi * Hi, acc[L::kSnr]);
                hm_gram_half<NT, SD, 0>(wg, ra, rb, Hr, Hi, acc);
            } else {
                acc[L::kSnr] = fma(wg, hp2 * Fp * Fp + hc2 * Fc * Fc, acc[L::kSnr]);      // Ap = |hp| Fp, Ac = |hc| Fc (signal.py:457-460, 727)
                hm_gram_half<NT, SD, 1>(wg, ra, rb, Hr, Hi, acc);
            }
        }
    }
}
#endif

// HM SNR as the reference defines it: Ap = |hp| Fp, Ac = |hc| Fc (the +/x cross term is dropped, signal.py:457-460, 727)
GWF_HD void hm_snr_point(const HMRec<4>& rec, const ModelCfg& cfg, const EvGeom& geom, const NetworkDev& net, const EventScratch& sc,
                         const HMExtra& ex, int g, bool group_rot, const FreqPoint& fp, double* __restrict__ snr2_arm) {
    const double f = fp.f, l2f = fp.lnf * 1.4426950408889634073599246810018921;
    const bool cut = !(cfg.flags & kFlagNoFcut);
    HMValueSink hs(ex.w);
    phenomhm_foreach_mode<4, false>(rec, g, fp, cut, hs);
    const double hpr = hs.hpr, hpi = hs.hpi, hcr = hs.hcr, hci = hs.hci;
    const double hp2 = hpr * hpr + hpi * hpi, hc2 = hcr * hcr + hci * hci;
    if (hp2 == 0.0 && hc2 == 0.0) return;
    double sBr = 0., cBr = 1.;
    if (group_rot) {
        double tau, dtau[2];
        const double xm13 = rec.sp.sm13 * fp.fm13, lpx3 = fma(fp.lnf, 1. / 3., rec.sp.lps3);
        tau_eval(rec.tau, 0.68278406325529568146702083315816 * xm13, lpx3, rec.lam, tau, dtau);
        rot_sincos(fma(-tau, kInvDay, geom.tcoal), &sBr, &cBr);
    }
    const double w4 = 4.0 * fp.w;
    for (int di = 0; di < net.ndet; ++di) {
        const DetDev& d = net.det[di];
        if (d.group != g) continue;
        const double Sn = psd_lookup(net.psd[d.psd], f, l2f);
        DetPoint dp;
        if (d.use_rot) det_point(sc.ed[di], cBr, sBr, dp);
        else dp = sc.fixed[di];
        const double wgt = w4 / Sn;
        for (int ai = d.arm_begin; ai < d.arm_end; ++ai) {
            double Fp, Fc;
            arm_pattern(dp, net.arm[ai], geom, Fp, Fc);
            double& slot = snr2_arm[GWF_SNR_SLOT(net.arm[ai].out)];
            slot = fma(wgt * net.arm[ai].weight, hp2 * Fp * Fp + hc2 * Fc * Fc, slot);
        }
    }
}

template <int NT> struct ModelTraits<kPhenomHM, NT> {
    typedef HMRec<NT> Rec;
    typedef HMExtra Extra;
    static GWF_HD void prologue(Rec& r, const EventIn& e, const ModelCfg& cfg, int opt_flags, const QnmTables&, const double* fmin_g, int ng, int = 3) {
        const Intrinsic<NT> p = seed_intrinsic<NT>(e, opt_flags, false);
        phenomhm_prologue(r, p, e.dL, fmin_g, ng, cfg, e.s_host, e.fcut_host);
    }
};

// uniform entry points used by the kernels and the emulation harness.  kAcc = number of per-lane accumulators; after the
// reduction entry(p) rebuilds packed Fisher element p (row-major lower triangle) and snr2() the SNR^2 of the pass.
template <int MODEL, int NT> struct PointFns {
    typedef NoExtra Extra;
    typedef typename ModelTraits<MODEL, NT>::Rec Rec;
    static constexpr int kAcc = Compact<NT>::kAcc;
    static GWF_HD void fisher(const Rec& rec, const ModelCfg& cfg, const EvGeom& geom, const NetworkDev& net, const EventScratch& sc, const Extra&, int g,
                              bool rot, const FreqPoint& fp, double* __restrict__ acc) {
        amp_phase_point<MODEL, NT>(rec, cfg, geom, net, sc, g, rot, fp, acc);
    }
    static constexpr bool kHasFast = true;
#ifdef __CUDA_ARCH__
    template <bool ROT, int SHAPE = 0>
    static __device__ __forceinline__ void fisher_fast(const Rec& rec, const ModelCfg& cfg, const EvGeom& geom, const NetworkDev& net, const EventScratch& sc,
                                                       const Extra&, const FreqPoint& fp, double* __restrict__ acc) {
        amp_phase_point_fast<MODEL, NT, ROT, SHAPE>(rec, cfg, geom, net, sc, fp, acc);
    }
    template <bool ROT, int SHAPE = 0>
    static __device__ __forceinline__ void snr_fast(const Rec& rec, const ModelCfg& cfg, const EvGeom& geom, const NetworkDev& net, const EventScratch& sc,
                                                    const Extra&, const FreqPoint& fp, double* __restrict__ s2) {
        amp_phase_snr_point_fast<MODEL, ROT, SHAPE>(rec, cfg, geom, net, sc, fp, s2);
    }
#endif
    static GWF_HD double entry(int i, int j, const double* __restrict__ red, const EvGeom& geom) { return compact_entry<NT>(i, j, red, geom); }
    // table form of entry() for the kernels (compact_entry_code / compact_coef)
    static constexpr bool kEntryTable = true;
    static GWF_HD void entry_code(int i, int j, unsigned char* code) { compact_entry_code<NT>(i, j, code); }
    static GWF_HD double snr2(const double* __restrict__ red, const EvGeom& geom) { return compact_snr2<NT>(red, geom); }
    static GWF_HD double snr2_integ(const double* __restrict__ red, const EvGeom& geom) { return compact_snr2<NT>(red, geom); }
    static GWF_HD double snr_deriv(int row, const double* __restrict__ red, const EvGeom& geom) { return compact_snr_deriv<NT>(row, red, geom); }
    static GWF_HD void snr(const Rec& rec, const ModelCfg& cfg, const EvGeom& geom, const NetworkDev& net, const EventScratch& sc, const Extra&, int g,
                           bool rot, const FreqPoint& fp, double* __restrict__ s2) {
        amp_phase_snr_point<MODEL>(rec, cfg, geom, net, sc, g, rot, fp, s2);
    }
};
// IMRPhenomHM: SD = true adds the (h | d_i h) accumulators (the (2,2)-only models get them from the compact Gram for free)
template <int NT, bool SD = false> struct PointFnsHM {
    typedef HMExtra Extra;
    typedef HMRec<NT> Rec;
    static constexpr int kAcc = HMAcc<NT>::kSd + (SD ? NT + 7 : 0);
    static GWF_HD void fisher(const Rec& rec, const ModelCfg& cfg, const EvGeom& geom, const NetworkDev& net, const EventScratch& sc, const Extra& ex, int g,
                              bool rot, const FreqPoint& fp, double* __restrict__ acc) {
        hm_point<NT, SD>(rec, cfg, geom, net, sc, ex, g, rot, fp, acc);
    }
    static GWF_HD double snr_deriv(int row, const double* __restrict__ red, const EvGeom&) { return SD ? red[HMAcc<NT>::kSd + row] : 0.0; }
    static constexpr bool kHasFast = false;
#ifdef __CUDA_ARCH__
    template <bool ROT, int SHAPE = 0>
    static __device__ __forceinline__ void fisher_fast(const Rec& rec, const ModelCfg& cfg, const EvGeom& geom, const NetworkDev& net, const EventScratch& sc,
                                                       const Extra& ex, const FreqPoint& fp, double* __restrict__ acc) {
        hm_point<NT, SD>(rec, cfg, geom, net, sc, ex, 0, ROT, fp, acc);
    }
    template <bool ROT, int SHAPE = 0>
    static __device__ __forceinline__ void snr_fast(const Rec& rec, const ModelCfg& cfg, const EvGeom& geom, const NetworkDev& net, const EventScratch& sc,
                                                    const Extra& ex, const FreqPoint& fp, double* __restrict__ s2) {
        hm_snr_point(rec, cfg, geom, net, sc, ex, 0, ROT, fp, s2);
    }
#endif
    static GWF_HD double entry(int i, int j, const double* __restrict__ red, const EvGeom&) { return red[tri(i, j)]; }
    static constexpr bool kEntryTable = false;      // packed index = accumulator index
    static GWF_HD void entry_code(int, int, unsigned char*) {}
    static GWF_HD double snr2(const double* __restrict__ red, const EvGeom&) { return red[HMAcc<NT>::kSnr2]; }
    static GWF_HD double snr2_integ(const double* __restrict__ red, const EvGeom&) { return red[HMAcc<NT>::kSnr2Integ]; }
    static GWF_HD void snr(const Rec& rec, const ModelCfg& cfg, const EvGeom& geom, const NetworkDev& net, const EventScratch& sc, const Extra& ex, int g,
                           bool rot, const FreqPoint& fp, double* __restrict__ s2) {
        hm_snr_point(rec, cfg, geom, net, sc, ex, g, rot, fp, s2);
    }
};

template <int NT> struct PointFns<kPhenomHM, NT> : PointFnsHM<NT, false> {};
// selects the accumulator set of a launch: SD only changes IMRPhenomHM
template <int MODEL, int NT, bool SD> struct PointFnsSel { typedef PointFns<MODEL, NT> type; };
template <int NT> struct PointFnsSel<kPhenomHM, NT, true> { typedef PointFnsHM<NT, true> type; };

// ------------------------------------------------------------------ stand-alone waveform values (WaveFormModel.Phi/Ampl/tau_star/hphc)
// out: phi[nm], amp[nm] (nm = 1, or 6 for IMRPhenomHM), tau, and for HM hp = (re, im), hc = (re, im)
struct WaveformOut {
    double phi[kHMModes], amp[kHMModes], tau, hp[2], hc[2];
};
template <int MODEL> struct WaveformFns;
template <> struct WaveformFns<kTaylorF2> {
    static constexpr int kModes = 1;
    static GWF_HD void eval(const TF2Rec<4>& r, const ModelCfg&, const HMWeights&, const FreqPoint& fp, WaveformOut& o) {
        VPow p;
        p.set(r.sp, fp);
        double pd[4], dtau[2];
        tf2_phase(r, p, o.phi[0], pd);
        o.amp[0] = r.C * fp.fm76;
        tau_eval(r.tau, p.vm1, p.lpx3, r.lam, o.tau, dtau);
    }
};
template <> struct WaveformFns<kPhenomD> {
    static constexpr int kModes = 1;
    static GWF_HD void eval(const PhenomDRec<4>& r, const ModelCfg& cfg, const HMWeights&, const FreqPoint& fp, WaveformOut& o) {
        XPow p;
        p.set(r.s, r.sp, fp);
        const bool cut = !(cfg.flags & kFlagNoFcut);
        double pd[4], ad[4], dtau[2];
        phenomd_phase(r, 0, p, cut, o.phi[0], pd);
        phenomd_amp(r, p, cut, o.amp[0], ad);
        tau_eval(r.tau, p.xm13 * 0.68278406325529568146702083315816, p.lpx3, r.lam, o.tau, dtau);
    }
};
template <> struct WaveformFns<kNRTidalv2> {
    static constexpr int kModes = 1;
    static GWF_HD void eval(const NRTidalRec<4>& r, const ModelCfg& cfg, const HMWeights&, const FreqPoint& fp, WaveformOut& o) {
        const PhenomDRec<4>& d = r.d;
        XPow p;
        p.set(d.s, d.sp, fp);
        const bool cut = !(cfg.flags & kFlagNoFcut);
        const double p13 = 1.4645918875615232630201425272637904 * p.x13;
        double pd[4], dv[4], dtau[2], v, T, Ty, Q, xQp, R, xRp;
        phenomd_phase(d, 0, p, cut, o.phi[0], pd);
        if (!cut || p.x < kMfCut) {
            nrt_phase_shape(p13, R, xRp);
            o.phi[0] += r.kph[0] * R;
        }
        phenomd_amp_core(d, p, true, v, dv);
        nrt_taper(p.x, r.ym[0], T, Ty);
        nrt_amp_shape(p13, p.lpx3, Q, xQp);
        o.amp[0] = d.C * fma(r.sm76 * fp.fm76, v, r.kam[0] * Q) * T;
        tau_eval(d.tau, p.xm13 * 0.68278406325529568146702083315816, p.lpx3, d.lam, o.tau, dtau);
    }
};
template <> struct WaveformFns<kNSBH> {
    static constexpr int kModes = 1;
    static GWF_HD void eval(const NSBHRec<4>& r, const ModelCfg& cfg, const HMWeights&, const FreqPoint& fp, WaveformOut& o) {
        PointWf<4> w;
        ModelTraits<kNSBH, 4>::eval(r, cfg, 0, fp, true, w);
        o.phi[0] = w.phi;
        o.amp[0] = w.A;
        o.tau = w.tau;
    }
};
template <> struct WaveformFns<kPhenomHM> {
    static constexpr int kModes = kHMModes;
    static GWF_HD void eval(const HMRec<4>& r, const ModelCfg& cfg, const HMWeights& w, const FreqPoint& fp, WaveformOut& o) {
        const bool cut = !(cfg.flags & kFlagNoFcut);
        phenomhm_amp_phase<4>(r, 0, fp, cut, o.amp, o.phi);
        o.hp[0] = o.hp[1] = o.hc[0] = o.hc[1] = 0.;
        for (int m = 0; m < kHMModes; ++m) {
            double sn, cs;
            sincos(o.phi[m], &sn, &cs);
            const double zr = o.amp[m] * cs, zi = -o.amp[m] * sn;
            o.hp[0] = fma(zr, w.wp[m], o.hp[0]); o.hp[1] = fma(zi, w.wp[m], o.hp[1]);
            o.hc[0] = fma(-zi, w.wc[m], o.hc[0]); o.hc[1] = fma(zr, w.wc[m], o.hc[1]);
        }
        double dtau[2];
        const double x13 = cbrt(r.s * fp.f), lpx3 = log(kPi * r.s * fp.f) * (1. / 3.);
        tau_eval(r.tau, 0.68278406325529568146702083315816 / x13, lpx3, r.lam, o.tau, dtau);
    }
};

}  // namespace gwf
