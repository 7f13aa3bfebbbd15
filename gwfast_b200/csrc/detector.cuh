// Detector response fused into the Fisher/SNR kernels: pattern functions (gwfast/signal.py:342-387), Earth-centre
// to site delay (signal.py:401-423), Earth-rotation time t(f) (signal.py:444-453, 564-580), the analytic
// derivative rows of signal.py:1303-1584, and the packed Gram accumulation of signal.py:922-931.
//
// Everything is expressed through  a(ang,dec), b(ang,dec)  with ang = ra - long - 2 pi t  and their partial
// derivatives in (ang, dec); a and b are linear in (sin 2(xax+rot), cos 2(xax+rot)), so an "arm" is just a
// coefficient pair (S2, C2) (with sin(angbtwArms) folded in) -- including the virtual arms used for the
// triangle (arm3 = -(arm1+arm2), signal.py:1057, and the u=r1+r2, v=r1-r2 two-Gram form of the 3-arm sum).
#pragma once
#include "model_common.cuh"

namespace gwf {

constexpr int kMaxDet = 8;
constexpr int kMaxArms = 16;
constexpr int kMaxPsd = 8;

struct PsdDev {
    const double4* tab;   // (f_j, S_j, slope_j, 0), slope_j = (S_{j+1}-S_j)/(f_{j+1}-f_j) as np.interp forms it
    const int* bucket;    // bucket[b] = largest j with f_j <= lower edge of log2-bucket b (clamped to [0, n-2])
    int n, nb;
    double lo, inv;       // bucket coordinate = (log2 f - lo) * inv
    double f_first, f_last;
};

struct DetDev {
    double sl, cl, s2l, c2l;   // sin/cos of latitude and of 2*latitude
    double slon, clon;         // sin/cos of longitude
    double fmin, fmax;         // Hz; fmax <= 0 means "no clip"
    int group;                 // frequency-grid group
    int psd;                   // index into psd[]
    int arm_begin, arm_end;    // arms[arm_begin:arm_end]
    int use_rot;               // useEarthMotion
    int no_motion;             // noMotion
};

struct ArmDev {
    double S2, C2;    // sin(angbtwArms) * sin/cos(2 (xax + rot))   (linear combinations for virtual arms)
    double weight;    // Gram weight (1, or 3/2 and 1/2 for the u/v form)
    int out;          // SNR kernel: output slot of this arm
    int pad;
};

struct NetworkDev {
    int ndet, narms, ngroups, npsd;
    int group_rot[kMaxGroups];      // any detector of the group uses the Earth rotation
    double group_fmin[kMaxGroups];
    double group_fmax[kMaxGroups];
    DetDev det[kMaxDet];
    ArmDev arm[kMaxArms];
    PsdDev psd[kMaxPsd];
};

// np.interp(f, strainFreq, noiseCurve, left=1., right=1.)  (signal.py:723, 901) with an O(1) bucket start.
// l2f = log2(f) is supplied by the caller (known in closed form on a geometric grid).
#ifdef __CUDA_ARCH__
#define GWF_LDG(ptr) __ldg(ptr)
#else
#define GWF_LDG(ptr) (*(ptr))
#endif
GWF_HD double psd_lookup(const PsdDev& p, double f, double l2f) {
    if (!(f >= p.f_first) || f > p.f_last) return 1.0;
    int b = (int)((l2f - p.lo) * p.inv);
    b = b < 1 ? 1 : (b > p.nb - 1 ? p.nb - 1 : b);
    int lo = GWF_LDG(p.bucket + b - 1);                    // one bucket of slack on each side against rounding of l2f
    int hi = (b + 2 < p.nb ? GWF_LDG(p.bucket + b + 2) : p.n - 2) + 1;
    hi = hi > p.n - 1 ? p.n - 1 : hi;
    // largest j in [lo, hi) with f_j <= f
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (GWF_LDG(&p.tab[mid].x) <= f) lo = mid; else hi = mid;
    }
    const double2* row = reinterpret_cast<const double2*>(p.tab + lo);
    const double2 fs = GWF_LDG(row), sl = GWF_LDG(row + 1);      // (f_j, S_j), (slope_j, 0): two 16-byte loads
    return fma(sl.x, f - fs.x, fs.y);
}

// per-event sky / orientation constants
struct EvGeom {
    double sd, cd, s2d, c2d;     // declination = pi/2 - theta
    double sra, cra;
    double c2psi, s2psi;
    double ci, si, K;            // cos iota, sin iota, (1 + cos^2 iota)/2
    double tcoal, inv_dL;
    GWF_HD void set(const EventIn& e) {
        // dec = pi/2 - theta (gwfastUtils.py:151-164): sin dec = cos theta, cos dec = sin theta
        const double dec = 0.5 * kPi - e.theta;
        sincos(dec, &sd, &cd);
        sincos(2.0 * dec, &s2d, &c2d);
        sincos(e.phi, &sra, &cra);
        sincos(2.0 * e.psi, &s2psi, &c2psi);
        sincos(e.iota, &si, &ci);
        K = 0.5 * (1.0 + ci * ci);
        tcoal = e.tcoal;
        inv_dL = 1.0 / e.dL;
    }
};

// a/b basis of one detector at one time: values, d/d(ang), d/d(dec); "S"/"C" = coefficient of S2 / C2
struct DetPoint {
    double aS, aC, bS, bC, aS_g, aC_g, bS_g, bC_g, aS_d, aC_d, bS_d, bC_d;
    double dt_th, dt_ph, dt_tn;   // d(Delta t)/d theta, /d phi [s], /d t_noloc [s per day]
    double dt;                    // Delta t [s]
};

// tn = detector-independent time in days BEFORE the Earth-centre->site delay; (cB, sB) = cos/sin(2 pi tn)
GWF_HD void det_point(const DetDev& d, const EvGeom& g, double cB, double sB, DetPoint& o) {
    const double Rc = kREarth / kClight;
    // A = ra - lon
    const double cA = g.cra * d.clon + g.sra * d.slon, sA = g.sra * d.clon - g.cra * d.slon;
    // ang0 = (ra - lon) - 2 pi tn
    const double c0 = cA * cB + sA * sB, s0 = sA * cB - cA * sB;
    o.dt = -Rc * (g.cd * d.cl * c0 + g.sd * d.sl);                       // signal.py:417-421
    o.dt_th = -Rc * (g.sd * d.cl * c0 - g.cd * d.sl);                    // signal.py:1483-1491
    o.dt_ph = Rc * g.cd * d.cl * s0;                                     // signal.py:1441-1448
    o.dt_tn = -2.0 * kPi * Rc * g.cd * d.cl * s0;                        // signal.py:1527-1534
    // ang = ang0 - 2 pi dt/86400 ; the shift is ~1e-6 rad: 3rd-order Taylor is exact to 1e-24
    const double del = 2.0 * kPi * o.dt / kDay;
    const double cdl = 1.0 - 0.5 * del * del, sdl = del * (1.0 - del * del * (1. / 6.));
    const double c1 = c0 * cdl + s0 * sdl, s1 = s0 * cdl - c0 * sdl;
    const double c2 = c1 * c1 - s1 * s1, s2 = 2.0 * s1 * c1;
    // signal.py:360-376 regrouped by (S2, C2)
    const double k3 = 3.0 - d.c2l, e3 = 3.0 - g.c2d;
    const double p1 = 0.0625 * k3 * e3, p2 = 0.25 * d.s2l * g.s2d, p3 = 0.75 * d.cl * d.cl * g.cd * g.cd;
    const double p4 = 0.25 * d.sl * e3, p5 = 0.5 * d.cl * g.s2d;
    const double q1 = d.sl * g.sd, q2 = d.cl * g.cd, q3 = 0.25 * k3 * g.sd, q4 = 0.5 * d.s2l * g.cd;
    o.aS = p1 * c2 + p2 * c1 + p3;
    o.aC = -(p4 * s2 + p5 * s1);
    o.bC = q1 * c2 + q2 * c1;
    o.bS = q3 * s2 + q4 * s1;
    o.aS_g = -2.0 * p1 * s2 - p2 * s1;
    o.aC_g = -(2.0 * p4 * c2 + p5 * c1);
    o.bC_g = -2.0 * q1 * s2 - q2 * s1;
    o.bS_g = 2.0 * q3 * c2 + q4 * c1;
    // d/d(dec) of the p's and q's
    const double p1d = 0.125 * k3 * g.s2d, p2d = 0.5 * d.s2l * g.c2d, p3d = -0.75 * d.cl * d.cl * g.s2d;
    const double p4d = 0.5 * d.sl * g.s2d, p5d = d.cl * g.c2d;
    const double q1d = d.sl * g.cd, q2d = -d.cl * g.sd, q3d = 0.25 * k3 * g.cd, q4d = -0.5 * d.s2l * g.sd;
    o.aS_d = p1d * c2 + p2d * c1 + p3d;
    o.aC_d = -(p4d * s2 + p5d * s1);
    o.bC_d = q1d * c2 + q2d * c1;
    o.bS_d = q3d * s2 + q4d * s1;
}

// value-only pattern functions of one arm (SNR path)
GWF_HD void arm_pattern(const DetPoint& p, const ArmDev& a, const EvGeom& g, double& Fp, double& Fc) {
    const double av = a.S2 * p.aS + a.C2 * p.aC, bv = a.C2 * p.bC + a.S2 * p.bS;
    Fp = av * g.c2psi + bv * g.s2psi;
    Fc = bv * g.c2psi - av * g.s2psi;
}

template <int NT>
struct PointWf {
    double f;            // Hz
    double A;            // waveform amplitude
    double lnA_d[NT];
    double phi_d[NT];
    double dtn[2];       // d t_noloc / d slot (days) = -dtau/86400 when the Earth rotates, else 0
    double tau;          // time to coalescence [s] (only when requested)
};

// packed lower-triangular index
GWF_HD constexpr int tri(int i, int j) { return i * (i + 1) / 2 + j; }

// rows of d h / d p (divided by A e^{i Psi}) for one arm and their weighted Gram; NP = NT + 7.
// Row order = ParNums (waveforms.py:78): Mc eta dL theta phi iota psi tcoal Phicoal chi1z chi2z [LambdaTilde deltaLambda]
template <int NT>
GWF_HD void arm_rows_accumulate(const PointWf<NT>& w, const DetPoint& p, const DetDev& d, const ArmDev& a, const EvGeom& g, double wgt,
                                double* __restrict__ acc, double& snr2) {
    constexpr int NP = NT + 7;
    const double av = a.S2 * p.aS + a.C2 * p.aC, bv = a.C2 * p.bC + a.S2 * p.bS;
    const double ag = a.S2 * p.aS_g + a.C2 * p.aC_g, bg = a.C2 * p.bC_g + a.S2 * p.bS_g;
    const double ad = a.S2 * p.aS_d + a.C2 * p.aC_d, bd = a.C2 * p.bC_d + a.S2 * p.bS_d;
    const double Fp = av * g.c2psi + bv * g.s2psi, Fc = bv * g.c2psi - av * g.s2psi;
    const double Gr = Fp * g.K, Gi = Fc * g.ci;                                       // signal.py:463-464
    const double Ggr = (ag * g.c2psi + bg * g.s2psi) * g.K, Ggi = (bg * g.c2psi - ag * g.s2psi) * g.ci;
    const double Gdr = (ad * g.c2psi + bd * g.s2psi) * g.K, Gdi = (bd * g.c2psi - ad * g.s2psi) * g.ci;
    const double W2 = 2.0 * kPi * w.f;
    const double twopi = 2.0 * kPi;
    const double tfac = d.no_motion ? 0.0 : 1.0 + p.dt_tn / kDay;       // d t / d t_noloc
    double ra[NP], rb[NP];
    // intrinsic rows (AD rows of the reference, signal.py:1153-1189)
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        const int row = j < 2 ? j : 7 + j;
        double ang_x = 0., psi_x = -w.phi_d[j];
        if (j < 2) {
            ang_x = -twopi * w.dtn[j] * tfac;
            psi_x = fma(W2 * p.dt_tn, w.dtn[j], psi_x);
        }
        ra[row] = w.lnA_d[j] * Gr + Ggr * ang_x - Gi * psi_x;
        rb[row] = w.lnA_d[j] * Gi + Ggi * ang_x + Gr * psi_x;
    }
    // dL, signal.py:1576
    ra[2] = -Gr * g.inv_dL;
    rb[2] = -Gi * g.inv_dL;
    // theta, signal.py:1482-1523 (dec = pi/2 - theta)
    {
        const double ang_t = -twopi * p.dt_th / kDay, ph = W2 * p.dt_th;
        ra[3] = Ggr * ang_t - Gdr - Gi * ph;
        rb[3] = Ggi * ang_t - Gdi + Gr * ph;
    }
    // phi, signal.py:1439-1480
    {
        const double ang_p = 1.0 - twopi * p.dt_ph / kDay, ph = W2 * p.dt_ph;
        ra[4] = Ggr * ang_p - Gi * ph;
        rb[4] = Ggi * ang_p + Gr * ph;
    }
    // iota, signal.py:1567-1571
    ra[5] = -Fp * g.ci * g.si;
    rb[5] = -Fc * g.si;
    // psi, signal.py:1432-1437
    ra[6] = 2.0 * Fc * g.K;
    rb[6] = -2.0 * Fp * g.ci;
    // tcoal, in 1/s (signal.py:1525-1565 and :920)
    {
        const double ang_c = -twopi * tfac / kDay, ph = W2 * (1.0 + (d.no_motion ? 0.0 : p.dt_tn / kDay));
        ra[7] = Ggr * ang_c - Gi * ph;
        rb[7] = Ggi * ang_c + Gr * ph;
    }
    // Phicoal, signal.py:1577
    ra[8] = Gi;
    rb[8] = -Gr;
    // 4 Re int conj(d_a h) d_b h / Sn df, signal.py:922-931
    const double wg = wgt * a.weight;
    snr2 = fma(wg, Gr * Gr + Gi * Gi, snr2);
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        const double wa = wg * ra[i], wb = wg * rb[i];
#pragma unroll
        for (int j = 0; j <= i; ++j) acc[tri(i, j)] = fma(wa, ra[j], fma(wb, rb[j], acc[tri(i, j)]));
    }
}

}  // namespace gwf
