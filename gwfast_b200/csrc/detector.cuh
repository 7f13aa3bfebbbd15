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
    // log-uniform node spacing (every PSD file shipped with gwfast): row index ~ (log2 f - u_lo) * u_inv, no bucket load
    int uni;
    double u_lo, u_inv;
    // window [c_j0, c_j0 + c_n) of the table rows cached in the CTA's dynamic shared memory at byte offset c_off (kernels
    // that call psd_cache_fill; planned per launch by plan_psd_cache, -1 = not cached); same 32-byte rows as `tab`
    int c_off, c_j0, c_n;
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

constexpr int kMaxFastDet = 4;
struct NetworkDev {
    int ndet, narms, ngroups, npsd;
    int group_rot[kMaxGroups];      // any detector of the group uses the Earth rotation
    double group_fmin[kMaxGroups];
    double group_fmax[kMaxGroups];
    DetDev det[kMaxDet];
    ArmDev arm[kMaxArms];
    PsdDev psd[kMaxPsd];
    // "fast" form of a single-group network with at most kMaxFastDet active detectors whose PSD windows all sit in shared
    // memory: the active detectors compacted to fdet[0..fnd) with their PSD descriptor copied next to them, so that a
    // fully unrolled detector loop reads every field as a kernel-parameter operand at a compile-time offset (no indexed
    // constant loads, no group / empty-detector / cache-miss branches).  fast = 0: not available (generic loop).
    int fast, fnd;
    DetDev fdet[kMaxFastDet];
    PsdDev fpsd[kMaxFastDet];
};

// np.interp(f, strainFreq, noiseCurve, left=1., right=1.)  (signal.py:723, 901).  The table row j is
// (f_j, S_j, slope_j, f_{j+1}); a log2-bucket index gives the row in O(1) (each bucket holds at most a few nodes),
// so the common case is one 32-byte row read by two 16-byte loads.  l2f = log2(f) comes from the caller (closed
// form on a geometric grid).
#ifdef __CUDA_ARCH__
#define GWF_LDG(ptr) __ldg(ptr)
#else
#define GWF_LDG(ptr) (*(ptr))
#endif
#ifdef __CUDA_ARCH__
extern __shared__ __align__(16) unsigned char gwf_dyn_smem[];
#endif
GWF_HD double psd_lookup(const PsdDev& p, double f, double l2f) {
    if (!(f >= p.f_first) || f > p.f_last) return 1.0;
#ifdef __CUDA_ARCH__
    if (p.c_off >= 0) {
        // shared-memory window: the whole row (f_j, S_j, slope_j, f_{j+1}) arrives with two 16-byte loads
        const double2* R = reinterpret_cast<const double2*>(gwf_dyn_smem + p.c_off);
        const int last = p.c_n - 1;
        int j = (int)((l2f - p.u_lo) * p.u_inv) - p.c_j0;
        j = j < 0 ? 0 : (j > last ? last : j);
        double2 fs = R[2 * j], sn = R[2 * j + 1];
        while (j > 0 && f < fs.x) { --j; fs = R[2 * j]; sn = R[2 * j + 1]; }        // rounding of the index guess (rare)
        while (j < last && f >= sn.y) { ++j; fs = R[2 * j]; sn = R[2 * j + 1]; }
        return fma(sn.x, f - fs.x, fs.y);
    }
#endif
    int b = (int)((l2f - p.lo) * p.inv);
    b = b < 0 ? 0 : (b > p.nb - 1 ? p.nb - 1 : b);
    int j = GWF_LDG(p.bucket + b);
    const double2* row = reinterpret_cast<const double2*>(p.tab + j);
    double2 fs = GWF_LDG(row), sn = GWF_LDG(row + 1);      // (f_j, S_j), (slope_j, f_{j+1})
    while (f < fs.x && j > 0) {                            // rounding of l2f put us one bucket too high (rare)
        --j;
        row = reinterpret_cast<const double2*>(p.tab + j);
        fs = GWF_LDG(row); sn = GWF_LDG(row + 1);
    }
    while (f >= sn.y && j < p.n - 2) {                     // more than one node inside the bucket
        ++j;
        row = reinterpret_cast<const double2*>(p.tab + j);
        fs = GWF_LDG(row); sn = GWF_LDG(row + 1);
    }
    return fma(sn.x, f - fs.x, fs.y);
}

#ifdef __CUDA_ARCH__
// the same interpolation for a table known to be log-uniform and resident in shared memory (NetworkDev::fpsd): no
// cache-miss path, the out-of-range case is a select, and the index fix-up is one rarely taken branch
__device__ __forceinline__ double psd_lookup_fast(const PsdDev& p, double f, double l2f) {
    const double2* R = reinterpret_cast<const double2*>(gwf_dyn_smem + p.c_off);
    const int last = p.c_n - 1;
    int j = (int)((l2f - p.u_lo) * p.u_inv) - p.c_j0;
    j = max(0, min(j, last));
    // The nodes of a "uni" table sit within a quarter of a bucket of the uniform positions (host_build.h), so the segment
    // that holds f is j - 1, j or j + 1: the step is a pair of selects on the node frequencies of row j and the row is read
    // again -- straight-line code the scheduler can spread over the surrounding arithmetic (the former fix-up branch waited
    // on its shared-memory row: 8 % of the Fisher kernel's stall samples, profiles/r01f).  f == node picks the segment to
    // its right, like np.interp.
    const double f_lo = R[2 * j].x, f_hi = R[2 * j + 1].y;
    j += (j < last && f >= f_hi) ? 1 : ((j > 0 && f < f_lo) ? -1 : 0);
    const double2 fs = R[2 * j], sn = R[2 * j + 1];
    const double v = fma(sn.x, f - fs.x, fs.y);
    return (f >= p.f_first && f <= p.f_last) ? v : 1.0;
}
#endif

// sin / cos of the Earth-rotation angle 2 pi t, t in days.  On the device the argument is reduced in turns (sincospi: exact
// reduction, no slow path for large arguments -- sincos() carries a Payne-Hanek branch that split every sample's code in two)
GWF_HD void rot_sincos(double t_days, double* s, double* c) {
#ifdef __CUDA_ARCH__
    sincospi(2.0 * t_days, s, c);
#else
    sincos(2.0 * kPi * t_days, s, c);
#endif
}

// per-event sky / orientation constants
struct EvGeom {
    double sd, cd, s2d, c2d;     // declination = pi/2 - theta
    double sra, cra;
    double c2psi, s2psi;
    double ci, si, K;            // cos iota, sin iota, (1 + cos^2 iota)/2
    double tcoal, inv_dL;
    double cT, sT;               // cos/sin(2 pi tcoal): detectors that do not follow the Earth rotation
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
        sincos(2.0 * kPi * e.tcoal, &sT, &cT);
    }
};

constexpr double kInvDay = 1.0 / kDay;
constexpr double kRc = kREarth / kClight;          // Earth radius in light-seconds

// f-independent products of the event's sky position with one detector's site constants
struct EvDet {
    double cA, sA;                 // cos/sin(ra - lon)
    double kc, k0;                 // Delta t      = -Rc (kc  c0 + k0 ),  kc  = cos(dec) cos(lat), k0  = sin(dec) sin(lat)
    double kc_th, k0_th;           // d Delta t/d theta = -Rc (kc_th c0 + k0_th)
    double p[5], q[4], pd[5], qd[4];   // a/b coefficient groups of signal.py:360-376 and their d/d(dec)
    GWF_HD void set(const DetDev& d, const EvGeom& g) {
        cA = g.cra * d.clon + g.sra * d.slon;
        sA = g.sra * d.clon - g.cra * d.slon;
        kc = g.cd * d.cl; k0 = g.sd * d.sl;
        kc_th = g.sd * d.cl; k0_th = -g.cd * d.sl;
        const double k3 = 3.0 - d.c2l, e3 = 3.0 - g.c2d;
        p[0] = 0.0625 * k3 * e3; p[1] = 0.25 * d.s2l * g.s2d; p[2] = 0.75 * d.cl * d.cl * g.cd * g.cd;
        p[3] = 0.25 * d.sl * e3; p[4] = 0.5 * d.cl * g.s2d;
        q[0] = d.sl * g.sd; q[1] = d.cl * g.cd; q[2] = 0.25 * k3 * g.sd; q[3] = 0.5 * d.s2l * g.cd;
        pd[0] = 0.125 * k3 * g.s2d; pd[1] = 0.5 * d.s2l * g.c2d; pd[2] = -0.75 * d.cl * d.cl * g.s2d;
        pd[3] = 0.5 * d.sl * g.s2d; pd[4] = d.cl * g.c2d;
        qd[0] = d.sl * g.cd; qd[1] = -d.cl * g.sd; qd[2] = 0.25 * k3 * g.cd; qd[3] = -0.5 * d.s2l * g.sd;
    }
};

// a/b basis of one detector at one time: values, d/d(ang), d/d(dec); "S"/"C" = coefficient of S2 / C2
struct DetPoint {
    double aS, aC, bS, bC, aS_g, aC_g, bS_g, bC_g, aS_d, aC_d, bS_d, bC_d;
    double dt_th, dt_ph, dt_tn;   // d(Delta t)/d theta, /d phi [s], /d t_noloc [s per day]
    double dt;                    // Delta t [s]
};

// a/b basis at ang = (ra - lon) - 2 pi t given as (c1, s1) = cos/sin(ang)  (signal.py:360-376)
GWF_HD void det_basis(const EvDet& e, double c1, double s1, DetPoint& o);
// (cB, sB) = cos/sin(2 pi tn), tn = time in days BEFORE the Earth-centre->site delay is added (signal.py:444-453)
GWF_HD void det_point(const EvDet& e, double cB, double sB, DetPoint& o) {
    // ang0 = (ra - lon) - 2 pi tn
    const double c0 = e.cA * cB + e.sA * sB, s0 = e.sA * cB - e.cA * sB;
    o.dt = -kRc * (e.kc * c0 + e.k0);                                    // signal.py:417-421
    o.dt_th = -kRc * (e.kc_th * c0 + e.k0_th);                           // signal.py:1483-1491
    o.dt_ph = kRc * e.kc * s0;                                           // signal.py:1441-1448
    o.dt_tn = -2.0 * kPi * o.dt_ph;                                      // signal.py:1527-1534
    // ang = ang0 - 2 pi dt/86400 ; the shift is ~1e-6 rad: 3rd-order Taylor is exact to 1e-24
    const double del = (2.0 * kPi * kInvDay) * o.dt;
    const double cdl = 1.0 - 0.5 * del * del, sdl = del * (1.0 - del * del * (1. / 6.));
    const double c1 = c0 * cdl + s0 * sdl, s1 = s0 * cdl - c0 * sdl;
    det_basis(e, c1, s1, o);
}
GWF_HD void det_basis(const EvDet& e, double c1, double s1, DetPoint& o) {
    const double c2 = c1 * c1 - s1 * s1, s2 = 2.0 * s1 * c1;
    o.aS = e.p[0] * c2 + e.p[1] * c1 + e.p[2];
    o.aC = -(e.p[3] * s2 + e.p[4] * s1);
    o.bC = e.q[0] * c2 + e.q[1] * c1;
    o.bS = e.q[2] * s2 + e.q[3] * s1;
    o.aS_g = -2.0 * e.p[0] * s2 - e.p[1] * s1;
    o.aC_g = -(2.0 * e.p[3] * c2 + e.p[4] * c1);
    o.bC_g = -2.0 * e.q[0] * s2 - e.q[1] * s1;
    o.bS_g = 2.0 * e.q[2] * c2 + e.q[3] * c1;
    o.aS_d = e.pd[0] * c2 + e.pd[1] * c1 + e.pd[2];
    o.aC_d = -(e.pd[3] * s2 + e.pd[4] * s1);
    o.bC_d = e.qd[0] * c2 + e.qd[1] * c1;
    o.bS_d = e.qd[2] * s2 + e.qd[3] * s1;
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
    double phi;          // waveform phase Phi(f) (only read by the derivative write-out kernel)
};

// packed lower-triangular index
GWF_HD constexpr int tri(int i, int j) { return i * (i + 1) / 2 + j; }

// detector-level chain-rule factors shared by the arms of one detector at one frequency
template <int NT>
struct DetRows {
    double ang_x[2], psi_x[NT];    // d ang / d slot (slots 0,1), d Psi / d slot
    double ang_t, ph_t, ang_p, ph_p, ang_c, ph_c;
    GWF_HD void set(const PointWf<NT>& w, const DetPoint& p, bool follows_rotation, bool no_motion) {
        const double W2 = 2.0 * kPi * w.f, twopi = 2.0 * kPi;
        const double tfac = no_motion ? 0.0 : fma(p.dt_tn, kInvDay, 1.0);          // d t / d t_noloc
#pragma unroll
        for (int j = 0; j < NT; ++j) psi_x[j] = -w.phi_d[j];
        ang_x[0] = ang_x[1] = 0.;
        if (follows_rotation) {
            const double wd = W2 * p.dt_tn;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                ang_x[j] = -twopi * w.dtn[j] * tfac;
                psi_x[j] = fma(wd, w.dtn[j], psi_x[j]);
            }
        }
        ang_t = -(twopi * kInvDay) * p.dt_th; ph_t = W2 * p.dt_th;                 // signal.py:1482-1523
        ang_p = 1.0 - (twopi * kInvDay) * p.dt_ph; ph_p = W2 * p.dt_ph;            // signal.py:1439-1480
        ang_c = -(twopi * kInvDay) * tfac;                                         // signal.py:1525-1565, per second (:920)
        ph_c = no_motion ? W2 : W2 * tfac;
    }
};

// ---------------------------------------------------------------------------------------------------------------
// Compact Gram of the (2,2)-only models.  Four of the rows are fixed per-event combinations of the pattern functions
// u = F+ and v = Fx alone:  dL: (-K/dL u, -c/dL v),  iota: (-c s u, -s v),  psi: (2K v, -2c u),  Phicoal: (c v, -K u)
// (K = (1+c^2)/2, c = cos iota, s = sin iota; signal.py:1432-1437, 1567-1577).  Instead of 4 full rows of the packed Gram
// they need only, per general row x = (a_x, b_x):  P = sum w u a_x, Q = sum w v b_x, R = sum w v a_x, S = sum w u b_x,
// plus UU, VV, UV -- 59 accumulators and 103 FP64 ops per arm-sample instead of 66 and 154 (84 vs 91 for 13 parameters).
// The packed Fisher entries are rebuilt once per event (compact_entry).
template <int NT>
struct Compact {
    static constexpr int NP = NT + 7;
    static constexpr int NG = NT + 3;                      // general rows: Mc eta theta phi tcoal chi1 chi2 [LambdaTilde deltaLambda]
    static constexpr int kGG = NG * (NG + 1) / 2;
    static constexpr int kUU = kGG + 4 * NG, kVV = kUU + 1, kUV = kUU + 2;
    static constexpr int kAcc = kUU + 3;
    GWF_HD static constexpr int row_of(int g) { return g < 2 ? g : (g == 2 ? 3 : (g == 3 ? 4 : (g == 4 ? 7 : 4 + g))); }
    // ParNums row -> general index, or -1 (dL), -2 (iota), -3 (psi), -4 (Phicoal)
    GWF_HD static constexpr int g_of(int row) {
        return row < 2 ? row : (row == 2 ? -1 : (row == 3 ? 2 : (row == 4 ? 3 : (row == 5 ? -2 : (row == 6 ? -3 : (row == 7 ? 4 : (row == 8 ? -4 : row - 4)))))));
    }
};

// Fisher entry (i, j), i >= j in ParNums order, from the reduced compact accumulators
template <int NT>
GWF_HD double compact_entry(int i, int j, const double* __restrict__ acc, const EvGeom& g) {
    typedef Compact<NT> C;
    const int gi = C::g_of(i), gj = C::g_of(j);
    // (alpha, beta): type I rows are (alpha u, beta v); type II rows are (beta v, alpha u)
    const double al[4] = {-g.K * g.inv_dL, -g.ci * g.si, -2.0 * g.ci, -g.K};
    const double be[4] = {-g.ci * g.inv_dL, -g.si, 2.0 * g.K, g.ci};
    if (gi >= 0 && gj >= 0) return acc[tri(gi > gj ? gi : gj, gi > gj ? gj : gi)];
    if (gi >= 0 || gj >= 0) {
        const int x = gi >= 0 ? gi : gj, sidx = -(gi >= 0 ? gj : gi) - 1;
        const double* c = acc + C::kGG + 4 * x;            // P, Q, R, S
        return sidx < 2 ? al[sidx] * c[0] + be[sidx] * c[1] : be[sidx] * c[2] + al[sidx] * c[3];
    }
    const int si = -gi - 1, sj = -gj - 1;
    const bool Ii = si < 2, Ij = sj < 2;
    if (Ii && Ij) return al[si] * al[sj] * acc[C::kUU] + be[si] * be[sj] * acc[C::kVV];
    if (!Ii && !Ij) return be[si] * be[sj] * acc[C::kVV] + al[si] * al[sj] * acc[C::kUU];
    const int a1 = Ii ? si : sj, a2 = Ii ? sj : si;       // a1 type I, a2 type II
    return (al[a1] * be[a2] + be[a1] * al[a2]) * acc[C::kUV];
}
// The same rebuild in table form, for the kernels: entry(i, j) = coef[ka] acc[ia] + coef[kb] acc[ib], where the codes (ia, ib, ka,
// kb) depend only on (i, j) -- computed once per kernel into shared memory -- and coef is a per-event vector of 58 products of
// (alpha, beta): 0: 1, 1: 0, 2+s: alpha_s, 6+s: beta_s, 10+4s+t: alpha_s alpha_t, 26+4s+t: beta_s beta_t,
// 42+4s+t: alpha_s beta_t + beta_s alpha_t.  (compact_entry's chain of selects, run three times per lane, both halves of every
// event, was 3 % of the Fisher kernel's stall samples, profiles/r01h.)
constexpr int kCompactCoefs = 58;
template <int NT>
GWF_HD void compact_entry_code(int i, int j, unsigned char* code) {
    typedef Compact<NT> C;
    const int gi = C::g_of(i), gj = C::g_of(j);
    int ia, ib, ka, kb;
    if (gi >= 0 && gj >= 0) {
        ia = ib = tri(gi > gj ? gi : gj, gi > gj ? gj : gi); ka = 0; kb = 1;
    } else if (gi >= 0 || gj >= 0) {
        const int x = gi >= 0 ? gi : gj, sidx = -(gi >= 0 ? gj : gi) - 1, base = C::kGG + 4 * x;
        if (sidx < 2) { ia = base; ka = 2 + sidx; ib = base + 1; kb = 6 + sidx; }
        else { ia = base + 2; ka = 6 + sidx; ib = base + 3; kb = 2 + sidx; }
    } else {
        const int si = -gi - 1, sj = -gj - 1;
        const bool Ii = si < 2, Ij = sj < 2;
        if (Ii == Ij) { ia = C::kUU; ka = 10 + 4 * si + sj; ib = C::kVV; kb = 26 + 4 * si + sj; }
        else { ia = ib = C::kUV; ka = 42 + 4 * (Ii ? si : sj) + (Ii ? sj : si); kb = 1; }
    }
    code[0] = (unsigned char)ia; code[1] = (unsigned char)ib; code[2] = (unsigned char)ka; code[3] = (unsigned char)kb;
}
// coefficient k of the per-event vector (see compact_entry_code)
GWF_HD double compact_coef(int k, const EvGeom& g) {
    const double al[4] = {-g.K * g.inv_dL, -g.ci * g.si, -2.0 * g.ci, -g.K};
    const double be[4] = {-g.ci * g.inv_dL, -g.si, 2.0 * g.K, g.ci};
    if (k < 2) return k == 0 ? 1.0 : 0.0;
    if (k < 6) return al[k - 2];
    if (k < 10) return be[k - 6];
    const int q = (k - 10) & 15, s = q >> 2, t = q & 3;
    if (k < 26) return al[s] * al[t];
    if (k < 42) return be[s] * be[t];
    return al[s] * be[t] + be[s] * al[t];
}

template <int NT>
GWF_HD double compact_snr2(const double* __restrict__ acc, const EvGeom& g) {
    return g.K * g.K * acc[Compact<NT>::kUU] + g.ci * g.ci * acc[Compact<NT>::kVV];     // 4 sum w |h|^2 / Sn, signal.py:727
}

// (h | d_row h) = 4 Re int conj(d_row h) h / Sn df from the same accumulators (return_SNR_derivatives, signal.py:938-945):
// h = A e^{i Psi} (K u + i c v), so a general row contributes K P + c Q and the fixed-combination rows only UU, VV, UV
template <int NT>
GWF_HD double compact_snr_deriv(int row, const double* __restrict__ acc, const EvGeom& g) {
    typedef Compact<NT> C;
    const int gx = C::g_of(row);
    if (gx >= 0) return g.K * acc[C::kGG + 4 * gx] + g.ci * acc[C::kGG + 4 * gx + 1];
    const double al[4] = {-g.K * g.inv_dL, -g.ci * g.si, -2.0 * g.ci, -g.K};
    const double be[4] = {-g.ci * g.inv_dL, -g.si, 2.0 * g.K, g.ci};
    const int sidx = -gx - 1;
    if (sidx < 2) return al[sidx] * g.K * acc[C::kUU] + be[sidx] * g.ci * acc[C::kVV];     // rows (alpha u, beta v)
    return (be[sidx] * g.K + al[sidx] * g.ci) * acc[C::kUV];                               // rows (beta v, alpha u)
}

// general rows of d h / d p (divided by A e^{i Psi}) for one arm: ra + i rb for the NG general parameters, and the
// pattern functions Fp, Fc that generate the four fixed-combination rows (see Compact)
// UNIT: the arm's coefficient pair is known at compile time -- 1: (S2, C2) = (1, 0), 2: (0, 1), the two virtual arms of a summed
// triangle (host_build.h:build_network) -- and the six linear combinations below are plain selections
template <int NT, int UNIT = 0>
GWF_HD void arm_rows(const PointWf<NT>& w, const DetPoint& p, const DetRows<NT>& dr, const ArmDev& a, const EvGeom& g,
                     double* __restrict__ ra, double* __restrict__ rb, double& Fp, double& Fc) {
    const double av = UNIT == 1 ? p.aS : (UNIT == 2 ? p.aC : a.S2 * p.aS + a.C2 * p.aC), bv = UNIT == 1 ? p.bS : (UNIT == 2 ? p.bC : a.C2 * p.bC + a.S2 * p.bS);
    const double ag = UNIT == 1 ? p.aS_g : (UNIT == 2 ? p.aC_g : a.S2 * p.aS_g + a.C2 * p.aC_g), bg = UNIT == 1 ? p.bS_g : (UNIT == 2 ? p.bC_g : a.C2 * p.bC_g + a.S2 * p.bS_g);
    const double ad = UNIT == 1 ? p.aS_d : (UNIT == 2 ? p.aC_d : a.S2 * p.aS_d + a.C2 * p.aC_d), bd = UNIT == 1 ? p.bS_d : (UNIT == 2 ? p.bC_d : a.C2 * p.bC_d + a.S2 * p.bS_d);
    Fp = av * g.c2psi + bv * g.s2psi;
    Fc = bv * g.c2psi - av * g.s2psi;
    const double Gr = Fp * g.K, Gi = Fc * g.ci;                                       // signal.py:463-464
    const double Ggr = (ag * g.c2psi + bg * g.s2psi) * g.K, Ggi = (bg * g.c2psi - ag * g.s2psi) * g.ci;
    const double Gdr = (ad * g.c2psi + bd * g.s2psi) * g.K, Gdi = (bd * g.c2psi - ad * g.s2psi) * g.ci;
    // intrinsic rows (the AD rows of the reference, signal.py:1153-1189): general indices 0, 1, 5, 6 [, 7, 8]
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        const int gx = j < 2 ? j : 3 + j;
        double re = fma(w.lnA_d[j], Gr, -Gi * dr.psi_x[j]), im = fma(w.lnA_d[j], Gi, Gr * dr.psi_x[j]);
        if (j < 2) {
            re = fma(Ggr, dr.ang_x[j], re);
            im = fma(Ggi, dr.ang_x[j], im);
        }
        ra[gx] = re;
        rb[gx] = im;
    }
    ra[2] = fma(Ggr, dr.ang_t, -Gdr) - Gi * dr.ph_t;          // theta (dec = pi/2 - theta), signal.py:1482-1523
    rb[2] = fma(Ggi, dr.ang_t, -Gdi) + Gr * dr.ph_t;
    ra[3] = fma(Ggr, dr.ang_p, -Gi * dr.ph_p);                // phi, signal.py:1439-1480
    rb[3] = fma(Ggi, dr.ang_p, Gr * dr.ph_p);
    ra[4] = fma(Ggr, dr.ang_c, -Gi * dr.ph_c);                // tcoal, signal.py:1525-1565
    rb[4] = fma(Ggi, dr.ang_c, Gr * dr.ph_c);
}

// compact weighted Gram of one arm-sample: 4 Re int conj(d_a h) d_b h / Sn df, signal.py:922-931
template <int NT>
GWF_HD void gram_accumulate(double wg, double Fp, double Fc, const double* __restrict__ ra, const double* __restrict__ rb,
                            double* __restrict__ acc) {
    typedef Compact<NT> C;
    constexpr int NG = C::NG;
    const double wu = wg * Fp, wv = wg * Fc;
    acc[C::kUU] = fma(wu, Fp, acc[C::kUU]);
    acc[C::kVV] = fma(wv, Fc, acc[C::kVV]);
    acc[C::kUV] = fma(wu, Fc, acc[C::kUV]);
#pragma unroll
    for (int i = 0; i < NG; ++i) {
        double* c = acc + C::kGG + 4 * i;
        c[0] = fma(wu, ra[i], c[0]);
        c[1] = fma(wv, rb[i], c[1]);
        c[2] = fma(wv, ra[i], c[2]);
        c[3] = fma(wu, rb[i], c[3]);
        const double wa = wg * ra[i], wb = wg * rb[i];
#pragma unroll
        for (int j = 0; j <= i; ++j) acc[tri(i, j)] = fma(wa, ra[j], fma(wb, rb[j], acc[tri(i, j)]));
    }
}

template <int NT, int UNIT = 0>
GWF_HD void arm_rows_accumulate(const PointWf<NT>& w, const DetPoint& p, const DetRows<NT>& dr, const ArmDev& a, const EvGeom& g, double wgt,
                                double* __restrict__ acc) {
    double ra[Compact<NT>::NG], rb[Compact<NT>::NG], Fp, Fc;
    arm_rows<NT, UNIT>(w, p, dr, a, g, ra, rb, Fp, Fc);
    gram_accumulate<NT>(wgt * a.weight, Fp, Fc, ra, rb, acc);
}

}  // namespace gwf
