// IMRPhenomD (gwfast/waveforms.py:959-1333) split B200-style into
//   * a per-event prologue in dual arithmetic that produces a COEFFICIENT RECORD: every region of the phase
//     and of the amplitude becomes  sum_k c_k * b_k(x)  with dual coefficients c_k (value + NT tangents) and a
//     few closed-form non-linear terms (the MRD arctan and the MRD Lorentzian), x = M GMsun/c^3 f;
//   * a per-frequency evaluation that needs no dual arithmetic at all: with lam_j = d ln(x)/d p_j,
//        d g / d p_j = sum_k dc_kj b_k(x) + lam_j * sum_k c_k (x b_k'(x)).
// The same record layout serves IMRPhenomD_NRTidalv2 (model_nrtidal.cuh adds its tidal terms).
#pragma once
#include "model_common.cuh"

namespace gwf {

constexpr double kPhiJoinIns = 0.018;   // waveforms.py:980
constexpr double kAmpJoinIns = 0.014;   // waveforms.py:978
constexpr double kMfCut = 0.2;          // waveforms.py:982

// arXiv:1508.07253 Tab. 5 as typed at waveforms.py:1039-1056, 1189-1197, 1223; row layout:
// t0 + t1 eta + xi (t2 + t3 eta + t4 eta^2) + xi^2 (t5 + t6 eta + t7 eta^2) + xi^3 (t8 + t9 eta + t10 eta^2)
enum FitId { SIG1, SIG2, SIG3, SIG4, BET1, BET2, BET3, ALP1, ALP2, ALP3, ALP4, ALP5, GAM1, GAM2, GAM3, RHO1, RHO2, RHO3, V2FIT, kNumFits };
#define GWF_PHENOMD_FIT_ROWS \
    {2096.551999295543, 1463.7493168261553, 1312.5493286098522, 18307.330017082117, -43534.1440746107, -833.2889543511114, 32047.31997183187, -108609.45037520859, 452.25136398112204, 8353.439546391714, -44531.3250037322}, \
    {-10114.056472621156, -44631.01109458185, -6541.308761668722, -266959.23419307504, 686328.3229317984, 3405.6372187679685, -437507.7208209015, 1631817.1307344697, -7462.648563007646, -114585.25177153319, 674402.4689098676}, \
    {22933.658273436497, 230960.00814979506, 14961.083974183695, 1194018.1342318142, -3104223.9693052764, -3038.166617199259, 1872032.2849093592, -7309145.012085539, 42738.22871475411, 467502.018616601, -3064853.498512499}, \
    {-14621.71522218357, -377812.8579387104, -9608.682631509726, -1710892.5257214056, 4332924.601416521, -22366.683262266528, -2501971.6386377467, 10274495.902259542, -85360.30079034246, -570025.3441737515, 4396844.346849777}, \
    {97.89747327985583, -42.659730877489224, 153.48421037904913, -1417.0620760768954, 2752.8614143665027, 138.7406469558649, -1433.6585075135881, 2857.7418952430758, 41.025109467376126, -423.680737974639, 850.3594335657173}, \
    {-3.282701958759534, -9.051384468245866, -12.415449742258042, 55.4716447709787, -106.05109938966335, -11.953044553690658, 76.80704618365418, -155.33172948098394, -3.4129261592393263, 25.572377569952536, -54.408036707740465}, \
    {-2.5156429818799565e-05, 1.9750256942201327e-05, -1.8370671469295915e-05, 2.1886317041311973e-05, 8.250240316860033e-05, 7.157371250566708e-06, -5.5780000112270685e-05, 0.00019142082884072178, 5.447166261464217e-06, -3.220610095021982e-05, 7.974016714984341e-05}, \
    {43.31514709695348, 638.6332679188081, -32.85768747216059, 2415.8938269370315, -5766.875169379177, -61.85459307173841, 2953.967762459948, -8986.29057591497, -21.571435779762044, 981.2158224673428, -3239.5664895930286}, \
    {-0.07020209449091723, -0.16269798450687084, -0.1872514685185499, 1.138313650449945, -2.8334196304430046, -0.17137955686840617, 1.7197549338119527, -4.539717148261272, -0.049983437357548705, 0.6062072055948309, -1.682769616644546}, \
    {9.5988072383479, -397.05438595557433, 16.202126189517813, -1574.8286986717037, 3600.3410843831093, 27.092429659075467, -1786.482357315139, 5152.919378666511, 11.175710130033895, -577.7999423177481, 1808.730762932043}, \
    {-0.02989487384493607, 1.4022106448583738, -0.07356049468633846, 0.8337006542278661, 0.2240008282397391, -0.055202870001177226, 0.5667186343606578, 0.7186931973380503, -0.015507437354325743, 0.15750322779277187, 0.21076815715176228}, \
    {0.9974408278363099, -0.007884449714907203, -0.059046901195591035, 1.3958712396764088, -4.516631601676276, -0.05585343136869692, 1.7516580039343603, -5.990208965347804, -0.017945336522161195, 0.5965097794825992, -2.0608879367971804}, \
    {0.006927402739328343, 0.03020474290328911, 0.006308024337706171, -0.12074130661131138, 0.26271598905781324, 0.0034151773647198794, -0.10779338611188374, 0.27098966966891747, 0.0007374185938559283, -0.02749621038376281, 0.0733150789135702}, \
    {1.010344404799477, 0.0008993122007234548, 0.283949116804459, -4.049752962958005, 13.207828172665366, 0.10396278486805426, -7.025059158961947, 24.784892370130475, 0.03093202475605892, -2.6924023896851663, 9.609374464684983}, \
    {1.3081615607036106, -0.005537729694807678, -0.06782917938621007, -0.6689834970767117, 3.403147966134083, -0.05296577374411866, -0.9923793203111362, 4.820681208409587, -0.006134139870393713, -0.38429253308696365, 1.7561754421985984}, \
    {3931.8979897196696, -17395.758706812805, 3132.375545898835, 343965.86092361377, -1216256.5819981997, -70698.00600428853, 1383907.177859705, -3966276.1890979446, -60017.52423652596, 803515.1181825735, -2091710.365941658}, \
    {-40105.47653771657, 112253.0169706701, 23561.696065836168, -3476180.699403351, 11375936.70849482, 754313.1127166454, -13084760.44625268, 36444584.853928134, 596226.612472288, -7427790.1143564405, 18928977.514040343}, \
    {83208.35471266537, -191237.7264145924, -210916.2454782992, 8717975.08352568, -26914942.420669552, -1988980.6527362722, 30888029.960154563, -83908702.79256162, -1453503.1953446497, 17063528.990822166, -42748659.731120914}, \
    {0.8149838730507785, 2.5747553517454658, 1.1610198035496786, -2.3627771785551537, 6.771038707057573, 0.7570782938606834, -2.7256896890432474, 7.1140380397149965, 0.1766934149293479, -0.7978690983168183, 2.1162391502005153}
__device__ __constant__ double kPhenomDFitsDev[kNumFits][11] = {GWF_PHENOMD_FIT_ROWS};
static const double kPhenomDFitsHost[kNumFits][11] = {GWF_PHENOMD_FIT_ROWS};

template <class T> GWF_HD T phenomd_fit(int id, const T& eta, const T& e2, const T& xi) {
#ifdef __CUDA_ARCH__
    const double* t = kPhenomDFitsDev[id];
#else
    const double* t = kPhenomDFitsHost[id];
#endif
    return t[0] + t[1] * eta + (t[2] + t[3] * eta + t[4] * e2 + (t[5] + t[6] * eta + t[7] * e2) * xi + (t[8] + t[9] * eta + t[10] * e2) * xi * xi) * xi;
}

// ------------------------------------------------------------------------------------------------
// coefficient record.  c[k][0] = value, c[k][1+j] = tangent w.r.t. intrinsic slot j.
constexpr int kPIns = 13;   // 1, x^2/3, x^1/3, x^1/3 L, L, x^-1/3, x^-2/3, x^-1, x^-5/3, x, x^4/3, x^5/3, x^2 ; L = log(pi x)/3
constexpr int kPInt = 5;    // 1, x, x^-3, log x, x^2/3 (the last only for NRTidalv2's 3.5PN SS/SSS term)
constexpr int kPMrd = 5;    // 1, x, 1/x, x^3/4, x^2/3 (idem)
constexpr int kAIns = 9;    // 1, x^2/3, x, x^4/3, x^5/3, x^2, x^7/3, x^8/3, x^3
constexpr int kAInt = 5;    // (x - 0.014)^k, k=0..4
// rows of the expanded tables are padded to an even number of doubles and 16-byte aligned: the per-frequency evaluation
// reads a row (value + NT tangents) with 16-byte shared-memory loads
template <int NT> struct RowLen { static constexpr int v = (NT + 2) & ~1; };
// the amplitude rows: IMRPhenomNSBH reuses the phase half of the record only (its amplitude is a different model) and leaves them out
// (AMP = false) -- 1.1 KB per staged event, the difference between fitting the PSD windows of ET+2CE into shared memory or not
template <int NT, bool AMP> struct PhenomDAmpRows {
    alignas(16) double ains[kAIns][RowLen<NT>::v];
    alignas(16) double aint[kAInt][RowLen<NT>::v];
    double amrd[4][1 + NT];   // fring, gamma2/(fdamp gamma3), fdamp*gamma3, fdamp*gamma3*gamma1
};
template <int NT> struct PhenomDAmpRows<NT, false> {};
template <int NT, bool AMP = true>
struct PhenomDRec : PhenomDAmpRows<NT, AMP> {
    double s;                 // x = s f      (s = M GMsun/c^3)
    ScalePow sp;              // powers of s used to build x^(1/3), x^(-1/3), ln(pi x)/3 from the grid's f-powers
    double lam[NT];           // d ln s / d slot
    double fcut_hz;           // model cut frequency in Hz (before the detector's fmax clip)
    double x_mrd, x_peak;     // phase int->MRD join (fring/2), amplitude int->MRD join (fpeak)
    double C, lnC_d[NT];      // overall amplitude factor 2 sqrt(5/64pi) M^2 GMsun_c2_Gpc GMsun_c3/dL * amp0, and d ln C
    double C76;               // C * s^(-7/6): A = C76 f^(-7/6) ampIMR(x)
    double pc[kMaxGroups][1 + NT];   // per grid group: t0*xRef - phiRef  (xRef = s*fmin unless fRef given)
    alignas(16) double pins[kPIns][RowLen<NT>::v];
    alignas(16) double pint[kPInt][RowLen<NT>::v];
    alignas(16) double pmrd[kPMrd][RowLen<NT>::v];
    double atn[3][1 + NT];    // MRD arctan term: alpha4/eta, alpha5*fring, fdamp
    TauRec tau;
};

template <int NT> GWF_HD void put(double* dst, const Dual<NT>& c) {
    dst[0] = c.v;
#pragma unroll
    for (int j = 0; j < NT; ++j) dst[1 + j] = c.d[j];
}
template <int NT> GWF_HD Dual<NT> get(const double* src) {
    Dual<NT> c; c.v = src[0];
#pragma unroll
    for (int j = 0; j < NT; ++j) c.d[j] = src[1 + j];
    return c;
}

// all f-independent PhenomD quantities as duals (waveforms.py:1002-1136, 1165-1245)
template <int NT>
struct PhenomDCore {
    typedef Dual<NT> D;
    D eta, fring, fdamp, fit[kNumFits], norm;
    PNPhase<D> pn;
    D C1Int, C2Int, C1MRD, C2MRD, fMRDJoin, fpeak_amp, t0;
    D A[10];      // inspiral amplitude coefficients of x^(k/3), k=2..9 (waveforms.py:1206-1213)
    D e[kAInt];   // intermediate amplitude in powers of (x - 0.014)

    GWF_HD D phi_ins(const D& x) const {   // waveforms.py:1081-1094 + 1145
        const double cp = cbrt(kPi);
        const D x13 = dpow(x, 1. / 3.), x23 = x13 * x13, lg = dlog(kPi * x) / 3.;
        const D pnv = pn.c5 * norm + pn.c7 * norm * (cp * cp) * x23 + pn.c6 * norm * cp * x13 + (-6848. / 21.) * norm * cp * x13 * lg +
                      3. * pn.c5 * norm * lg + pn.c4 * norm / cp / x13 + pn.c3 * norm / (cp * cp) / x23 + pn.c2 * norm / kPi / x +
                      norm / (kPi * cp * cp) / (x * x23);
        return pnv + (fit[SIG1] * x + fit[SIG2] * 0.75 * x * x13 + fit[SIG3] * 0.6 * x * x23 + fit[SIG4] * 0.5 * x * x) / eta;
    }
    GWF_HD D dphi_ins(const D& x) const {  // waveforms.py:1108-1109
        const D px = kPi * x, p13 = dpow(px, 1. / 3.), p23 = p13 * p13;
        const D d = (2.0 * pn.c7 * norm * (px * px * p13) + (pn.c6 * norm + (-6848. / 21.) * norm * (1.0 + dlog(px) / 3.)) * (px * px) +
                     3. * pn.c5 * norm * (px * p23) - pn.c4 * norm * (px * p13) - 2. * pn.c3 * norm * px - 3. * pn.c2 * norm * p23 - 5. * norm) *
                    kPi / (3. * (px * px * p23));
        const D x13 = dpow(x, 1. / 3.);
        return d + (fit[SIG1] + fit[SIG2] * x13 + fit[SIG3] * (x13 * x13) + fit[SIG4] * x) / eta;
    }
    GWF_HD D phi_int_raw(const D& x) const { return fit[BET1] * x - fit[BET3] / (3. * x * x * x) + fit[BET2] * dlog(x); }
    GWF_HD D dphi_int(const D& x) const { const D x2 = x * x; return (fit[BET1] + fit[BET3] / (x2 * x2) + fit[BET2] / x) / eta; }
    GWF_HD D phi_mrd_raw(const D& x) const {
        return -(fit[ALP2] / x) + (4.0 / 3.0) * (fit[ALP3] * dpow(x, 0.75)) + fit[ALP1] * x + fit[ALP4] * datan((x - fit[ALP5] * fring) / fdamp);
    }
    GWF_HD D dphi_mrd(const D& x) const {
        const D u = x - fit[ALP5] * fring;
        return (fit[ALP1] + fit[ALP2] / (x * x) + fit[ALP3] / dpow(x, 0.25) + fit[ALP4] / (fdamp * (1. + u * u / (fdamp * fdamp)))) / eta;
    }
    GWF_HD D phi_regions(const D& x, bool apply_cut) const {   // the nested where of waveforms.py:1143-1151
        if (x.v < kPhiJoinIns) return phi_ins(x);
        if (x.v < fMRDJoin.v) return phi_int_raw(x) / eta + C1Int + C2Int * x;
        if (!apply_cut || x.v < kMfCut) return phi_mrd_raw(x) / eta + C1MRD + C2MRD * x;
        return D(0.0);
    }
    GWF_HD D amp_ins(const D& x) const {    // waveforms.py:1215
        const D x13 = dpow(x, 1. / 3.), x23 = x13 * x13, x2 = x * x;
        return 1. + x23 * A[2] + (x * x13) * A[4] + (x * x23) * A[5] + (x2 * x13) * A[7] + (x2 * x23) * A[8] + x * (A[3] + x * A[6] + x2 * A[9]);
    }
    GWF_HD D damp_ins(const D& x) const {   // waveforms.py:1217 (d/dx of the line above)
        const D x13 = dpow(x, 1. / 3.), x23 = x13 * x13;
        return (2. / 3.) * A[2] / x13 + A[3] + (4. / 3.) * A[4] * x13 + (5. / 3.) * A[5] * x23 + 2. * A[6] * x + (7. / 3.) * (x * x13) * A[7] +
               (8. / 3.) * (x * x23) * A[8] + 3. * (x * x) * A[9];
    }
    GWF_HD D amp_mrd(const D& x) const {    // waveforms.py:1219
        const D fd3 = fdamp * fit[GAM3], u = x - fring;
        return dexp(-u * fit[GAM2] / fd3) * (fd3 * fit[GAM1]) / (u * u + fd3 * fd3);
    }
    GWF_HD D damp_mrd(const D& x) const {   // waveforms.py:1221
        const D fd3 = fdamp * fit[GAM3], u = x - fring, den = u * u + fd3 * fd3;
        return ((-2. * fdamp * u * fit[GAM3] * fit[GAM1]) / den - (fit[GAM2] * fit[GAM1])) / (dexp(u * fit[GAM2] / fd3) * den);
    }

    // ringdown/damping frequencies from the QNM tables (IMRPhenomD, NRTidalv2; waveforms.py:1034-1035)
    // parts: 1 = everything the phase needs, 2 = everything the amplitude needs, 3 = both.  The prologue kernel gives the two
    // halves of an IMRPhenomD event to two different warps: the kernel is ~200 KB of straight-line code that each warp runs
    // once, so its duration is the instruction-fetch latency of the code a warp walks through (profiles/r01i), and each half
    // walks through about half of it.  The arithmetic of every quantity is unchanged.
    GWF_HD void build(const D& eta_, const D& chi1, const D& chi2, const D& qm1, const D& qm2, const QnmTables& q, int parts = 3) {
        const D aeff = final_spin(eta_, chi1, chi2), erad = radiated_energy(eta_, chi1, chi2);
        build_rd(eta_, chi1, chi2, qm1, qm2, qnm_interp(q, q.fring, aeff) / (1.0 - erad), qnm_interp(q, q.fdamp, aeff) / (1.0 - erad), parts);
    }
    // ... or supplied by the caller (IMRPhenomHM uses polynomial fits, waveforms.py:2336-2345)
    GWF_HD void build_rd(const D& eta_, const D& chi1, const D& chi2, const D& qm1, const D& qm2, const D& fring_, const D& fdamp_, int parts = 3) {
        eta = eta_;
        fring = fring_;
        fdamp = fdamp_;
        const D e2 = eta * eta, sq = seta_of(eta);
        const D xs = 0.5 * (chi1 + chi2), xa = 0.5 * (chi1 - chi2);
        const D xi = -1.0 + (xs * (1.0 - eta * 76.0 / 113.0) + sq * xa);
        // fits used by the phase: sigma, beta, alpha (+ gamma2, gamma3 for the peak); by the amplitude: gamma, rho, v2
        const unsigned need = ((parts & 1) ? ((1u << (ALP5 + 1)) - 1u) | (1u << GAM2) | (1u << GAM3) : 0u) |
                              ((parts & 2) ? (1u << GAM1) | (1u << GAM2) | (1u << GAM3) | (1u << RHO1) | (1u << RHO2) | (1u << RHO3) | (1u << V2FIT) : 0u);
        for (int k = 0; k < kNumFits; ++k)
            if (need & (1u << k)) fit[k] = phenomd_fit(k, eta, e2, xi);
        fMRDJoin = 0.5 * fring;
        // peak frequency, waveforms.py:1134 (phase, |.| on both branches) and :1193 (amplitude)
        const D g2 = fit[GAM2], g3 = fit[GAM3];
        if (g2.v >= 1.0) fpeak_amp = dfabs(fring - (fdamp * g3) / g2);
        else fpeak_amp = fring + (fdamp * (-1.0 + dsqrt(1.0 - g2 * g2)) * g3) / g2;
        if (parts & 1) {
            pn = pn_phase_coeffs(eta, chi1, chi2, qm1, qm2, false);
            pn.c6 = pn.c6 - pn.ss6;                     // waveforms.py:1077
            norm = 3. / (128. * eta);
            // C(1) joins, waveforms.py:1098-1129
            const D fj(kPhiJoinIns);
            C2Int = dphi_ins(fj) - dphi_int(fj);
            C1Int = phi_ins(fj) - phi_int_raw(fj) / eta - C2Int * fj;
            C2MRD = (C2Int + dphi_int(fMRDJoin)) - dphi_mrd(fMRDJoin);
            C1MRD = (phi_int_raw(fMRDJoin) / eta + C1Int + C2Int * fMRDJoin) - phi_mrd_raw(fMRDJoin) / eta - C2MRD * fMRDJoin;
            t0 = dphi_mrd(dfabs(fpeak_amp));
        }
        if (!(parts & 2)) return;
        // inspiral amplitude, waveforms.py:1204-1213
        const D sp1 = 1.0 + sq, c12 = chi1 * chi1, c22 = chi2 * chi2;
        const double cp = cbrt(kPi), cp2 = cp * cp;
        A[2] = ((-969. + 1804. * eta) * cp2) / 672.;
        A[3] = ((chi1 * (81. * sp1 - 44. * eta) + chi2 * (81. - 81. * sq - 44. * eta)) * kPi) / 48.;
        A[4] = ((-27312085.0 - 10287648. * c22 - 10287648. * c12 * sp1 + 10287648. * c22 * sq +
                 24. * (-1975055. + 857304. * c12 - 994896. * chi1 * chi2 + 857304. * c22) * eta + 35371056. * e2) * (kPi * cp)) / 8.128512e6;
        A[5] = ((kPi * cp2) * (chi2 * (-285197. * (-1. + sq) + 4. * (-91902. + 1579. * sq) * eta - 35632. * e2) +
                               chi1 * (285197. * sp1 - 4. * (91902. + 1579. * sq) * eta - 35632. * e2) + 42840. * (-1.0 + 4. * eta) * kPi)) / 32256.;
        A[6] = -((kPi * kPi) * (-336. * (-3248849057.0 + 2943675504. * c12 - 3339284256. * chi1 * chi2 + 2943675504. * c22) * e2 - 324322727232. * e2 * eta -
                                7. * (-177520268561. + 107414046432. * c22 + 107414046432. * c12 * sp1 - 107414046432. * c22 * sq +
                                      11087290368. * (chi1 + chi2 + chi1 * sq - chi2 * sq) * kPi) +
                                12. * eta * (-545384828789. - 176491177632. * chi1 * chi2 + 202603761360. * c22 + 77616. * c12 * (2610335. + 995766. * sq) -
                                             77287373856. * c22 * sq + 5841690624. * (chi1 + chi2) * kPi + 21384760320. * kPi * kPi))) / 6.0085960704e10;
        A[7] = fit[RHO1]; A[8] = fit[RHO2]; A[9] = fit[RHO3];
        // intermediate amplitude: the quartic through (f1;v1,d1), (f2;v2), (f3;v3,d3) (waveforms.py:1199-1245),
        // built by divided differences on the nodes f1,f1,f2,f3,f3 and re-expanded in powers of (x - f1)
        const D f1(kAmpJoinIns), f3 = fpeak_amp, f2 = f1 + 0.5 * (f3 - f1);
        const D v1 = amp_ins(f1), d1 = damp_ins(f1), v3 = amp_mrd(f3), d3 = damp_mrd(f3), v2 = fit[V2FIT];
        const D h2 = f2 - f1, h3 = f3 - f1, h23 = f3 - f2;
        const D a12 = (v2 - v1) / h2, a23 = (v3 - v2) / h23;
        const D b112 = (a12 - d1) / h2, b123 = (a23 - a12) / h3, b233 = (d3 - a23) / h23;
        const D c1123 = (b123 - b112) / h3, c1233 = (b233 - b123) / h3;
        const D dd = (c1233 - c1123) / h3;
        e[0] = v1; e[1] = d1; e[2] = b112 - h2 * c1123 + h2 * h3 * dd; e[3] = c1123 - (h2 + h3) * dd; e[4] = dd;
    }
};

// fill the record from the core; xref[g] = dimensionless reference frequency of grid group g
template <int NT, bool AMP>
GWF_HD void phenomd_fill(PhenomDRec<NT, AMP>& r, const PhenomDCore<NT>& c, const Dual<NT>& M, const Dual<NT>& dL, const double* fmin_g, int ngroups,
                         const ModelCfg& cfg, double s_host = 0.0, double fcut_host = 0.0, int parts = 3) {
    typedef Dual<NT> D;
    D s = M * kGMsunC3;
    if (s_host > 0.0) s.v = s_host;       // the host's M*GMsun_over_c3: x = s f then rounds like the reference's fgrid
    if constexpr (AMP) if (parts & 2) {
        ScalePow sp;
        sp.set(s.v);
        r.x_peak = c.fpeak_amp.v;
        // waveforms.py:1248, 1204, 1254
        const D amp0 = dsqrt(2.0 * c.eta / 3.0) * pow(kPi, -1. / 6.);
        const D Cc = 2. * sqrt(5. / (64. * kPi)) * M * kGMsunC2Gpc * M * kGMsunC3 / dL * amp0;
        r.C = Cc.v;
        r.C76 = Cc.v * (sp.sm13 * sp.sm13 * sp.sm13 * sqrt(sp.sm13));
#pragma unroll
        for (int j = 0; j < NT; ++j) r.lnC_d[j] = Cc.d[j] / Cc.v;
        // amplitude
        put(r.ains[0], D(1.0));
        put(r.ains[1], c.A[2]); put(r.ains[2], c.A[3]); put(r.ains[3], c.A[4]); put(r.ains[4], c.A[5]);
        put(r.ains[5], c.A[6]); put(r.ains[6], c.A[7]); put(r.ains[7], c.A[8]); put(r.ains[8], c.A[9]);
        for (int k = 0; k < kAInt; ++k) put(r.aint[k], c.e[k]);
        const D fd3 = c.fdamp * c.fit[GAM3];
        put(r.amrd[0], c.fring);
        put(r.amrd[1], c.fit[GAM2] / fd3);
        put(r.amrd[2], fd3);
        put(r.amrd[3], fd3 * c.fit[GAM1]);
        tau_fill(r.tau, s, c.eta);
    }
    if (!(parts & 1)) return;
    r.s = s.v;
    r.sp.set(s.v);
#pragma unroll
    for (int j = 0; j < NT; ++j) r.lam[j] = s.d[j] / s.v;
    r.fcut_hz = fcut_host > 0.0 ? fcut_host : kMfCut / s.v;   // waveforms.py:1333
    r.x_mrd = c.fMRDJoin.v;
    const bool apply_cut = !(cfg.flags & kFlagNoFcut);
    for (int g = 0; g < ngroups; ++g) {
        const D xref = (cfg.flags & kFlagHasFRef) ? s * cfg.fRef : s * fmin_g[g];   // waveforms.py:1139-1141
        put(r.pc[g], c.t0 * xref - c.phi_regions(xref, apply_cut));
    }
    const double cp = cbrt(kPi), cp2 = cp * cp;
    const D n = c.norm, ie = 1.0 / c.eta;
    // inspiral phase, waveforms.py:1081-1094; -t0 folded into the coefficient of x
    put(r.pins[0], c.pn.c5 * n);
    put(r.pins[1], c.pn.c7 * n * cp2);
    put(r.pins[2], c.pn.c6 * n * cp);
    put(r.pins[3], (-6848. / 21.) * n * cp);
    put(r.pins[4], 3. * c.pn.c5 * n);
    put(r.pins[5], c.pn.c4 * n / cp);
    put(r.pins[6], c.pn.c3 * n / cp2);
    put(r.pins[7], c.pn.c2 * n / kPi);
    put(r.pins[8], n / (kPi * cp2));
    put(r.pins[9], c.fit[SIG1] * ie - c.t0);
    put(r.pins[10], c.fit[SIG2] * 0.75 * ie);
    put(r.pins[11], c.fit[SIG3] * 0.6 * ie);
    put(r.pins[12], c.fit[SIG4] * 0.5 * ie);
    // intermediate phase, waveforms.py:1145 second branch
    put(r.pint[0], c.C1Int);
    put(r.pint[1], c.fit[BET1] * ie + c.C2Int - c.t0);
    put(r.pint[2], -c.fit[BET3] * ie / 3.);
    put(r.pint[3], c.fit[BET2] * ie);
    put(r.pint[4], D(0.0));
    // merger-ringdown phase
    put(r.pmrd[0], c.C1MRD);
    put(r.pmrd[1], c.fit[ALP1] * ie + c.C2MRD - c.t0);
    put(r.pmrd[2], -c.fit[ALP2] * ie);
    put(r.pmrd[3], (4.0 / 3.0) * c.fit[ALP3] * ie);
    put(r.pmrd[4], D(0.0));
    put(r.atn[0], c.fit[ALP4] * ie);
    put(r.atn[1], c.fit[ALP5] * c.fring);
    put(r.atn[2], c.fdamp);
}

template <int NT>
GWF_HD void phenomd_prologue(PhenomDRec<NT>& r, const Intrinsic<NT>& p, double dL, const QnmTables& q, const double* fmin_g, int ngroups,
                             const ModelCfg& cfg, double s_host = 0.0, double fcut_host = 0.0, int parts = 3) {
    typedef Dual<NT> D;
    PhenomDCore<NT> c;
    c.build(p.eta, p.chi1, p.chi2, D(1.0), D(1.0), q, parts);
    const D M = p.Mc / dpow(p.eta, 3. / 5.);
    phenomd_fill(r, c, M, D(dL), fmin_g, ngroups, cfg, s_host, fcut_host, parts);
}

// ------------------------------------------------------------------------------------------------
// per-frequency evaluation
template <int K, int NT>
GWF_HD void expand(const double (*c)[RowLen<NT>::v], const double* b, const double* bx, double& v, double* d, double& dx) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double row[RowLen<NT>::v];
        const double2* src = reinterpret_cast<const double2*>(c[k]);
#pragma unroll
        for (int q = 0; q < RowLen<NT>::v / 2; ++q) {
            const double2 t = src[q];
            row[2 * q] = t.x;
            row[2 * q + 1] = t.y;
        }
        v = fma(row[0], b[k], v);
        dx = fma(row[0], bx[k], dx);
#pragma unroll
        for (int j = 0; j < NT; ++j) d[j] = fma(row[1 + j], b[k], d[j]);
    }
}

// powers of x shared by the regions
struct XPow {
    double x, x13, x23, xm13, lpx3;   // x^(1/3), x^(2/3), x^(-1/3), log(pi x)/3
    double fm76;                      // f^(-7/6)
    GWF_HD void set(double s, const ScalePow& sp, const FreqPoint& fp) {
        x = s * fp.f;
        x13 = sp.s13 * fp.f13;
        x23 = x13 * x13;
        xm13 = sp.sm13 * fp.fm13;
        lpx3 = fma(fp.lnf, 1. / 3., sp.lps3);
        fm76 = fp.fm76;
    }
};

// phase tangents (and value) at x; g = grid group.  Returns false beyond the cut (phase/amp identically 0).
template <int NT, bool AMP>
GWF_HD void phenomd_phase(const PhenomDRec<NT, AMP>& r, int g, const XPow& p, bool apply_cut, double& phi, double* phi_d) {
    const double x = p.x;
    double v = 0., dx = 0., d[NT];
#pragma unroll
    for (int j = 0; j < NT; ++j) d[j] = 0.;
    if (x < kPhiJoinIns) {
        const double xm23 = p.xm13 * p.xm13, xm1 = xm23 * p.xm13, xm53 = xm1 * xm23, x43 = x * p.x13, x53 = x * p.x23, x2 = x * x;
        const double L = p.lpx3, x13L = p.x13 * L;
        const double b[kPIns] = {1., p.x23, p.x13, x13L, L, p.xm13, xm23, xm1, xm53, x, x43, x53, x2};
        const double bx[kPIns] = {0., 2. / 3. * p.x23, 1. / 3. * p.x13, 1. / 3. * (x13L + p.x13), 1. / 3., -1. / 3. * p.xm13, -2. / 3. * xm23, -xm1,
                                  -5. / 3. * xm53, x, 4. / 3. * x43, 5. / 3. * x53, 2. * x2};
        expand<kPIns, NT>(r.pins, b, bx, v, d, dx);
    } else if (x < r.x_mrd) {
        const double xm1 = p.xm13 * p.xm13 * p.xm13, xm3 = xm1 * xm1 * xm1, lx = 3.0 * p.lpx3 - 1.1447298858494001741434273513530587;   // ln x = ln(pi x) - ln pi
        const double b[kPInt] = {1., x, xm3, lx, p.x23};
        const double bx[kPInt] = {0., x, -3. * xm3, 1., 2. / 3. * p.x23};
        expand<kPInt, NT>(r.pint, b, bx, v, d, dx);
    } else if (!apply_cut || x < kMfCut) {
        const double xm1 = p.xm13 * p.xm13 * p.xm13, sx = sqrt(x), x34 = sx * sqrt(sx);
        const double b[kPMrd] = {1., x, xm1, x34, p.x23};
        const double bx[kPMrd] = {0., x, -xm1, 0.75 * x34, 2. / 3. * p.x23};
        expand<kPMrd, NT>(r.pmrd, b, bx, v, d, dx);
        // alpha4/eta * atan((x - alpha5 fring)/fdamp)
        const double ifd = rcp_fast(r.atn[2][0]);
        const double u = (x - r.atn[1][0]) * ifd;
        const double at = atan(u), w = r.atn[0][0] * rcp_fast(1.0 + u * u) * ifd;   // c * d(atan)/du * (1/fd)
        v = fma(r.atn[0][0], at, v);
        dx = fma(w, x, dx);
#pragma unroll
        for (int j = 0; j < NT; ++j) d[j] += r.atn[0][1 + j] * at - w * (r.atn[1][1 + j] + u * r.atn[2][1 + j]);
    } else {
        phi = 0.;
#pragma unroll
        for (int j = 0; j < NT; ++j) phi_d[j] = 0.;
        return;
    }
    phi = v + r.pc[g][0];
#pragma unroll
    for (int j = 0; j < NT; ++j) phi_d[j] = d[j] + dx * r.lam[j] + r.pc[g][1 + j];
}

// amplitude shape function ampIMR(x) (the three-region where of waveforms.py:1250) and its TOTAL tangents
// dv[j] = sum_k dc_kj b_k + lam_j x d/dx.  Returns false beyond the cut (ampIMR identically 0).
template <int NT>
GWF_HD bool phenomd_amp_core(const PhenomDRec<NT>& r, const XPow& p, bool apply_cut, double& v_out, double* dv) {
    const double x = p.x;
    double v = 0., dx = 0., d[NT];
#pragma unroll
    for (int j = 0; j < NT; ++j) d[j] = 0.;
    if (x < kAmpJoinIns) {
        const double x43 = x * p.x13, x53 = x * p.x23, x2 = x * x;
        const double b[kAIns] = {1., p.x23, x, x43, x53, x2, x2 * p.x13, x2 * p.x23, x2 * x};
        const double bx[kAIns] = {0., 2. / 3. * b[1], x, 4. / 3. * x43, 5. / 3. * x53, 2. * x2, 7. / 3. * b[6], 8. / 3. * b[7], 3. * b[8]};
        expand<kAIns, NT>(r.ains, b, bx, v, d, dx);
    } else if (x < r.x_peak) {
        const double u = x - kAmpJoinIns, u2 = u * u;
        const double b[kAInt] = {1., u, u2, u2 * u, u2 * u2};
        const double bx[kAInt] = {0., x, 2. * x * u, 3. * x * u2, 4. * x * u2 * u};
        expand<kAInt, NT>(r.aint, b, bx, v, d, dx);
    } else if (!apply_cut || x < kMfCut) {
        // exp(-(x-fr) g) * P / ((x-fr)^2 + w^2);  amrd = {fr, g, w, P}
        const double u = x - r.amrd[0][0], g = r.amrd[1][0], w = r.amrd[2][0], P = r.amrd[3][0];
        const double iden = rcp_fast(u * u + w * w), iP = rcp_fast(P);
        v = exp(-u * g) * P * iden;
        // d ln v = -(du) g - u dg + dP/P - (2 u du + 2 w dw)/den, with du_j = x lam_j - dfr_j
        const double ku = -g - 2. * u * iden;
        dx = v * ku * x;
#pragma unroll
        for (int j = 0; j < NT; ++j)
            d[j] = v * (-ku * r.amrd[0][1 + j] - u * r.amrd[1][1 + j] + r.amrd[3][1 + j] * iP - 2. * w * iden * r.amrd[2][1 + j]);
    } else {
        v_out = 0.;
#pragma unroll
        for (int j = 0; j < NT; ++j) dv[j] = 0.;
        return false;
    }
    v_out = v;
#pragma unroll
    for (int j = 0; j < NT; ++j) dv[j] = fma(dx, r.lam[j], d[j]);
    return true;
}

// value-only ampIMR(x) (SNR path: no tangents needed)
template <int NT>
GWF_HD double phenomd_amp_value(const PhenomDRec<NT>& r, const XPow& p, bool apply_cut) {
    const double x = p.x;
    if (x < kAmpJoinIns) {
        const double x2 = x * x;
        return r.ains[0][0] + r.ains[1][0] * p.x23 + r.ains[2][0] * x + r.ains[3][0] * (x * p.x13) + r.ains[4][0] * (x * p.x23) + r.ains[5][0] * x2 +
               r.ains[6][0] * (x2 * p.x13) + r.ains[7][0] * (x2 * p.x23) + r.ains[8][0] * (x2 * x);
    }
    if (x < r.x_peak) {
        const double u = x - kAmpJoinIns;
        return r.aint[0][0] + u * (r.aint[1][0] + u * (r.aint[2][0] + u * (r.aint[3][0] + u * r.aint[4][0])));
    }
    if (!apply_cut || x < kMfCut) {
        const double u = x - r.amrd[0][0], w = r.amrd[2][0];
        return exp(-u * r.amrd[1][0]) * r.amrd[3][0] * rcp_fast(u * u + w * w);
    }
    return 0.0;
}

// amplitude A = C x^(-7/6) ampIMR(x) and d ln A at x
template <int NT>
GWF_HD void phenomd_amp(const PhenomDRec<NT>& r, const XPow& p, bool apply_cut, double& A, double* lnA_d) {
    double v, dv[NT];
    if (!phenomd_amp_core(r, p, apply_cut, v, dv)) {
        A = 0.;
#pragma unroll
        for (int j = 0; j < NT; ++j) lnA_d[j] = 0.;
        return;
    }
    A = r.C76 * p.fm76 * v;
    const double iv = rcp_fast(v);
#pragma unroll
    for (int j = 0; j < NT; ++j) lnA_d[j] = fma(dv[j], iv, fma(-7. / 6., r.lam[j], r.lnC_d[j]));
}

}  // namespace gwf
