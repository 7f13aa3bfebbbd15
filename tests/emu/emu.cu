// CPU emulation harness for the device math (TEST INFRASTRUCTURE, never shipped or loaded by gwfast_b200).
//
// The per-event prologue and the per-(event, frequency) point functions of gwfast_b200/csrc are written as
// __host__ __device__ code; this file drives exactly those functions in a plain double loop on the host so that
// `pytest -m "not gpu"` can check the formulas against the oracle in a container that has no GPU.  It does not
// exercise the kernels' thread mapping, shared-memory staging or warp reductions -- the `-m gpu` tests do.
#include <vector>
#include <cstring>
#include "../../gwfast_b200/csrc/fisher_core.cuh"
#include "../../gwfast_b200/csrc/host_build.h"
#include "../../gwfast_b200/csrc/covariance.cuh"

using namespace gwf;

struct EmuPsd {
    std::vector<double4> tab;
    std::vector<int> bucket;
    PsdDev dev;
};

static QnmTables g_q = {nullptr, nullptr, nullptr, 0, nullptr};
static std::vector<double> g_qbuf;

template <int MODEL, int NT, bool SD = false>
static int emu_run(const gwf_model* model, const gwf_detector* dets, int ndet, const PsdDev* pd, int npsd, const double* const* ev, long long n,
                   const gwf_opts* opts, double* fisher, double* snr2, double* sd = nullptr) {
    typedef typename ModelTraits<MODEL, NT>::Rec Rec;
    constexpr int NP = NT + 7, NPACK = NP * (NP + 1) / 2;
    ModelCfg cfg = {model->id, model->flags, model->fcutPar, model->fRef};
    const int npass = opts->per_arm ? gwf_num_arms(dets, ndet) : 1;
    for (int pass = 0; pass < npass; ++pass) {
        NetworkDev net;
        int rc = build_network(dets, ndet, pd, npsd, opts->per_arm ? pass : -1, false, net);
        if (rc) return rc;
        for (long long e = 0; e < n; ++e) {
            EventIn in;
            in.Mc = ev[0][e]; in.eta = ev[1][e]; in.dL = ev[2][e]; in.theta = ev[3][e]; in.phi = ev[4][e]; in.iota = ev[5][e];
            in.psi = ev[6][e]; in.tcoal = ev[7][e]; in.Phicoal = ev[8][e]; in.chi1z = ev[9][e]; in.chi2z = ev[10][e];
            in.Lambda1 = ev[11] ? ev[11][e] : 0.; in.Lambda2 = ev[12] ? ev[12][e] : 0.;
            in.fcut_host = ev[13] ? ev[13][e] : 0.; in.s_host = ev[14] ? ev[14][e] : 0.; in.ecc = ev[15] ? ev[15][e] : 0.;
            in.fmax_g = net.group_fmax;
            Rec rec;
            ModelTraits<MODEL, NT>::prologue(rec, in, cfg, opts->flags, g_q, net.group_fmin, net.ngroups);
            EvGeom geom;
            geom.set(in);
            EventScratch sc;
            for (int di = 0; di < net.ndet; ++di) scratch_set(sc, net, geom, di);
            typename PointFns<MODEL, NT>::Extra ex;
            ex.set(in);
            typedef typename PointFnsSel<MODEL, NT, SD>::type PF;
            double acc[PF::kAcc];
            for (int p = 0; p < PF::kAcc; ++p) acc[p] = 0.;
            for (int g = 0; g < net.ngroups; ++g) {
                double fcut = rec.fcut_hz;
                if (net.group_fmax[g] > 0.0 && fcut > net.group_fmax[g]) fcut = net.group_fmax[g];
                Grid grid;
                grid.set(net.group_fmin[g], fcut, opts->res, (opts->flags & GWF_OPT_LIN_GRID) != 0, 32);
                // same walk as the kernels: 32 interleaved strided walks ("lanes")
                for (int lane = 0; lane < 32 && lane < opts->res; ++lane) {
                    FreqPoint fp;
                    grid.start(lane, fp);
                    for (int k = lane; k < opts->res; k += 32) {
                        if (k != lane) grid.advance(k, fp);
                        PF::fisher(rec, cfg, geom, net, sc, ex, g, net.group_rot[g] != 0, fp, acc);
                    }
                }
            }
            double* o = fisher + ((size_t)pass * n + e) * NPACK;
            for (int i = 0; i < NP; ++i)
                for (int j = 0; j <= i; ++j) o[tri(i, j)] = PF::entry(i, j, acc, geom);
            if (snr2) snr2[(size_t)pass * n + e] = PF::snr2(acc, geom);
            if (sd)
                for (int i = 0; i < NP; ++i) sd[((size_t)pass * n + e) * NP + i] = PF::snr_deriv(i, acc, geom);
        }
    }
    return 0;
}

template <int MODEL>
static int emu_run_snr(const gwf_model* model, const gwf_detector* dets, int ndet, const PsdDev* pd, int npsd, const double* const* ev, long long n,
                       const gwf_opts* opts, double* snr2_arm) {
    typedef typename ModelTraits<MODEL, 4>::Rec Rec;
    ModelCfg cfg = {model->id, model->flags, model->fcutPar, model->fRef};
    NetworkDev net;
    int rc = build_network(dets, ndet, pd, npsd, -1, true, net);
    if (rc) return rc;
    for (long long e = 0; e < n; ++e) {
        EventIn in;
        in.Mc = ev[0][e]; in.eta = ev[1][e]; in.dL = ev[2][e]; in.theta = ev[3][e]; in.phi = ev[4][e]; in.iota = ev[5][e];
        in.psi = ev[6][e]; in.tcoal = ev[7][e]; in.Phicoal = ev[8][e]; in.chi1z = ev[9][e]; in.chi2z = ev[10][e];
        in.Lambda1 = ev[11] ? ev[11][e] : 0.; in.Lambda2 = ev[12] ? ev[12][e] : 0.;
            in.fcut_host = ev[13] ? ev[13][e] : 0.; in.s_host = ev[14] ? ev[14][e] : 0.; in.ecc = ev[15] ? ev[15][e] : 0.;
        in.fmax_g = net.group_fmax;
        Rec rec;
        ModelTraits<MODEL, 4>::prologue(rec, in, cfg, 0, g_q, net.group_fmin, net.ngroups);
        EvGeom geom;
        geom.set(in);
        EventScratch sc;
        for (int di = 0; di < net.ndet; ++di) scratch_set(sc, net, geom, di);
        typename PointFns<MODEL, 4>::Extra ex;
        ex.set(in);
        double s2[kMaxArms] = {0};
        for (int g = 0; g < net.ngroups; ++g) {
            double fcut = rec.fcut_hz;
            if (net.group_fmax[g] > 0.0 && fcut > net.group_fmax[g]) fcut = net.group_fmax[g];
            Grid grid;
            grid.set(net.group_fmin[g], fcut, opts->res, (opts->flags & GWF_OPT_LIN_GRID) != 0, 32);
            for (int lane = 0; lane < 32 && lane < opts->res; ++lane) {
                FreqPoint fp;
                grid.start(lane, fp);
                for (int k = lane; k < opts->res; k += 32) {
                    if (k != lane) grid.advance(k, fp);
                    PointFns<MODEL, 4>::snr(rec, cfg, geom, net, sc, ex, g, net.group_rot[g] != 0, fp, s2);
                }
            }
        }
        for (int a = 0; a < net.narms; ++a) snr2_arm[(size_t)a * n + e] = s2[a];
    }
    return 0;
}

template <int MODEL>
static int emu_run_waveform(const gwf_model* model, const double* const* ev, long long n, const double* f, int res, int f2d, double* phi, double* ampl,
                            double* tau, double* hphc, double* fcut) {
    typedef typename ModelTraits<MODEL, 4>::Rec Rec;
    typedef WaveformFns<MODEL> WF;
    ModelCfg cfg = {model->id, model->flags, model->fcutPar, model->fRef};
    const long long plane = (long long)res * n;
    for (long long e = 0; e < n; ++e) {
        EventIn in;
        in.Mc = ev[0][e]; in.eta = ev[1][e]; in.dL = ev[2][e]; in.theta = ev[3][e]; in.phi = ev[4][e]; in.iota = ev[5][e];
        in.psi = ev[6][e]; in.tcoal = ev[7][e]; in.Phicoal = ev[8][e]; in.chi1z = ev[9][e]; in.chi2z = ev[10][e];
        in.Lambda1 = ev[11] ? ev[11][e] : 0.; in.Lambda2 = ev[12] ? ev[12][e] : 0.;
            in.fcut_host = ev[13] ? ev[13][e] : 0.; in.s_host = ev[14] ? ev[14][e] : 0.; in.ecc = ev[15] ? ev[15][e] : 0.;
        double fm = 1.0, fM = 0.0;
        if (res > 0) {
            fm = fM = f2d ? f[e] : f[0];
            for (int k = 1; k < res; ++k) {
                fm = std::min(fm, f2d ? f[(long long)k * n + e] : f[k]);
                fM = std::max(fM, f2d ? f[(long long)k * n + e] : f[k]);
            }
        }
        double fmin_g[kMaxGroups] = {fm, fm, fm, fm}, fmax_g[kMaxGroups] = {fM, fM, fM, fM};
        in.fmax_g = fmax_g;
        in.fmax_exact = true;
        Rec rec;
        ModelTraits<MODEL, 4>::prologue(rec, in, cfg, 0, g_q, fmin_g, 1);
        HMWeights w;
        w.set(in.iota);
        if (fcut) fcut[e] = rec.fcut_hz;
        for (int k = 0; k < res; ++k) {
            FreqPoint fp;
            fp.from_f(f2d ? f[(long long)k * n + e] : f[k]);
            fp.w = 0.;
            WaveformOut o;
            WF::eval(rec, cfg, w, fp, o);
            const long long at = (long long)k * n + e;
            for (int m = 0; m < WF::kModes; ++m) {
                if (phi) phi[m * plane + at] = o.phi[m];
                if (ampl) ampl[m * plane + at] = o.amp[m];
            }
            if (tau) tau[at] = o.tau;
            if (hphc && MODEL == kPhenomHM) { hphc[at] = o.hp[0]; hphc[plane + at] = o.hp[1]; hphc[2 * plane + at] = o.hc[0]; hphc[3 * plane + at] = o.hc[1]; }
        }
    }
    return 0;
}

extern "C" {

// fisherTools row: the double-double covariance / eigen routines of csrc/covariance.cuh driven on the host
int emu_covariance(const double* F, long long n, int nP, int method, double thresh, double* cov, double* inv_err, int* status) {
    std::vector<dd> A(kCovMaxP * kCovMaxP), L(kCovMaxP * kCovMaxP), V(kCovMaxP * kCovMaxP);
    for (long long e = 0; e < n; ++e) status[e] = cov_one(F + e, n, nP, method, thresh, cov + e, inv_err + e, A.data(), L.data(), V.data());
    return 0;
}
int emu_eigen(const double* F, long long n, int nP, double* evals, double* evecs, double* cond) {
    std::vector<dd> A(kCovMaxP * kCovMaxP), V(kCovMaxP * kCovMaxP);
    for (long long e = 0; e < n; ++e) eig_one(F + e, n, nP, evals + e, evecs ? evecs + e : nullptr, cond + e, A.data(), V.data());
    return 0;
}

int emu_psd_lookup(const double* psd_f, const double* psd_S, int n, const double* f, int nf, double* out) {
    EmuPsd P;
    int rc = build_psd_tables(psd_f, psd_S, n, P.tab, P.bucket, P.dev);
    if (rc) return rc;
    P.dev.tab = P.tab.data();
    P.dev.bucket = P.bucket.data();
    for (int i = 0; i < nf; ++i) out[i] = psd_lookup(P.dev, f[i], std::log(f[i]) * 1.4426950408889634073599246810018921);
    return 0;
}

int emu_waveform(const gwf_model* model, const double* const* ev, long long n, const double* f, int res, int f2d, double* phi, double* ampl, double* tau,
                 double* hphc, double* fcut) {
    switch (model->id) {
        case GWF_TAYLORF2: return emu_run_waveform<kTaylorF2>(model, ev, n, f, res, f2d, phi, ampl, tau, hphc, fcut);
        case GWF_IMRPHENOMD: return emu_run_waveform<kPhenomD>(model, ev, n, f, res, f2d, phi, ampl, tau, hphc, fcut);
        case GWF_IMRPHENOMD_NRTIDALV2: return emu_run_waveform<kNRTidalv2>(model, ev, n, f, res, f2d, phi, ampl, tau, hphc, fcut);
        case GWF_IMRPHENOMHM: return emu_run_waveform<kPhenomHM>(model, ev, n, f, res, f2d, phi, ampl, tau, hphc, fcut);
        case GWF_IMRPHENOMNSBH: return emu_run_waveform<kNSBH>(model, ev, n, f, res, f2d, phi, ampl, tau, hphc, fcut);
    }
    return -2;
}

int gwf_num_arms(const gwf_detector* dets, int32_t ndet) {
    int n = 0;
    for (int i = 0; i < ndet; ++i) n += dets[i].shape == 0 ? 1 : 3;
    return n;
}

const char* emu_last_error(void) { return g_err.c_str(); }

int emu_set_qnm(const double* a, const double* fr, const double* fd, int n) {
    g_qbuf.assign(a, a + n);
    g_qbuf.insert(g_qbuf.end(), fr, fr + n);
    g_qbuf.insert(g_qbuf.end(), fd, fd + n);
    g_q.a = g_qbuf.data(); g_q.fring = g_qbuf.data() + n; g_q.fdamp = g_qbuf.data() + 2 * n; g_q.n = n;
    return 0;
}

// IMRPhenomNSBH: use a given 200^3 xi_tide table (nullptr: every node is evaluated on the spot by xitide_node)
int emu_set_xitide(const double* table) {
    g_q.xitide = table;
    return 0;
}
// one slab of the xi_tide table: the 200 x 200 nodes (q, chi) of compactness index i, as the device kernel computes them
int emu_xitide_slab(int i, double* out) {
    for (int j = 0; j < kXiRes; ++j)
        for (int k = 0; k < kXiRes; ++k)
            out[j * kXiRes + k] = xitide_node(xi_node_coord(i, kXiCompMin, kXiCompMax), xi_node_coord(j, kXiQMin, kXiQMax), xi_node_coord(k, kXiChiMin, kXiChiMax));
    return 0;
}

// debugging aid: the waveform point (A, d ln A, d Phi; NT = 6 tidal parametrisation) of IMRPhenomNSBH on the samples f[0..res) of one event
int emu_nsbh_point(const gwf_model* model, const double* evv, double fmin, double fmax, const double* f, int res, double* A, double* lnA_d, double* phi_d) {
    ModelCfg cfg = {model->id, model->flags, model->fcutPar, model->fRef};
    EventIn in;
    in.Mc = evv[0]; in.eta = evv[1]; in.dL = evv[2]; in.theta = evv[3]; in.phi = evv[4]; in.iota = evv[5]; in.psi = evv[6]; in.tcoal = evv[7];
    in.Phicoal = evv[8]; in.chi1z = evv[9]; in.chi2z = evv[10]; in.Lambda1 = evv[11]; in.Lambda2 = evv[12]; in.fcut_host = 0; in.s_host = 0; in.ecc = 0;
    double fmin_g[kMaxGroups] = {fmin, fmin, fmin, fmin}, fmax_g[kMaxGroups] = {fmax, fmax, fmax, fmax};
    in.fmax_g = fmax_g;
    NSBHRec<6> rec;
    ModelTraits<kNSBH, 6>::prologue(rec, in, cfg, 0, g_q, fmin_g, 1);
    for (int k = 0; k < res; ++k) {
        FreqPoint fp;
        fp.from_f(f[k]);
        fp.w = 0.;
        PointWf<6> w;
        ModelTraits<kNSBH, 6>::eval(rec, cfg, 0, fp, false, w);
        A[k] = w.A;
        for (int j = 0; j < 6; ++j) { lnA_d[k * 6 + j] = w.lnA_d[j]; phi_d[k * 6 + j] = w.phi_d[j]; }
    }
    return 0;
}

// psd_f/psd_S: concatenated host tables, psd_n[i] rows each
int emu_fisher(const gwf_model* model, const gwf_detector* dets, int ndet, const double* const* psd_f, const double* const* psd_S, const int* psd_n,
               int npsd, const double* const* ev, long long n, const gwf_opts* opts, double* fisher, double* snr2, int snr_mode) {
    std::vector<EmuPsd> P(npsd);
    PsdDev pd[kMaxPsd];
    for (int i = 0; i < npsd; ++i) {
        int rc = build_psd_tables(psd_f[i], psd_S[i], psd_n[i], P[i].tab, P[i].bucket, P[i].dev);
        if (rc) return rc;
        P[i].dev.tab = P[i].tab.data();
        P[i].dev.bucket = P[i].bucket.data();
        pd[i] = P[i].dev;
    }
    if (snr_mode) {
        switch (model->id) {
            case GWF_TAYLORF2: return emu_run_snr<kTaylorF2>(model, dets, ndet, pd, npsd, ev, n, opts, fisher);
            case GWF_IMRPHENOMD: return emu_run_snr<kPhenomD>(model, dets, ndet, pd, npsd, ev, n, opts, fisher);
            case GWF_IMRPHENOMD_NRTIDALV2: return emu_run_snr<kNRTidalv2>(model, dets, ndet, pd, npsd, ev, n, opts, fisher);
            case GWF_IMRPHENOMHM: return emu_run_snr<kPhenomHM>(model, dets, ndet, pd, npsd, ev, n, opts, fisher);
            case GWF_IMRPHENOMNSBH: return emu_run_snr<kNSBH>(model, dets, ndet, pd, npsd, ev, n, opts, fisher);
        }
        return -2;
    }
    switch (model->id) {
        case GWF_TAYLORF2:
            if (model->flags & GWF_MODEL_ECCENTRIC) {
                if (model->flags & GWF_MODEL_TIDAL) return emu_run<kTaylorF2, 7>(model, dets, ndet, pd, npsd, ev, n, opts, fisher, snr2);
                return emu_run<kTaylorF2, 5>(model, dets, ndet, pd, npsd, ev, n, opts, fisher, snr2);
            }
            if (model->flags & GWF_MODEL_TIDAL) return emu_run<kTaylorF2, 6>(model, dets, ndet, pd, npsd, ev, n, opts, fisher, snr2);
            return emu_run<kTaylorF2, 4>(model, dets, ndet, pd, npsd, ev, n, opts, fisher, snr2);
        case GWF_IMRPHENOMD: return emu_run<kPhenomD, 4>(model, dets, ndet, pd, npsd, ev, n, opts, fisher, snr2);
        case GWF_IMRPHENOMD_NRTIDALV2: return emu_run<kNRTidalv2, 6>(model, dets, ndet, pd, npsd, ev, n, opts, fisher, snr2);
        case GWF_IMRPHENOMHM: return emu_run<kPhenomHM, 4>(model, dets, ndet, pd, npsd, ev, n, opts, fisher, snr2);
        case GWF_IMRPHENOMNSBH: return emu_run<kNSBH, 6>(model, dets, ndet, pd, npsd, ev, n, opts, fisher, snr2);
    }
    return -2;
}

// Fisher + (h | d_i h) (gwf_fisher_ex)
int emu_fisher_sd(const gwf_model* model, const gwf_detector* dets, int ndet, const double* const* psd_f, const double* const* psd_S, const int* psd_n,
                  int npsd, const double* const* ev, long long n, const gwf_opts* opts, double* fisher, double* snr2, double* sd) {
    std::vector<EmuPsd> P(npsd);
    PsdDev pd[kMaxPsd];
    for (int i = 0; i < npsd; ++i) {
        int rc = build_psd_tables(psd_f[i], psd_S[i], psd_n[i], P[i].tab, P[i].bucket, P[i].dev);
        if (rc) return rc;
        P[i].dev.tab = P[i].tab.data();
        P[i].dev.bucket = P[i].bucket.data();
        pd[i] = P[i].dev;
    }
    switch (model->id) {
        case GWF_TAYLORF2:
            if (model->flags & GWF_MODEL_ECCENTRIC) {
                if (model->flags & GWF_MODEL_TIDAL) return emu_run<kTaylorF2, 7>(model, dets, ndet, pd, npsd, ev, n, opts, fisher, snr2, sd);
                return emu_run<kTaylorF2, 5>(model, dets, ndet, pd, npsd, ev, n, opts, fisher, snr2, sd);
            }
            if (model->flags & GWF_MODEL_TIDAL) return emu_run<kTaylorF2, 6>(model, dets, ndet, pd, npsd, ev, n, opts, fisher, snr2, sd);
            return emu_run<kTaylorF2, 4>(model, dets, ndet, pd, npsd, ev, n, opts, fisher, snr2, sd);
        case GWF_IMRPHENOMD: return emu_run<kPhenomD, 4>(model, dets, ndet, pd, npsd, ev, n, opts, fisher, snr2, sd);
        case GWF_IMRPHENOMD_NRTIDALV2: return emu_run<kNRTidalv2, 6>(model, dets, ndet, pd, npsd, ev, n, opts, fisher, snr2, sd);
        case GWF_IMRPHENOMHM: return emu_run<kPhenomHM, 4, true>(model, dets, ndet, pd, npsd, ev, n, opts, fisher, snr2, sd);
        case GWF_IMRPHENOMNSBH: return emu_run<kNSBH, 6>(model, dets, ndet, pd, npsd, ev, n, opts, fisher, snr2, sd);
    }
    return -2;
}
}
