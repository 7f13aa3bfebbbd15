// Host-side construction of the device descriptors (PSD tables, network/arm lists).  Shared by the library
// (gwfast_b200.cu) and by the CPU emulation harness used in tests (tests/emu/emu.cu).
#pragma once
#include <string>
#include <vector>
#include <cstring>
#include <cmath>
#include <algorithm>
#include "../../include/gwfast_b200.h"
#include "detector.cuh"

namespace gwf {

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

// PSD rows (f, S, slope, 0) and the log2-bucket start index used by psd_lookup
static int build_psd_tables(const double* f, const double* S, int n, std::vector<double4>& tab, std::vector<int>& bucket, PsdDev& dev) {
    if (!f || !S || n < 2) return fail(GWF_ERR_ARG, "gwf_psd_create: need at least two (f, S) rows");
    for (int i = 1; i < n; ++i)
        if (!(f[i] > f[i - 1])) return fail(GWF_ERR_ARG, "gwf_psd_create: frequencies must be strictly increasing");
    if (!(f[0] > 0.0)) return fail(GWF_ERR_ARG, "gwf_psd_create: frequencies must be positive");
    tab.resize(n);
    for (int i = 0; i < n; ++i) {
        // slope exactly as numpy's interp builds it: (fp[j+1]-fp[j])/(xp[j+1]-xp[j])
        const double slope = i + 1 < n ? (S[i + 1] - S[i]) / (f[i + 1] - f[i]) : 0.0;
        tab[i] = make_double4(f[i], S[i], slope, i + 1 < n ? f[i + 1] : f[i]);
    }
    const int nb = std::max(1024, std::min(1 << 17, 4 * n));
    const double lo = std::log2(f[0]), hi = std::log2(f[n - 1]);
    const double inv = nb / (hi - lo);
    bucket.resize(nb);
    int j = 0;
    for (int b = 0; b < nb; ++b) {
        const double edge = std::exp2(lo + b / inv);
        while (j + 1 < n && f[j + 1] <= edge) ++j;
        bucket[b] = std::min(j, n - 2);
    }
    dev.tab = nullptr;
    dev.bucket = nullptr;
    dev.n = n;
    dev.nb = nb;
    dev.lo = lo;
    dev.inv = inv;
    dev.f_first = f[0];
    dev.f_last = f[n - 1];
    // log-uniform nodes: the row index follows from log2 f directly (within +-1, fixed up by the lookup)
    const double dl = (hi - lo) / (n - 1);
    double dev_max = 0.0;
    for (int i = 0; i < n; ++i) dev_max = std::max(dev_max, std::fabs(std::log2(f[i]) - (lo + i * dl)));
    dev.uni = dev_max < 0.25 * dl ? 1 : 0;
    dev.u_lo = lo;
    dev.u_inv = 1.0 / dl;
    dev.c_off = -1;
    dev.c_j0 = dev.c_n = 0;
    return GWF_OK;
}

// A PSD handle keeps the host copy of its tables and uploads them to a device the first time a call runs there, so one
// process can drive several devices with the same handles (SURVEY.md 8(b) threading contract).
constexpr int kMaxDevices = 64;
constexpr int kMaxPeers = 8;      // GPUs of one NVSwitch box
struct PsdHost {
    std::vector<double4> tab_h;
    std::vector<int> bucket_h;
    double4* tab[kMaxDevices] = {};
    int* bucket[kMaxDevices] = {};
    PsdDev dev;      // descriptor with null table pointers; collect_psds fills in the current device's
};

static int model_nt(const gwf_model& m) {
    switch (m.id) {
        case GWF_TAYLORF2: return ((m.flags & GWF_MODEL_TIDAL) ? 6 : 4) + ((m.flags & GWF_MODEL_ECCENTRIC) ? 1 : 0);
        case GWF_IMRPHENOMD: return 4;
        case GWF_IMRPHENOMD_NRTIDALV2: return 6;
        case GWF_IMRPHENOMHM: return 4;
        case GWF_IMRPHENOMNSBH: return 6;
    }
    return -1;
}

// NetworkDev for one pass.  pass < 0: everything summed (triangles as the u/v pair); pass >= 0: only arm `pass`
// of the flattened arm list (L: 1 arm, T: 3 arms).  snr_mode: all physical arms, each to its own output slot.
static int build_network(const gwf_detector* dets, int ndet, const PsdDev* psds, int npsd, int pass, bool snr_mode, NetworkDev& net) {
    if (ndet < 1 || ndet > kMaxDet) return fail(GWF_ERR_ARG, "number of detectors must be in [1, 8]");
    if (npsd < 1 || npsd > kMaxPsd) return fail(GWF_ERR_ARG, "number of PSD tables must be in [1, 8]");
    std::memset(&net, 0, sizeof(net));
    net.ndet = ndet;
    net.npsd = npsd;
    for (int i = 0; i < npsd; ++i) net.psd[i] = psds[i];
    int flat = 0;
    for (int i = 0; i < ndet; ++i) {
        const gwf_detector& d = dets[i];
        if (d.shape != 0 && d.shape != 1) return fail(GWF_ERR_ARG, "Enter valid detector configuration");
        if (d.psd < 0 || d.psd >= npsd) return fail(GWF_ERR_ARG, "detector PSD index out of range");
        if (!(d.fmin > 0.0)) return fail(GWF_ERR_ARG, "fmin must be positive");
        DetDev& o = net.det[i];
        o.sl = std::sin(d.lat_rad); o.cl = std::cos(d.lat_rad);
        o.s2l = std::sin(2.0 * d.lat_rad); o.c2l = std::cos(2.0 * d.lat_rad);
        o.slon = std::sin(d.long_rad); o.clon = std::cos(d.long_rad);
        o.fmin = d.fmin; o.fmax = d.fmax > 0.0 ? d.fmax : 0.0;
        o.psd = d.psd;
        o.no_motion = d.no_motion != 0;
        o.use_rot = (d.use_earth_motion != 0) && !o.no_motion;   // signal.py:138-140
        int g = -1;
        for (int k = 0; k < net.ngroups; ++k)
            if (net.group_fmin[k] == o.fmin && net.group_fmax[k] == o.fmax) g = k;
        if (g < 0) {
            if (net.ngroups == kMaxGroups) return fail(GWF_ERR_ARG, "more than 4 distinct (fmin, fmax) pairs in one network");
            g = net.ngroups++;
            net.group_fmin[g] = o.fmin;
            net.group_fmax[g] = o.fmax;
        }
        o.group = g;
        if (o.use_rot) net.group_rot[g] = 1;
        o.arm_begin = net.narms;
        const double sarm = d.shape == 0 ? 1.0 : std::sin(kPi / 3.);          // sin(angbtwArms), signal.py:144-147
        const double s1 = sarm * std::sin(2.0 * d.xax_rad), c1 = sarm * std::cos(2.0 * d.xax_rad);
        const double x2 = d.xax_rad + 60. * kPi / 180.;                       // rot = 60 deg, signal.py:749, 1023
        const double s2 = sarm * std::sin(2.0 * x2), c2 = sarm * std::cos(2.0 * x2);
        auto push = [&](double S, double C, double w, int out) {
            ArmDev& a = net.arm[net.narms++];
            a.S2 = S; a.C2 = C; a.weight = w; a.out = out;
        };
        const int narm_d = d.shape == 0 ? 1 : 3;
        if (net.narms + narm_d > kMaxArms) return fail(GWF_ERR_ARG, "too many arms in one network");
        if (d.shape == 0) {
            if (snr_mode || pass < 0 || pass == flat) push(s1, c1, 1.0, snr_mode ? flat : 0);
        } else if (snr_mode) {
            push(s1, c1, 1.0, flat);
            push(s2, c2, 1.0, flat + 1);
            push(-(s1 + s2), -(c1 + c2), 1.0, flat + 2);                      // signal.py:751
        } else if (pass < 0) {
            // The rows of an arm are linear in its pair (S2, C2) = sin(60 deg) (sin, cos)(2 (xax + rot)), and the three arms sit at
            // 2 xax + {0, 120, 240} deg: sum S2^2 = sum C2^2 = 3/2 sin^2(60 deg) = 9/8 and sum S2 C2 = 0, so
            //     G(r1) + G(r2) + G(-(r1 + r2)) = 9/8 [ G(r_S) + G(r_C) ]
            // with the unit arms r_S: (S2, C2) = (1, 0) and r_C: (0, 1) -- two Grams instead of three, whatever xax is, and rows that
            // need no linear combination at all (detector.cuh:arm_rows<NT, UNIT>; the shape-specialised kernels rely on a two-arm
            // detector being exactly this pair, in this order)
            (void)s1; (void)c1; (void)s2; (void)c2;
            push(1.0, 0.0, 1.125, 0);
            push(0.0, 1.0, 1.125, 0);
        } else {
            if (pass == flat) push(s1, c1, 1.0, 0);
            if (pass == flat + 1) push(s2, c2, 1.0, 0);
            if (pass == flat + 2) push(-(s1 + s2), -(c1 + c2), 1.0, 0);       // signal.py:1057
        }
        flat += narm_d;
        o.arm_end = net.narms;
    }
    return GWF_OK;
}


}  // namespace gwf
