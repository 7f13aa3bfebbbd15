// libgwfast_b200.so -- kernels and C ABI (include/gwfast_b200.h).  sm_100a, FP64 throughout.
//
// Launch structure for one gwf_fisher call:
//   K1  prologue_kernel   one thread per event: dual-number evaluation of every f-independent quantity
//                         -> coefficient record (HBM, ~2 KB/event, read once by K2)
//   K2  fisher_kernel     one warp per event, lanes = frequency samples; per sample: waveform amplitude and
//                         tangents by basis expansion against the record (staged in shared memory), detector
//                         geometry + analytic derivative rows per arm, packed Gram accumulated in FP64 registers;
//                         butterfly reduction over the warp; one packed Fisher (+ SNR^2) written per event.
// The derivative strain never exists in memory (the reference materialises (nP,N,res) complex128 per arm,
// gwfast/signal.py:917-922).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include <cmath>
#include <mutex>
#include <unordered_map>

#include "../../include/gwfast_b200.h"
#include "fisher_core.cuh"
#include "host_build.h"
#include "covariance.cuh"

namespace gwf {

#define GWF_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) return fail(GWF_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

struct EventsDev {
    const double* p[GWF_NPARAM_IN];
};
__device__ __forceinline__ EventIn load_event(const EventsDev& ev, long long e) {
    EventIn in;
    in.Mc = ev.p[0][e]; in.eta = ev.p[1][e]; in.dL = ev.p[2][e]; in.theta = ev.p[3][e]; in.phi = ev.p[4][e];
    in.iota = ev.p[5][e]; in.psi = ev.p[6][e]; in.tcoal = ev.p[7][e]; in.Phicoal = ev.p[8][e];
    in.chi1z = ev.p[9][e]; in.chi2z = ev.p[10][e];
    in.Lambda1 = ev.p[11] ? ev.p[11][e] : 0.0;
    in.Lambda2 = ev.p[12] ? ev.p[12][e] : 0.0;
    in.fcut_host = ev.p[13] ? ev.p[13][e] : 0.0;
    in.s_host = ev.p[14] ? ev.p[14][e] : 0.0;
    in.ecc = ev.p[15] ? ev.p[15][e] : 0.0;
    return in;
}

#ifndef GWF_UNR_TF2
#define GWF_UNR_TF2 1
#endif
#ifndef GWF_UNR_D
#define GWF_UNR_D 1
#endif
#ifndef GWF_UNR_NRT
#define GWF_UNR_NRT 1
#endif
#ifndef GWF_FISHER_THREADS
#define GWF_FISHER_THREADS 256
#endif
#ifndef GWF_FISHER_MINBLOCKS
#define GWF_FISHER_MINBLOCKS 1
#endif
constexpr int kFisherThreads = GWF_FISHER_THREADS;
constexpr int kWarpsPerCta = kFisherThreads / 32;
constexpr int kSnrThreads = 512, kSnrWarps = kSnrThreads / 32;     // SNR kernel: 16 warps, one CTA per SM

struct GroupInfo {
    int n;
    double fmin[kMaxGroups];
    double fmax[kMaxGroups];    // upper end of the group's band in Hz (0 = none); IMRPhenomNSBH needs the last grid sample in its prologue
    int fmax_exact;             // fmax[0] is the last sample of a user grid, not a clip
};
static GroupInfo group_info(const NetworkDev& net) {
    GroupInfo gi;
    gi.n = net.ngroups;
    gi.fmax_exact = 0;
    for (int g = 0; g < kMaxGroups; ++g) {
        gi.fmin[g] = net.group_fmin[g];
        gi.fmax[g] = net.group_fmax[g];
    }
    return gi;
}
static GroupInfo group_info_user_grid() {
    GroupInfo gi;
    gi.n = 1;
    gi.fmax_exact = 0;
    for (int g = 0; g < kMaxGroups; ++g) {
        gi.fmin[g] = 1.0;
        gi.fmax[g] = 0.0;
    }
    return gi;
}
// which device tables a model's prologue reads: 1 = QNM ringdown tables, 2 = the xi_tide table of IMRPhenomNSBH
template <int MODEL> constexpr int model_tables() { return (MODEL == kPhenomD || MODEL == kNRTidalv2) ? 1 : (MODEL == kNSBH ? 3 : 0); }

// ------------------------------------------------------------------------------------------- K1
// what the Fisher launch needs to know to fill EventAux (out == nullptr: not wanted)
struct AuxPlan {
    EventAux* out;
    double fmax[kMaxGroups];
    int res, lin, stride;
};

// per-event status word (gwf_fisher_out.status / gwf_snr_ex): input checks, done by the prologue
__device__ __forceinline__ int input_status(const EventIn& in) {
    const bool finite = isfinite(in.Mc) && isfinite(in.eta) && isfinite(in.dL) && isfinite(in.theta) && isfinite(in.phi) && isfinite(in.iota) &&
                        isfinite(in.psi) && isfinite(in.tcoal) && isfinite(in.Phicoal) && isfinite(in.chi1z) && isfinite(in.chi2z) &&
                        isfinite(in.Lambda1) && isfinite(in.Lambda2) && isfinite(in.ecc);
    const bool domain = in.Mc > 0.0 && in.eta > 0.0 && in.eta <= 0.25 && in.dL > 0.0;
    return (finite ? 0 : GWF_EV_NONFINITE_INPUT) | ((finite && !domain) ? GWF_EV_OUT_OF_DOMAIN : 0);
}

template <int MODEL, int NT>
__global__ void __launch_bounds__(128) prologue_kernel(EventsDev ev, long long n, ModelCfg cfg, int opt_flags, QnmTables q, GroupInfo gi,
                                                      typename ModelTraits<MODEL, NT>::Rec* __restrict__ recs,
                                                      const double* __restrict__ fmin_per_event = nullptr, AuxPlan aux = AuxPlan(),
                                                      int* __restrict__ status = nullptr) {
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= n) return;
    EventIn in = load_event(ev, e);
    if (fmin_per_event) {                                   // stand-alone waveform calls: fRef = min of the user's grid,
        gi.fmin[0] = fmin_per_event[e];
        gi.fmax[0] = fmin_per_event[n + e];                 // and its max (IMRPhenomNSBH's t0, waveforms.py:2994)
        gi.fmax_exact = 1;
    }
    in.fmax_g = gi.fmax;
    in.fmax_exact = gi.fmax_exact != 0;
    // gridDim.y = 2: the blocks with blockIdx.y = 0 / 1 compute the two independent halves of every record (IMRPhenomD: phase /
    // amplitude) -- the kernel's duration is the instruction-fetch latency of the code one warp walks through
    const int parts = gridDim.y == 2 ? 1 + (int)blockIdx.y : 3;
    ModelTraits<MODEL, NT>::prologue(recs[e], in, cfg, opt_flags, q, gi.fmin, gi.n, parts);
    int st = status ? input_status(in) : 0;
    if (aux.out && (parts & 1)) {
        EventAux& a = aux.out[e];
        a.geom.set(in);
        for (int g = 0; g < gi.n; ++g) {
            double fcut = recs[e].fcut_hz;
            if (aux.fmax[g] > 0.0 && fcut > aux.fmax[g]) fcut = aux.fmax[g];   // signal.py:717-718
            a.grid[g].set(gi.fmin[g], fcut, aux.res, aux.lin != 0, aux.stride);
            if (!(fcut > gi.fmin[g])) st |= GWF_EV_EMPTY_GRID;
        }
    }
    if (status && (parts & 1)) status[e] = st;
}

// per-event minimum of a user grid f[res][n] (or the shared f[res])
// (fmin[e] = min, fmin[n + e] = max)
__global__ void grid_min_kernel(const double* __restrict__ f, int res, long long n, int f2d, double* __restrict__ fmin) {
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= n) return;
    double m = f2d ? f[e] : f[0], M = m;
    for (int k = 1; k < res; ++k) {
        const double v = f2d ? f[(long long)k * n + e] : f[k];
        m = v < m ? v : m;
        M = v > M ? v : M;
    }
    fmin[e] = m;
    fmin[n + e] = M;
}

// IMRPhenomNSBH: the 200^3 nodes of the xi_tide table (model_nsbh.cuh), one thread per node
__global__ void __launch_bounds__(128) xitide_table_kernel(double* __restrict__ tab) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= kXiRes * kXiRes * kXiRes) return;
    const int k = idx % kXiRes, j = (idx / kXiRes) % kXiRes, i = idx / (kXiRes * kXiRes);
    tab[idx] = xitide_node(xi_node_coord(i, kXiCompMin, kXiCompMax), xi_node_coord(j, kXiQMin, kXiQMax), xi_node_coord(k, kXiChiMin, kXiChiMax));
}

template <class Rec> struct WaveBlk {
    Rec rec;
    HMWeights w;
};

// WaveFormModel.Phi / Ampl / tau_star / hphc on a user grid: one warp per event, lanes over the grid samples
template <int MODEL>
__global__ void __launch_bounds__(kFisherThreads)
waveform_kernel(const typename ModelTraits<MODEL, 4>::Rec* __restrict__ recs, EventsDev ev, long long n, const double* __restrict__ f, int res, int f2d,
                ModelCfg cfg, double* __restrict__ phi, double* __restrict__ ampl, double* __restrict__ tau, double* __restrict__ hphc,
                double* __restrict__ fcut) {
    typedef typename ModelTraits<MODEL, 4>::Rec Rec;
    typedef WaveformFns<MODEL> WF;
    constexpr int kRecDoubles = (int)(sizeof(Rec) / sizeof(double));
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    typedef WaveBlk<Rec> Blk;
    Blk* mine = reinterpret_cast<Blk*>(smem_raw) + wid;
    const long long nwarps = (long long)gridDim.x * kWarpsPerCta;
    for (long long e = (long long)blockIdx.x * kWarpsPerCta + wid; e < n; e += nwarps) {
        const double* src = reinterpret_cast<const double*>(recs + e);
        double* dst = reinterpret_cast<double*>(&mine->rec);
        __syncwarp();
        for (int i = lane; i < kRecDoubles; i += 32) dst[i] = __ldg(src + i);
        if (lane == 0) mine->w.set(ev.p[5][e]);
        __syncwarp();
        if (lane == 0 && fcut) fcut[e] = mine->rec.fcut_hz;
        const long long plane = (long long)res * n;
        for (int k = lane; k < res; k += 32) {
            FreqPoint fp;
            fp.from_f(f2d ? f[(long long)k * n + e] : f[k]);
            fp.w = 0.;
            WaveformOut o;
            WF::eval(mine->rec, cfg, mine->w, fp, o);
            const long long at = (long long)k * n + e;
            for (int m = 0; m < WF::kModes; ++m) {
                if (phi) phi[m * plane + at] = o.phi[m];
                if (ampl) ampl[m * plane + at] = o.amp[m];
            }
            if (tau) tau[at] = o.tau;
            if (hphc && MODEL == kPhenomHM) {
                hphc[at] = o.hp[0]; hphc[plane + at] = o.hp[1]; hphc[2 * plane + at] = o.hc[0]; hphc[3 * plane + at] = o.hc[1];
            }
        }
    }
}

// GWSignal.GWAmplitudes / GWPhase / GWstrain on a USER grid (signal.py:425-655) for one detector and one arm orientation
// (xax + rot): one warp per event, lanes over the grid samples.  Any output may be null.
struct SignalOut {
    double *Ap, *Ac, *psi, *strain, *Fp, *Fc, *dt;
};
template <int MODEL>
__global__ void __launch_bounds__(kFisherThreads)
signal_grid_kernel(const typename ModelTraits<MODEL, 4>::Rec* __restrict__ recs, EventsDev ev, long long n, const double* __restrict__ f, int res, int f2d,
                   ModelCfg cfg, DetDev det, ArmDev arm, SignalOut out) {
    typedef typename ModelTraits<MODEL, 4>::Rec Rec;
    typedef WaveformFns<MODEL> WF;
    constexpr int kRecDoubles = (int)(sizeof(Rec) / sizeof(double));
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    typedef WaveBlk<Rec> Blk;
    Blk* mine = reinterpret_cast<Blk*>(smem_raw) + wid;
    const long long nwarps = (long long)gridDim.x * kWarpsPerCta;
    for (long long e = (long long)blockIdx.x * kWarpsPerCta + wid; e < n; e += nwarps) {
        const double* src = reinterpret_cast<const double*>(recs + e);
        double* dst = reinterpret_cast<double*>(&mine->rec);
        __syncwarp();
        for (int i = lane; i < kRecDoubles; i += 32) dst[i] = __ldg(src + i);
        const EventIn in = load_event(ev, e);
        if (lane == 0) mine->w.set(in.iota);
        __syncwarp();
        EvGeom geom;
        geom.set(in);
        EvDet ed;
        ed.set(det, geom);
        const bool rot = det.use_rot != 0;
        for (int k = lane; k < res; k += 32) {
            FreqPoint fp;
            fp.from_f(f2d ? f[(long long)k * n + e] : f[k]);
            fp.w = 0.;
            WaveformOut o;
            WF::eval(mine->rec, cfg, mine->w, fp, o);
            // time the response is evaluated at (signal.py:444-453, 564-580), before the Earth-centre -> site delay is added
            double sB = 0., cB = 1.;
            if (!det.no_motion) sincos(2.0 * kPi * (rot ? fma(-o.tau, kInvDay, geom.tcoal) : geom.tcoal), &sB, &cB);
            DetPoint dp;
            det_point(ed, cB, sB, dp);
            double Fp, Fc;
            arm_pattern(dp, arm, geom, Fp, Fc);
            const long long at = (long long)k * n + e;
            const double W2 = 2.0 * kPi * fp.f;
            const double carrier = W2 * (geom.tcoal * 3600. * 24.) - in.Phicoal;                 // signal.py:484
            double Ap, Ac, hr, hi;
            if (MODEL == kPhenomHM) {
                // Ap = |hp| Fp, Ac = |hc| Fc (signal.py:457-460); strain = (hp Fp + hc Fc) e^{i (phiL + carrier)} (signal.py:584-607)
                Ap = sqrt(o.hp[0] * o.hp[0] + o.hp[1] * o.hp[1]) * Fp;
                Ac = sqrt(o.hc[0] * o.hc[0] + o.hc[1] * o.hc[1]) * Fc;
                double sP, cP;
                sincos(fma(W2, dp.dt, carrier), &sP, &cP);
                const double zr = o.hp[0] * Fp + o.hc[0] * Fc, zi = o.hp[1] * Fp + o.hc[1] * Fc;
                hr = zr * cP - zi * sP; hi = zr * sP + zi * cP;
            } else {
                Ap = o.amp[0] * Fp * geom.K;                                                      // signal.py:463-464
                Ac = o.amp[0] * Fc * geom.ci;
                double sP, cP;
                sincos((carrier - o.phi[0]) + W2 * dp.dt, &sP, &cP);                              // Psi + phiL, signal.py:580, 641
                hr = Ap * cP - Ac * sP; hi = Ap * sP + Ac * cP;
            }
            if (out.Ap) out.Ap[at] = Ap;
            if (out.Ac) out.Ac[at] = Ac;
            if (out.psi) out.psi[at] = carrier - o.phi[0];
            if (out.strain) reinterpret_cast<double2*>(out.strain)[at] = make_double2(hr, hi);
            if (out.Fp) out.Fp[at] = Fp;
            if (out.Fc) out.Fc[at] = Fc;
            if (out.dt) out.dt[at] = dp.dt;
        }
    }
}

// GWSignal._PatternFunction / _DeltLoc (signal.py:342-423), element-wise over m points: the pattern functions at EXACTLY the given
// time (no delay added) and the Earth-centre -> site delay at that time
__global__ void __launch_bounds__(256) pattern_kernel(DetDev det, ArmDev arm, const double* __restrict__ theta, const double* __restrict__ phi,
                                                      const double* __restrict__ t, const double* __restrict__ psi, long long m,
                                                      double* __restrict__ Fp, double* __restrict__ Fc, double* __restrict__ dt) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= m) return;
    EventIn in = EventIn();
    in.theta = theta[i]; in.phi = phi[i]; in.psi = psi ? psi[i] : 0.0; in.dL = 1.0;
    EvGeom geom;
    geom.set(in);
    EvDet ed;
    ed.set(det, geom);
    double sB, cB;
    sincos(2.0 * kPi * t[i], &sB, &cB);
    const double c1 = ed.cA * cB + ed.sA * sB, s1 = ed.sA * cB - ed.cA * sB;
    if (dt) dt[i] = -kRc * (ed.kc * c1 + ed.k0);                                                  // signal.py:417-421
    if (Fp || Fc) {
        DetPoint dp;
        det_basis(ed, c1, s1, dp);
        double p, c;
        arm_pattern(dp, arm, geom, p, c);
        if (Fp) Fp[i] = p;
        if (Fc) Fc[i] = c;
    }
}

// ------------------------------------------------------------------------------------------- K2

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Warp reduction of L per-lane accumulators by recursive halving: at every step a lane keeps one half of its
// vector and sends the other half to its partner, so the whole reduction costs ~L shuffles instead of 5 L, and each
// lane ends up owning ceil(L/32) fully reduced entries.  All indexing is compile-time (the vector stays in registers).
__host__ __device__ constexpr int fold_half(int L) { return (L + 1) / 2; }
template <int L>
__device__ __forceinline__ void fold_step(double* v, bool upper, int offset) {
    constexpr int H = fold_half(L);
#pragma unroll
    for (int i = 0; i < H; ++i) {
        const double lo = v[i];
        const double hi = (i + H < L) ? v[i + H] : 0.0;
        const double recv = __shfl_xor_sync(0xffffffffu, upper ? lo : hi, offset);
        v[i] = (upper ? hi : lo) + recv;
    }
}
// after fold_all<L>, entry i (i < kFoldOut<L>) of lane `lane` holds the total of original index fold_index<L>(lane, i)
// (or nothing if that index is >= L)
template <int L> struct Fold {
    static constexpr int L1 = fold_half(L), L2 = fold_half(L1), L3 = fold_half(L2), L4 = fold_half(L3), L5 = fold_half(L4);
    static constexpr int kOut = L5;
    static __device__ __forceinline__ void run(double* v, int lane) {
        fold_step<L>(v, (lane & 16) != 0, 16);
        fold_step<L1>(v, (lane & 8) != 0, 8);
        fold_step<L2>(v, (lane & 4) != 0, 4);
        fold_step<L3>(v, (lane & 2) != 0, 2);
        fold_step<L4>(v, (lane & 1) != 0, 1);
    }
    // original index of final slot i, or -1 if the slot only ever held padding
    static __device__ __forceinline__ int index(int lane, int i) {
        int pos = i;
        bool ok = true;
        pos += (lane & 1) ? L5 : 0;  ok = ok && pos < L4;
        pos += (lane & 2) ? L4 : 0;  ok = ok && pos < L3;
        pos += (lane & 4) ? L3 : 0;  ok = ok && pos < L2;
        pos += (lane & 8) ? L2 : 0;  ok = ok && pos < L1;
        pos += (lane & 16) ? L1 : 0; ok = ok && pos < L;
        return ok ? pos : -1;
    }
};

// ---- PSD windows in shared memory -------------------------------------------------------------------------------
// The three PSD tables of a network are ~240 KB of rows + bucket index: together with the spill slots they do not
// fit L1, and a lookup was two dependent L1/L2 round trips per detector and sample (19 % of the Fisher kernel's stall
// samples in profiles/r01).  Log-uniform tables need no bucket index, and the rows above the lowest grid frequency fit
// the CTA's shared memory as the table's own 32-byte rows (f_j, S_j, slope_j, f_{j+1}): one index, two 16-byte loads.
constexpr size_t kSmemLimit = 227 * 1024;
// host: choose the windows (byte offsets from the dynamic-smem base, behind `base` bytes of staging blocks); returns the
// dynamic shared memory size of the launch.  Tables that are not log-uniform or do not fit stay on the global path.
static size_t plan_psd_cache(NetworkDev& net, size_t base, size_t limit) {
    size_t off = (base + 15) & ~(size_t)15;
    for (int i = 0; i < net.npsd; ++i) net.psd[i].c_off = -1;
    for (int i = 0; i < net.npsd; ++i) {
        PsdDev& p = net.psd[i];
        double fq = 0.0;
        bool used = false;
        for (int d = 0; d < net.ndet; ++d)
            if (net.det[d].psd == i && net.det[d].arm_begin != net.det[d].arm_end) {
                const double f0 = net.group_fmin[net.det[d].group];
                fq = used ? std::min(fq, f0) : f0;
                used = true;
            }
        if (!used || !p.uni || !(fq > 0.0)) continue;
        bool shared = false;
        for (int k = 0; k < i && !shared; ++k)
            if (net.psd[k].tab == p.tab && net.psd[k].c_off >= 0 && net.psd[k].c_j0 <= (int)std::floor((std::log2(fq) - p.u_lo) * p.u_inv) - 2) {
                p.c_off = net.psd[k].c_off; p.c_j0 = net.psd[k].c_j0; p.c_n = net.psd[k].c_n;
                shared = true;
            }
        if (shared) continue;
        int j0 = (int)std::floor((std::log2(fq) - p.u_lo) * p.u_inv) - 2;
        j0 = std::max(0, std::min(j0, p.n - 2));
        const int cn = (p.n - 1) - j0;
        const size_t bytes = (size_t)cn * sizeof(double4);
        if (off + bytes > limit) continue;
        p.c_off = (int)off; p.c_j0 = j0; p.c_n = cn;
        off += bytes;
    }
    return off;
}
// host: fast form of the network (see NetworkDev::fast) -- call after plan_psd_cache.  Returns 0 (generic loop), 1 (fast,
// no detector follows the Earth rotation) or 2 (fast, every active detector does).
static int plan_fast(NetworkDev& net) {
    net.fast = 0;
    net.fnd = 0;
    if (net.ngroups != 1) return 0;
    int nrot = 0;
    for (int d = 0; d < net.ndet; ++d) {
        const DetDev& D = net.det[d];
        if (D.arm_begin == D.arm_end) continue;
        if (net.fnd == kMaxFastDet) return 0;
        const PsdDev& P = net.psd[D.psd];
        if (P.c_off < 0) return 0;
        net.fdet[net.fnd] = D;
        net.fpsd[net.fnd] = P;
        nrot += D.use_rot ? 1 : 0;
        ++net.fnd;
    }
    if (net.fnd == 0 || (nrot != 0 && nrot != net.fnd)) { net.fnd = 0; return 0; }
    net.fast = nrot ? 2 : 1;
    return net.fast;
}
// host: shape code of a network in fast form (fisher_core.cuh: base-4 digits = arms per compact detector), or 0 if it has none
// (a detector with more than 3 arms, or arms that are not consecutive from 0 in net.arm[])
static int fast_shape(const NetworkDev& net) {
    if (!net.fast) return 0;
    int shape = 0, first = 0;
    for (int i = 0; i < net.fnd; ++i) {
        const int na = net.fdet[i].arm_end - net.fdet[i].arm_begin;
        if (na < 1 || na > 3 || net.fdet[i].arm_begin != first) return 0;
        shape |= na << (2 * i);
        first += na;
    }
    return shape;
}
// network shapes with their own kernel instantiation: a single L / T, ET + 2 CE, three and four L (Fisher: a triangle is the
// (u, v) pair of virtual arms; SNR: its three physical arms)
#define GWF_FISHER_SHAPES(X) X(1) X(2) X(22) X(21) X(85)
#define GWF_SNR_SHAPES(X) X(1) X(3) X(23) X(21) X(85)

// device: all threads of the CTA copy the windows; ends with a CTA barrier
__device__ __forceinline__ void psd_cache_fill(const NetworkDev& net, unsigned char* smem) {
    for (int i = 0; i < net.npsd; ++i) {
        const PsdDev& p = net.psd[i];
        if (p.c_off < 0) continue;
        bool dup = false;
        for (int k = 0; k < i; ++k) dup = dup || net.psd[k].c_off == p.c_off;
        if (dup) continue;
        double2* dst = reinterpret_cast<double2*>(smem + p.c_off);
        const double2* rows = reinterpret_cast<const double2*>(p.tab + p.c_j0);
        for (int j = threadIdx.x; j < 2 * p.c_n; j += blockDim.x) dst[j] = __ldg(rows + j);
    }
    __syncthreads();
}

// per-warp shared memory: the event's coefficient record followed by the detector scratch
template <class Rec, class Extra> struct WarpSmem {
    Rec rec;
    EventScratch sc;
    EvGeom geom;       // per-event sky/orientation constants: read by broadcast instead of living in 30 registers
    Extra ex;
    double gridc[8];   // geometric-grid walk constants of the current group: r, r13, rm13, rm76, dln, hw_in, hw_hi, fcut
    Grid grid[kMaxGroups];   // Fisher kernel: the event's grids as the prologue left them (EventAux)
};

template <bool FAST, class Rec, class Extra>
__device__ __forceinline__ void stage_event(WarpSmem<Rec, Extra>* mine, const Rec* recs, long long e, const NetworkDev& net, const EvGeom& geom,
                                            const EventIn& in, int lane) {
    constexpr int kRecDoubles = (int)(sizeof(Rec) / sizeof(double));
    const double* src = reinterpret_cast<const double*>(recs + e);
    double* dst = reinterpret_cast<double*>(&mine->rec);
    __syncwarp();
    // coefficient record: coalesced 8-byte loads, later read by broadcast
    for (int i = lane; i < kRecDoubles; i += 32) dst[i] = __ldg(src + i);
    if (FAST) {
        if (lane < net.fnd) scratch_set_fast(mine->sc, net, geom, lane);
    } else if (lane < net.ndet) scratch_set(mine->sc, net, geom, lane);
    if (lane == 31) mine->ex.set(in);
    if (lane == 30) mine->geom = geom;
    __syncwarp();
}

// Fisher kernel: WarpSmem plus the table form of the entry rebuild
constexpr int kXferDoubles = 128;     // packed entries (<= 108) + the two SNR sums + (h | d_i h) (<= 14)
template <class Rec, class Extra> struct FisherSmem : WarpSmem<Rec, Extra> {
    double coef[64];                  // per-event coefficient vector (kCompactCoefs used)
    unsigned char code[4 * 108];      // (ia, ib, ka, kb) per packed entry, filled once per kernel
};
// pair mode: behind the per-warp blocks, two hand-over slots per pair (the second warp's entries on their way to the first, used alternately)
constexpr size_t kXferBytes = sizeof(double) * (kFisherThreads / 64) * 2 * kXferDoubles;

// Fisher kernel: the record and the prologue's EventAux are copied (coalesced), the detector scratch is then set from the
// staged geometry by one lane per detector
template <bool FAST, class Rec, class Extra>
__device__ __forceinline__ void stage_event_aux(WarpSmem<Rec, Extra>* mine, const Rec* recs, const EventAux* aux, const EventsDev& ev, long long e,
                                                const NetworkDev& net, int lane) {
    constexpr int kRecDoubles = (int)(sizeof(Rec) / sizeof(double));
    constexpr int kGeomDoubles = (int)(sizeof(EvGeom) / sizeof(double)), kGridDoubles = (int)(sizeof(Grid) / sizeof(double));
    const double* src = reinterpret_cast<const double*>(recs + e);
    const double* asrc = reinterpret_cast<const double*>(aux + e);
    double* dst = reinterpret_cast<double*>(&mine->rec);
    double* gdst = reinterpret_cast<double*>(&mine->geom);
    double* qdst = reinterpret_cast<double*>(&mine->grid[0]);
    __syncwarp();
    for (int i = lane; i < kRecDoubles; i += 32) dst[i] = __ldg(src + i);
    if (lane < kGeomDoubles) gdst[lane] = __ldg(asrc + lane);
    for (int i = lane; i < kGridDoubles * net.ngroups; i += 32) qdst[i] = __ldg(asrc + kGeomDoubles + i);
    if (sizeof(Extra) > 1 && lane == 31) mine->ex.set(load_event(ev, e));
    __syncwarp();
    if (FAST) {
        if (lane < net.fnd) scratch_set_fast(mine->sc, net, mine->geom, lane);
    } else if (lane < net.ndet) scratch_set(mine->sc, net, mine->geom, lane);
    __syncwarp();
}

// outputs of one Fisher pass (any pointer but `fisher` may be null)
// this rank's slot in the gathered buffer of every rank of the box (device pointers, peers mapped through CUDA IPC)
struct PeerSlots {
    double* p[kMaxPeers];
    int n;
};
struct FisherOut {
    double* fisher;       // [n][NPACK]
    double* snr2;         // [n]  4 int |h|^2/Sn df of the arms of the pass
    double* snr2_integ;   // [n]  the integral SNRInteg forms (signal.py:727): = snr2, except IMRPhenomHM (cross term dropped)
    double* snr_derivs;   // [n][NP]
    int* status;          // [n]  GWF_EV_* bits; the prologue stores the input bits, the kernel ORs in GWF_EV_NONFINITE_OUTPUT
    PeerSlots peers;      // multi-GPU: the finished packed row of every event is also stored into these [n][NPACK] slots (n = 0: none)
};

// Multi-GPU gather fused into the Fisher kernel: the finished packed row of event e goes to this rank's slot of every rank's
// gathered buffer as soon as the event is done -- 8-byte stores over NVLink spread over the whole lifetime of the kernel, so the
// exchange step of the path (one all-gather of the Fisher matrices) costs no time of its own.
__device__ __forceinline__ void forward_row(const PeerSlots& peers, long long e, int npack, int p, double v) {
    // fully unrolled with a predicate per slot: the slot pointers are then read from the kernel's parameter bank at compile-time
    // offsets.  A loop over peers.n indexed the by-value array dynamically, which made the compiler keep a copy of it in LOCAL memory --
    // three dependent local loads per slot and event that miss the L1 this kernel's 206 KB of shared memory leaves (measured with local
    // slots, scripts/forward_probe.py: +10 us per slot and launch of 1e4 events, 8 % at 8 slots)
    const long long off = e * npack + p;
#pragma unroll
    for (int q = 0; q < kMaxPeers; ++q)
        if (q < peers.n) peers.p[q][off] = v;
}

template <int MODEL, int NT, int FAST, bool SD, int SHAPE = 0>
#ifdef GWF_FISHER_MAXNREG
__global__ void __maxnreg__(GWF_FISHER_MAXNREG)
#else
__global__ void __launch_bounds__(kFisherThreads, GWF_FISHER_MINBLOCKS)
#endif
fisher_kernel(const typename ModelTraits<MODEL, NT>::Rec* __restrict__ recs, const EventAux* __restrict__ aux, EventsDev ev, long long n, int res, int lin,
              ModelCfg cfg, const __grid_constant__ NetworkDev net, const FisherOut fo, int pair) {
    double* __restrict__ out = fo.fisher;
    double* __restrict__ snr2_out = fo.snr2;
    double* __restrict__ sd_out = fo.snr_derivs;
    typedef typename ModelTraits<MODEL, NT>::Rec Rec;
    typedef typename PointFnsSel<MODEL, NT, SD>::type PF;
    typedef FisherSmem<Rec, typename PF::Extra> WS;
    constexpr int NP = NT + 7, NPACK = NP * (NP + 1) / 2;
    static_assert(NPACK <= 108, "entry-code table too small");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    WS* mine = reinterpret_cast<WS*>(smem_raw) + wid;
    const Rec& rec = mine->rec;
    if (PF::kEntryTable) {
        for (int p = lane; p < NPACK; p += 32) {
            int i = 0;
            while (tri(i + 1, 0) <= p) ++i;                   // row of packed index p
            PF::entry_code(i, p - tri(i, 0), mine->code + 4 * p);
        }
    }
    psd_cache_fill(net, smem_raw);
    // pair mode: warps w and w + kWarpsPerCta/2 take the two halves of an event (every other block of 32 samples each), so the
    // work unit of the persistent loop is half an event and its last round wastes half as much.  Each half rebuilds the Fisher
    // entries of its own partial sums (the entries are linear in the accumulators); the second warp hands its entries to the
    // first through shared memory at the pair's per-event barrier (see the end of the event loop).
    constexpr int kHalfWarps = kWarpsPerCta / 2;
    const int half = pair ? wid / kHalfWarps : 0, slot = pair ? wid % kHalfWarps : wid;
    const int per_cta = pair ? kHalfWarps : kWarpsPerCta, stride = pair ? 64 : 32, k0 = lane + 32 * half;
    int xpar = 0;
    for (long long e = (long long)blockIdx.x * per_cta + slot; e < n; e += (long long)gridDim.x * per_cta) {
        stage_event_aux<FAST != 0>(mine, recs, aux, ev, e, net, lane);
        const EvGeom& geom = mine->geom;
        double acc[PF::kAcc];
#pragma unroll
        for (int p = 0; p < PF::kAcc; ++p) acc[p] = 0.0;
        for (int g = 0; g < net.ngroups; ++g) {
            const Grid& grid = mine->grid[g];
            const bool rot = net.group_rot[g] != 0;
            // the walk constants are warp-uniform: re-read from shared memory (volatile) every step instead of occupying 16
            // registers next to the Gram accumulators
            const volatile Grid* gv = &mine->grid[g];
            FreqPoint fp;
            if (k0 < res) grid.start(k0, fp);
            // kUniformLoop: warp-uniform trip count with the last block predicated (needed for a barrier inside the loop);
            // otherwise lanes leave the loop on their own.  Same arithmetic either way; which form the compiler schedules
            // better differs per model (measured per 1e4 events: IMRPhenomD 1.229 -> 1.208 ms, NRTidalv2 3.131 -> 3.287 ms).
            constexpr bool kUniformLoop = MODEL == kPhenomHM || MODEL == kPhenomD;
            constexpr int kSampleUnroll = MODEL == kTaylorF2 ? GWF_UNR_TF2 : (MODEL == kPhenomD ? GWF_UNR_D : (MODEL == kNRTidalv2 ? GWF_UNR_NRT : 1));
            // kb0 runs over the blocks of `stride` samples; both halves of a pair make the same number of trips (the
            // block barrier below needs every arrival), the half whose block of 32 lies beyond the grid is predicated off
            // kSampleUnroll = 2 lets the scheduler interleave the Gram of one sample with the serial head (waveform, rotation phase, PSD row,
            // detector basis) of the next; 1 = rolled (experiment knobs GWF_UNR_*)
#pragma unroll(kSampleUnroll)
            for (int kb0 = 0; kUniformLoop ? kb0 < res : kb0 + k0 < res; kb0 += stride) {
                const int k = kb0 + k0;
                if (!kUniformLoop || k < res) {
                    if (k != k0) {
                        if (lin) grid.advance(k, fp);
                        else {
                            const bool last = k == res - 1;
                            fp.f = last ? gv->fcut : fp.f * gv->r;
                            fp.f13 *= gv->r13; fp.fm13 *= gv->rm13; fp.fm76 *= gv->rm76; fp.lnf += gv->dln;
                            fp.w = fp.f * (last ? gv->hw_hi : gv->hw_in);
                        }
                    }
                    if (FAST == 2) PF::template fisher_fast<true, SHAPE>(rec, cfg, geom, net, mine->sc, mine->ex, fp, acc);
                    else if (FAST == 1) PF::template fisher_fast<false, SHAPE>(rec, cfg, geom, net, mine->sc, mine->ex, fp, acc);
                    else PF::fisher(rec, cfg, geom, net, mine->sc, mine->ex, g, rot, fp, acc);
                }
                if (kUniformLoop && (pair & 4)) asm volatile("bar.sync %0, 64;" ::"r"(1 + slot) : "memory");
            }
        }
        // warp reduction by recursive halving; the reduced accumulators land in shared memory (the record is no longer
        // needed), from where every lane rebuilds its share of the packed Fisher matrix: coalesced stores
        typedef Fold<PF::kAcc> F;
        static_assert(sizeof(Rec) >= sizeof(double) * PF::kAcc, "record too small to hold the reduced accumulators");
        F::run(acc, lane);
        double* red = reinterpret_cast<double*>(&mine->rec);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < F::kOut; ++i) {
            const int idx = F::index(lane, i);
            if (idx >= 0) red[idx] = acc[i];
        }
        __syncwarp();
        double* o = out + e * NPACK;
        if (PF::kEntryTable) {
            for (int k = lane; k < kCompactCoefs; k += 32) mine->coef[k] = compact_coef(k, geom);
            __syncwarp();
        }
        // Every lane rebuilds its share of the packed entries from this warp's sums (the entries are linear in the accumulators).
        // One warp per event: they are the result.  Pair mode: the second warp of the pair leaves its entries (and its SNR sums) in
        // the first warp's hand-over slot, the pair meets at its per-event barrier -- which also keeps the two warps in the same
        // stretch of code: they sit on the same scheduler and share its instruction fetches (measured per 1e4 events, free-running
        // -> lock step per event: IMRPhenomD 1.251 -> 1.229 ms, NRTidalv2 3.166 -> 3.131 ms) -- and the first warp adds its own and
        // writes the event out: plain stores, no atomics, no zero-initialised output, and in a multi-GPU run the finished row goes to
        // the peers from registers.  a + b is commutative: the result is bitwise what two atomic additions to zero gave.
        constexpr int kVals = (NPACK + 31) / 32;
        constexpr int kXSnr = NPACK, kXInteg = NPACK + 1, kXSd = NPACK + 2;          // layout of a hand-over slot
        static_assert(kXSd + NP <= kXferDoubles, "hand-over slot too small");
        double vals[kVals];
#pragma unroll
        for (int i = 0; i < kVals; ++i) {
            const int p = lane + 32 * i;
            double v = 0.0;
            if (p < NPACK) {
                if (PF::kEntryTable) {
                    const uchar4 c = *reinterpret_cast<const uchar4*>(mine->code + 4 * p);
                    v = mine->coef[c.z] * red[c.x] + mine->coef[c.w] * red[c.y];
                } else v = red[p];
            }
            vals[i] = v;
        }
        double s_snr = (lane == 0 && snr2_out) ? PF::snr2(red, geom) : 0.0;
        double s_integ = (lane == 1 && fo.snr2_integ) ? PF::snr2_integ(red, geom) : 0.0;
        double s_sd = (sd_out && lane < NP) ? PF::snr_deriv(lane, red, geom) : 0.0;      // (h | d_i h), signal.py:938-945
        if (pair) {
            double* slot_x = reinterpret_cast<double*>(smem_raw + sizeof(WS) * kWarpsPerCta) + (slot * 2 + xpar) * kXferDoubles;
            if (half == 1) {
#pragma unroll
                for (int i = 0; i < kVals; ++i)
                    if (lane + 32 * i < NPACK) slot_x[lane + 32 * i] = vals[i];
                if (lane == 0) slot_x[kXSnr] = s_snr;
                if (lane == 1) slot_x[kXInteg] = s_integ;
                if (lane < NP) slot_x[kXSd + lane] = s_sd;
            }
            asm volatile("bar.sync %0, 64;" ::"r"(1 + slot) : "memory");
            if (half == 0) {
#pragma unroll
                for (int i = 0; i < kVals; ++i)
                    if (lane + 32 * i < NPACK) vals[i] += slot_x[lane + 32 * i];
                if (lane == 0) s_snr += slot_x[kXSnr];
                if (lane == 1) s_integ += slot_x[kXInteg];
                if (lane < NP) s_sd += slot_x[kXSd + lane];
            }
            xpar ^= 1;      // two slots: the second warp's next hand-over cannot overtake this read (it passes the next barrier first)
        }
        if (!pair || half == 0) {
            bool bad = false;
#pragma unroll
            for (int i = 0; i < kVals; ++i) {
                const int p = lane + 32 * i;
                if (p < NPACK) {
                    bad = bad || !isfinite(vals[i]);
                    o[p] = vals[i];
                    if (fo.peers.n) forward_row(fo.peers, e, NPACK, p, vals[i]);
                }
            }
            if (fo.status) {
                const unsigned anybad = __ballot_sync(0xffffffffu, bad);
                if (lane == 0 && anybad) atomicOr(fo.status + e, GWF_EV_NONFINITE_OUTPUT);
            }
            if (lane == 0 && snr2_out) snr2_out[e] = s_snr;
            if (lane == 1 && fo.snr2_integ) fo.snr2_integ[e] = s_integ;
            if (sd_out && lane < NP) sd_out[e * NP + lane] = s_sd;
        }
    }
}

// IMRPhenomHM, pair mode: the two warps of a pair share ONE staged event and work on the same blocks of 32 samples (hm_point_split:
// three modes each, exchanged partial sums, half of the packed entries each).  Outputs are added to the zero-initialised arrays like
// fisher_kernel's pair mode does; an entry is only ever added by the warp that owns it.
template <int NT, bool SD>
__global__ void __launch_bounds__(kFisherThreads, 1)
fisher_hm_split_kernel(const HMRec<NT>* __restrict__ recs, const EventAux* __restrict__ aux, EventsDev ev, long long n, int res, int lin, ModelCfg cfg,
                       const __grid_constant__ NetworkDev net, const FisherOut fo) {
    typedef HMRec<NT> Rec;
    typedef HMSplitAcc<NT, SD> L;
    typedef WarpSmem<Rec, HMExtra> WS;
    constexpr int NP = NT + 7, NPACK = NP * (NP + 1) / 2;
    constexpr int kPairs = kWarpsPerCta / 2, kW = HMStrainSink<NT>::kWords;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int half = wid / kPairs, slot = wid % kPairs;          // warps w and w + kPairs sit on the same scheduler
    WS* mine = reinterpret_cast<WS*>(smem_raw) + slot;
    double* xbuf = reinterpret_cast<double*>(smem_raw + sizeof(WS) * kPairs) + slot * (2 * kW * 32);
    const Rec& rec = mine->rec;
    const int bar_id = 1 + slot;
    psd_cache_fill(net, smem_raw);
    for (long long e = (long long)blockIdx.x * kPairs + slot; e < n; e += (long long)gridDim.x * kPairs) {
        // both warps stage the event into the shared block (identical values; neither reads before its own copy is complete)
        stage_event_aux<false>(mine, recs, aux, ev, e, net, lane);
        const EvGeom& geom = mine->geom;
        double acc[L::kAcc];
#pragma unroll
        for (int p = 0; p < L::kAcc; ++p) acc[p] = 0.0;
        for (int g = 0; g < net.ngroups; ++g) {
            const Grid& grid = mine->grid[g];
            const bool rot = net.group_rot[g] != 0;
            const volatile Grid* gv = &mine->grid[g];
            FreqPoint fp;
            if (lane < res) grid.start(lane, fp);
            for (int kb0 = 0; kb0 < res; kb0 += 32) {
                const int k = kb0 + lane;
                const bool active = k < res;
                if (active && kb0 != 0) {
                    if (lin) grid.advance(k, fp);
                    else {
                        const bool last = k == res - 1;
                        fp.f = last ? gv->fcut : fp.f * gv->r;
                        fp.f13 *= gv->r13; fp.fm13 *= gv->rm13; fp.fm76 *= gv->rm76; fp.lnf += gv->dln;
                        fp.w = fp.f * (last ? gv->hw_hi : gv->hw_in);
                    }
                }
                hm_point_split<NT, SD>(rec, cfg, geom, net, mine->sc, mine->ex, g, rot, fp, acc, half, xbuf, bar_id, lane, active);
            }
        }
        // both warps are past the last exchange: nothing reads the record any more, its space takes the reduced accumulators
        typedef Fold<L::kAcc> F;
        F::run(acc, lane);
        double* red = reinterpret_cast<double*>(&mine->rec) + half * 64;
        static_assert(L::kAcc <= 64 && sizeof(Rec) >= sizeof(double) * 128, "reduced accumulators of the two halves live in the record's space");
        __syncwarp();
#pragma unroll
        for (int i = 0; i < F::kOut; ++i) {
            const int idx = F::index(lane, i);
            if (idx >= 0) red[idx] = acc[i];
        }
        __syncwarp();
        // every entry has exactly one owner: plain stores (and, in a multi-GPU run, the same value to the peers)
        double* o = fo.fisher + e * NPACK;
        bool bad = false;
        for (int p = 2 * lane + half; p < NPACK; p += 64) {
            const double v = red[p >> 1];
            bad = bad || !isfinite(v);
            o[p] = v;
            if (fo.peers.n) forward_row(fo.peers, e, NPACK, p, v);
        }
        if (fo.status) {
            const unsigned anybad = __ballot_sync(0xffffffffu, bad);
            if (lane == 0 && anybad) atomicOr(fo.status + e, GWF_EV_NONFINITE_OUTPUT);
        }
        if (lane == 0) {
            double* dst = half == 0 ? fo.snr2 : fo.snr2_integ;
            if (dst) dst[e] = red[L::kSnr];
        }
        if (SD && fo.snr_derivs && lane < NP && (lane & 1) == half) fo.snr_derivs[e * NP + lane] = red[L::kSd + (lane >> 1)];
        // the shared block is restaged for the next event only when both warps are done with this one
        asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
    }
}

// return_derivatives (signal.py:917-945): the derivative strain d h / d p_i of ONE arm (a per-arm pass network) written out as
// complex128 [nP][n][res] -- the array the Fisher kernel deliberately never materialises.  Same rows as arm_rows / Compact, times
// A e^{i Psi} with Psi = 2 pi f tcoal 86400 - Phicoal - Phi(f) + 2 pi f Delta t (signal.py:484, 580, 641); HBM-write bound.
template <int MODEL, int NT>
__global__ void __launch_bounds__(kFisherThreads)
derivs_kernel(const typename ModelTraits<MODEL, NT>::Rec* __restrict__ recs, EventsDev ev, long long n, int res, int lin, ModelCfg cfg,
              const __grid_constant__ NetworkDev net, double2* __restrict__ out) {
    typedef typename ModelTraits<MODEL, NT>::Rec Rec;
    typedef PointFns<MODEL, NT> PF;
    typedef WarpSmem<Rec, typename PF::Extra> WS;
    typedef Compact<NT> CP;
    constexpr int NP = NT + 7;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    WS* mine = reinterpret_cast<WS*>(smem_raw) + wid;
    const Rec& rec = mine->rec;
    const long long nwarps = (long long)gridDim.x * kWarpsPerCta;
    for (long long e = (long long)blockIdx.x * kWarpsPerCta + wid; e < n; e += nwarps) {
        double phicoal;
        {
            EvGeom g0;
            const EventIn in = load_event(ev, e);
            phicoal = in.Phicoal;
            g0.set(in);
            stage_event<false>(mine, recs, e, net, g0, in, lane);
        }
        const EvGeom& geom = mine->geom;
        // the one detector / arm of this pass
        int di = 0;
        while (di < net.ndet - 1 && net.det[di].arm_begin == net.det[di].arm_end) ++di;
        const DetDev& d = net.det[di];
        const ArmDev& arm = net.arm[d.arm_begin];
        const int g = d.group;
        double fcut = rec.fcut_hz;
        if (net.group_fmax[g] > 0.0 && fcut > net.group_fmax[g]) fcut = net.group_fmax[g];
        Grid grid;
        grid.set(net.group_fmin[g], fcut, res, lin != 0, 32);
        const bool rot = d.use_rot != 0;
        for (int k = lane; k < res; k += 32) {
            FreqPoint fp;
            grid.start(k, fp);
            double2 row[NP];
#pragma unroll
            for (int i = 0; i < NP; ++i) row[i] = make_double2(0.0, 0.0);
            if constexpr (MODEL == kPhenomHM) {
                // rows of hm_arm_rows times the carrier e^{i (2 pi f tcoal 86400 - Phicoal + 2 pi f Delta t)} (signal.py:584-607, 1378-1380)
                HMStrainSink<NT> hs(mine->ex.w);
                phenomhm_foreach_mode<NT, true>(rec, g, fp, !(cfg.flags & kFlagNoFcut), hs);
                if (!(hs.hpr == 0.0 && hs.hpi == 0.0 && hs.hcr == 0.0 && hs.hci == 0.0)) {
                    PointWf<NT> w;
                    w.f = fp.f;
#pragma unroll
                    for (int j = 0; j < NT; ++j) w.phi_d[j] = 0.;
                    w.dtn[0] = w.dtn[1] = 0.;
                    double sBr = 0., cBr = 1.;
                    if (rot) {
                        double tau, dtau[2];
                        const double xm13 = rec.sp.sm13 * fp.fm13, lpx3 = fma(fp.lnf, 1. / 3., rec.sp.lps3);
                        tau_eval(rec.tau, 0.68278406325529568146702083315816 * xm13, lpx3, rec.lam, tau, dtau);
                        w.dtn[0] = -dtau[0] * kInvDay;
                        w.dtn[1] = -dtau[1] * kInvDay;
                        sincos(2.0 * kPi * fma(-tau, kInvDay, geom.tcoal), &sBr, &cBr);
                    }
                    DetPoint dp;
                    if (rot) det_point(mine->sc.ed[di], cBr, sBr, dp);
                    else dp = mine->sc.fixed[di];
                    DetRows<NT> dr;
                    dr.set(w, dp, rot, d.no_motion != 0);
                    double ra[NP], rb[NP], Hr, Hi, Fp, Fc;
                    hm_arm_rows<NT>(hs, dp, dr, arm, geom, ra, rb, Hr, Hi, Fp, Fc);
                    const double W2 = 2.0 * kPi * fp.f;
                    double sP, cP;
                    sincos(fma(W2, dp.dt, W2 * (geom.tcoal * 3600. * 24.) - phicoal), &sP, &cP);
#pragma unroll
                    for (int i = 0; i < NP; ++i) row[i] = make_double2(ra[i] * cP - rb[i] * sP, ra[i] * sP + rb[i] * cP);
                }
            } else {
            PointWf<NT> w;
            ModelTraits<MODEL, NT>::eval(rec, cfg, g, fp, rot, w);
            w.f = fp.f;
            if (w.A != 0.0) {
                double sBr = 0., cBr = 1.;
                if (rot) sincos(2.0 * kPi * fma(-w.tau, kInvDay, geom.tcoal), &sBr, &cBr);
                DetPoint dp;
                if (rot) det_point(mine->sc.ed[di], cBr, sBr, dp);
                else dp = mine->sc.fixed[di];
                DetRows<NT> dr;
                dr.set(w, dp, rot, d.no_motion != 0);
                double ra[CP::NG], rb[CP::NG], u, v;
                arm_rows<NT>(w, dp, dr, arm, geom, ra, rb, u, v);
                const double W2 = 2.0 * kPi * fp.f;
                const double Psi = (W2 * (geom.tcoal * 3600. * 24.) - phicoal - w.phi) + W2 * dp.dt;
                double sP, cP;
                sincos(Psi, &sP, &cP);
                const double zr = w.A * cP, zi = w.A * sP;
                const double al[4] = {-geom.K * geom.inv_dL, -geom.ci * geom.si, -2.0 * geom.ci, -geom.K};
                const double be[4] = {-geom.ci * geom.inv_dL, -geom.si, 2.0 * geom.K, geom.ci};
#pragma unroll
                for (int i = 0; i < NP; ++i) {
                    const int gx = CP::g_of(i);
                    double a, b;
                    if (gx >= 0) { a = ra[gx]; b = rb[gx]; }
                    else {
                        const int sidx = -gx - 1;
                        if (sidx < 2) { a = al[sidx] * u; b = be[sidx] * v; }
                        else { a = be[sidx] * v; b = al[sidx] * u; }
                    }
                    row[i] = make_double2(a * zr - b * zi, a * zi + b * zr);
                }
            }
            }
#pragma unroll
            for (int i = 0; i < NP; ++i) out[((long long)i * n + e) * res + k] = row[i];
        }
    }
}

// GWSignal.GWstrain on the engine's own grid (signal.py:484-655): the strain h = A e^{i Psi} (K F+ + i cos(iota) Fx) of ONE arm
// (a per-arm pass network) written out as complex128 [n][res]; samples beyond the waveform cut are 0.  Used by WFOverlap, whose
// two waveforms differ in model and parameters but share the grid (the host passes the common upper frequency as fcut_host).
template <int MODEL>
__global__ void __launch_bounds__(kFisherThreads)
strain_kernel(const typename ModelTraits<MODEL, 4>::Rec* __restrict__ recs, EventsDev ev, long long n, int res, int lin, ModelCfg cfg,
              const __grid_constant__ NetworkDev net, double2* __restrict__ out) {
    constexpr int NT = 4;
    typedef typename ModelTraits<MODEL, NT>::Rec Rec;
    typedef PointFns<MODEL, NT> PF;
    typedef WarpSmem<Rec, typename PF::Extra> WS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    WS* mine = reinterpret_cast<WS*>(smem_raw) + wid;
    const Rec& rec = mine->rec;
    const long long nwarps = (long long)gridDim.x * kWarpsPerCta;
    for (long long e = (long long)blockIdx.x * kWarpsPerCta + wid; e < n; e += nwarps) {
        double phicoal;
        {
            EvGeom g0;
            const EventIn in = load_event(ev, e);
            phicoal = in.Phicoal;
            g0.set(in);
            stage_event<false>(mine, recs, e, net, g0, in, lane);
        }
        const EvGeom& geom = mine->geom;
        int di = 0;
        while (di < net.ndet - 1 && net.det[di].arm_begin == net.det[di].arm_end) ++di;
        const DetDev& d = net.det[di];
        const ArmDev& arm = net.arm[d.arm_begin];
        const int g = d.group;
        double fcut = rec.fcut_hz;
        if (net.group_fmax[g] > 0.0 && fcut > net.group_fmax[g]) fcut = net.group_fmax[g];
        Grid grid;
        grid.set(net.group_fmin[g], fcut, res, lin != 0, 32);
        const bool rot = d.use_rot != 0;
        for (int k = lane; k < res; k += 32) {
            FreqPoint fp;
            grid.start(k, fp);
            double2 h = make_double2(0.0, 0.0);
            if constexpr (MODEL == kPhenomHM) {
                // h = (hp Fp + hc Fc) e^{i (2 pi f Delta t + 2 pi f tcoal 86400 - Phicoal)}, signal.py:584-607
                HMValueSink hs(mine->ex.w);
                phenomhm_foreach_mode<NT, false>(rec, g, fp, !(cfg.flags & kFlagNoFcut), hs);
                if (!(hs.hpr == 0.0 && hs.hpi == 0.0 && hs.hcr == 0.0 && hs.hci == 0.0)) {
                    double sBr = 0., cBr = 1.;
                    if (rot) {
                        double tau, dtau[2];
                        const double xm13 = rec.sp.sm13 * fp.fm13, lpx3 = fma(fp.lnf, 1. / 3., rec.sp.lps3);
                        tau_eval(rec.tau, 0.68278406325529568146702083315816 * xm13, lpx3, rec.lam, tau, dtau);
                        sincos(2.0 * kPi * fma(-tau, kInvDay, geom.tcoal), &sBr, &cBr);
                    }
                    DetPoint dp;
                    if (rot) det_point(mine->sc.ed[di], cBr, sBr, dp);
                    else dp = mine->sc.fixed[di];
                    double Fp, Fc;
                    arm_pattern(dp, arm, geom, Fp, Fc);
                    const double W2 = 2.0 * kPi * fp.f;
                    double sP, cP;
                    sincos(fma(W2, dp.dt, W2 * (geom.tcoal * 3600. * 24.) - phicoal), &sP, &cP);
                    const double zr = hs.hpr * Fp + hs.hcr * Fc, zi = hs.hpi * Fp + hs.hci * Fc;
                    h = make_double2(zr * cP - zi * sP, zr * sP + zi * cP);
                }
            } else {
            PointWf<NT> w;
            ModelTraits<MODEL, NT>::eval(rec, cfg, g, fp, rot, w);
            if (w.A != 0.0) {
                double sBr = 0., cBr = 1.;
                if (rot) sincos(2.0 * kPi * fma(-w.tau, kInvDay, geom.tcoal), &sBr, &cBr);
                DetPoint dp;
                if (rot) det_point(mine->sc.ed[di], cBr, sBr, dp);
                else dp = mine->sc.fixed[di];
                double Fp, Fc;
                arm_pattern(dp, arm, geom, Fp, Fc);
                const double W2 = 2.0 * kPi * fp.f;
                const double Psi = (W2 * (geom.tcoal * 3600. * 24.) - phicoal - w.phi) + W2 * dp.dt;      // signal.py:484, 580, 641
                double sP, cP;
                sincos(Psi, &sP, &cP);
                const double ar = w.A * geom.K * Fp, ai = w.A * geom.ci * Fc;                            // Ap, Ac (signal.py:463-464)
                h = make_double2(ar * cP - ai * sP, ar * sP + ai * cP);
            }
            }
            out[e * res + k] = h;
        }
    }
}

// GWSignal.WFOverlap integrals (signal.py:1866-1925) for one arm: 4 int Re(h1 conj h2)/Sn df, 4 int |h1|^2/Sn df, 4 int |h2|^2/Sn df on the
// grid geomspace(fmin, fcut[e], res); one warp per event, trapezoid weights as in the Fisher/SNR kernels.
__global__ void __launch_bounds__(256) overlap_kernel(const double2* __restrict__ h1, const double2* __restrict__ h2, const double* __restrict__ fcut,
                                                      long long n, int res, double fmin, PsdDev psd, double* __restrict__ ov, double* __restrict__ s1,
                                                      double* __restrict__ s2) {
    const int lane = threadIdx.x & 31;
    const long long e = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    if (e >= n) return;
    Grid grid;
    grid.set(fmin, fcut[e], res, false, 32);
    double a = 0., b = 0., c = 0.;
    for (int k = lane; k < res; k += 32) {
        FreqPoint fp;
        grid.start(k, fp);
        const double wgt = 4.0 * fp.w / psd_lookup(psd, fp.f, fp.lnf * 1.4426950408889634073599246810018921);
        const double2 x = h1[e * res + k], y = h2[e * res + k];
        a = fma(wgt, x.x * y.x + x.y * y.y, a);
        b = fma(wgt, x.x * x.x + x.y * x.y, b);
        c = fma(wgt, y.x * y.x + y.y * y.y, c);
    }
    a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
    if (lane == 0) { ov[e] = a; s1[e] = b; s2[e] = c; }
}

// SNR: per-arm integrals, value-only.  kSplit warps share an event (warp `sub` takes every kSplit-th block of 32 samples), so a
// CTA of 16 warps stages only 16/kSplit coefficient records and the PSD windows still fit next to them in shared memory.
// Warps per event: two where the staging blocks of eight events still leave room for the PSD windows (measured, prologue + kernel per
// 1e4 events, 4 -> 2: IMRPhenomD 0.460 -> 0.444 ms, TaylorF2 0.239 -> 0.228 ms), four for IMRPhenomHM (2.62 vs 2.72 ms)
template <int MODEL> struct SnrMap {
    static constexpr int kSplit = MODEL == kPhenomHM ? 4 : 2;
    static constexpr int kGroups = kSnrWarps / kSplit;
};
template <int MODEL, int FAST, int SHAPE = 0>
__global__ void __launch_bounds__(kSnrThreads, 1)
snr_kernel(const typename ModelTraits<MODEL, 4>::Rec* __restrict__ recs, const EventAux* __restrict__ aux, EventsDev ev, long long n, int res, int lin,
           ModelCfg cfg, const __grid_constant__ NetworkDev net, int narm_out, double* __restrict__ snr2_arm) {
    typedef typename ModelTraits<MODEL, 4>::Rec Rec;
    typedef PointFns<MODEL, 4> PF;
    typedef WarpSmem<Rec, typename PF::Extra> WS;
    constexpr int kSnrSplit = SnrMap<MODEL>::kSplit, kSnrGroups = SnrMap<MODEL>::kGroups;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int slot = wid % kSnrGroups, sub = wid / kSnrGroups, gt = sub * 32 + lane;      // gt: thread index inside the group
    WS* mine = reinterpret_cast<WS*>(smem_raw) + slot;
    // per-lane, per-arm accumulators [warp][arm][lane], then the per-warp totals [group][sub][arm]
    double* s2_base = reinterpret_cast<double*>(reinterpret_cast<WS*>(smem_raw) + kSnrGroups);
    double* s2 = s2_base + (wid * narm_out) * 32 + lane;
    double* comb = s2_base + kSnrWarps * narm_out * 32 + slot * kSnrSplit * narm_out;
    const Rec& rec = mine->rec;
    psd_cache_fill(net, smem_raw);
    constexpr int kRecDoubles = (int)(sizeof(Rec) / sizeof(double));
    constexpr int stride = 32 * kSnrSplit;
    const int k0 = gt;
    for (long long e = (long long)blockIdx.x * kSnrGroups + slot; e < n; e += (long long)gridDim.x * kSnrGroups) {
        {
            // the threads of the group stage the event: coefficient record and the prologue's EventAux (geometry, grids) by
            // coalesced copies, then the detector scratch from the staged geometry
            constexpr int kGeomDoubles = (int)(sizeof(EvGeom) / sizeof(double)), kGridDoubles = (int)(sizeof(Grid) / sizeof(double));
            const double* src = reinterpret_cast<const double*>(recs + e);
            const double* asrc = reinterpret_cast<const double*>(aux + e);
            double* dst = reinterpret_cast<double*>(&mine->rec);
            double* gdst = reinterpret_cast<double*>(&mine->geom);
            double* qdst = reinterpret_cast<double*>(&mine->grid[0]);
            for (int i = gt; i < kRecDoubles; i += stride) dst[i] = __ldg(src + i);
            if (sub == 0) {
                if (lane < kGeomDoubles) gdst[lane] = __ldg(asrc + lane);
                for (int i = lane; i < kGridDoubles * net.ngroups; i += 32) qdst[i] = __ldg(asrc + kGeomDoubles + i);
                if (sizeof(typename PF::Extra) > 1 && lane == 31) mine->ex.set(load_event(ev, e));
                __syncwarp();
                if (FAST) {
                    if (lane < net.fnd) scratch_set_fast(mine->sc, net, mine->geom, lane);
                } else if (lane < net.ndet) scratch_set(mine->sc, net, mine->geom, lane);
            }
        }
        asm volatile("bar.sync %0, %1;" ::"r"(1 + slot), "r"(stride) : "memory");
        const EvGeom& geom = mine->geom;
        // per-arm sums: registers when the shape (hence every arm index) is a compile-time constant, else [arm][lane] in shared memory
        constexpr bool kRegSums = FAST != 0 && SHAPE != 0;
        constexpr int kRegArms = kRegSums ? shape_total(SHAPE) : 1;
        double s2r[kRegArms];
#pragma unroll
        for (int a = 0; a < kRegArms; ++a) s2r[a] = 0.0;
        if (!kRegSums)
            for (int a = 0; a < narm_out; ++a) s2[a * 32] = 0.0;
        for (int g = 0; g < net.ngroups; ++g) {
            const Grid& grid = mine->grid[g];
            const bool rot = net.group_rot[g] != 0;
            FreqPoint fp;
            if (k0 < res) grid.start(k0, fp);
            for (int k = k0; k < res; k += stride) {
                if (k != k0) grid.advance(k, fp);
                if (FAST == 2) PF::template snr_fast<true, SHAPE>(rec, cfg, geom, net, mine->sc, mine->ex, fp, kRegSums ? s2r : s2);
                else if (FAST == 1) PF::template snr_fast<false, SHAPE>(rec, cfg, geom, net, mine->sc, mine->ex, fp, kRegSums ? s2r : s2);
                else PF::snr(rec, cfg, geom, net, mine->sc, mine->ex, g, rot, fp, s2);
            }
        }
        if (kRegSums) {
#pragma unroll
            for (int a = 0; a < kRegArms; ++a) {
                const double v = warp_sum(s2r[a]);
                if (lane == 0) comb[sub * narm_out + a] = v;
            }
        } else {
            for (int a = 0; a < narm_out; ++a) {
                const double v = warp_sum(s2[a * 32]);
                if (lane == 0) comb[sub * narm_out + a] = v;
            }
        }
        // all warps of the group are done with the record and have left their totals
        asm volatile("bar.sync %0, %1;" ::"r"(1 + slot), "r"(stride) : "memory");
        if (sub == 0)
            for (int a = lane; a < narm_out; a += 32) {
                double v = 0.0;
                for (int w = 0; w < kSnrSplit; ++w) v += comb[w * narm_out + a];
                snr2_arm[(long long)a * n + e] = v;
            }
        // comb is next written after the following event's staging barrier, which warp 0 only reaches after this read
    }
}

// FP64 FMA throughput probe: 8 independent chains per thread, all in registers
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    const double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 12345.678) out[0] = s;
}

// packed [n][npack] -> (nP, nP, ld) planes, event axis fastest; ld >= n lets a chunk of events land in its slice of a larger array.
// One thread block transposes a tile of 32 events through shared memory: coalesced reads of the packed rows, coalesced plane writes.
__global__ void __launch_bounds__(256) unpack_kernel(const double* __restrict__ packed, long long n, int nP, double* __restrict__ full, long long ld) {
    __shared__ double tile[32][109];                 // up to nP = 14 (npack = 105), odd pitch: conflict-free columns
    const int npack = nP * (nP + 1) / 2;
    const long long e0 = (long long)blockIdx.x * 32;
    const int ne = (int)((n - e0) < 32 ? (n - e0) : 32);
    const double* src = packed + e0 * npack;
    for (int t = threadIdx.x; t < ne * npack; t += blockDim.x) tile[t / npack][t % npack] = src[t];
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int q = wid; q < nP * nP; q += 8) {
        const int i = q / nP, j = q % nP;
        if (lane < ne) full[(long long)q * ld + e0 + lane] = tile[lane][i >= j ? tri(i, j) : tri(j, i)];
    }
}

// Multi-GPU: the same transposition fused with the write side of an all-gather over NVLink peer memory.  The packed rows of this
// rank's events are also stored into this rank's slot of every peer's gathered buffer (peer pointers obtained by the caller through
// CUDA IPC): plain coalesced 16-byte stores that the NVSwitch fabric routes to the peers while the tile is transposed -- no
// separate collective kernel, no copy of the result through a communication buffer, and the stores of one tile overlap the
// loads of the next.  The caller synchronises the ranks (any barrier) before the gathered buffers are read.
__global__ void __launch_bounds__(256) unpack_gather_kernel(const double* __restrict__ packed, long long n, int nP, double* __restrict__ full, long long ld,
                                                            const PeerSlots peers) {
    __shared__ __align__(16) double tile[32][110];   // even pitch: rows stay 16-byte aligned for the vector stores
    const int npack = nP * (nP + 1) / 2;
    const long long e0 = (long long)blockIdx.x * 32;
    const int ne = (int)((n - e0) < 32 ? (n - e0) : 32);
    const double* src = packed + e0 * npack;
    const int total = ne * npack;
    for (int t = threadIdx.x; t < total; t += blockDim.x) {
        const double v = src[t];
        tile[t / npack][t % npack] = v;
    }
    // the tile's rows are contiguous in the packed array: forward them to the peers as they are
    if (((e0 * npack) & 1) == 0) {
        const double2* s2 = reinterpret_cast<const double2*>(src);
#pragma unroll
        for (int q = 0; q < kMaxPeers; ++q) {          // unrolled: slot pointers from the parameter bank, not from a local copy (see forward_row)
            if (q >= peers.n) break;
            double2* d2 = reinterpret_cast<double2*>(peers.p[q] + e0 * npack);
            for (int t = threadIdx.x; t < total / 2; t += blockDim.x) d2[t] = s2[t];
            if ((total & 1) && threadIdx.x == 0) peers.p[q][e0 * npack + total - 1] = src[total - 1];
        }
    } else {
#pragma unroll
        for (int q = 0; q < kMaxPeers; ++q) {
            if (q >= peers.n) break;
            for (int t = threadIdx.x; t < total; t += blockDim.x) peers.p[q][e0 * npack + t] = src[t];
        }
    }
    __syncthreads();
    if (full) {
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        for (int q = wid; q < nP * nP; q += 8) {
            const int i = q / nP, j = q % nP;
            if (lane < ne) full[(long long)q * ld + e0 + lane] = tile[lane][i >= j ? tri(i, j) : tri(j, i)];
        }
    }
}

// ------------------------------------------------------------------------------------------- host side
// Per-device state: SM count, the device copy of the QNM tables, and the largest dynamic shared memory size already
// requested for each kernel -- looked up once per (device, kernel) instead of on every call.  Tables set with
// gwf_set_qnm_tables and PSD handles keep a host copy and are uploaded to a device the first time a call runs there.
struct DeviceCtx {
    int sms = 0;
    QnmTables qnm = {nullptr, nullptr, nullptr, 0, nullptr};
    int qnm_version = 0;
    std::unordered_map<const void*, size_t> smem_set;
};
static std::mutex g_mu;
static DeviceCtx g_ctx[kMaxDevices];
static std::vector<double> g_qnm_host;       // a | fring | fdamp
static int g_qnm_n = 0, g_qnm_version = 0;

// context of the current device (needs_qnm: make sure the device copy of the QNM tables is current)
static int device_ctx(int needs_tables, DeviceCtx** out, int* dev_out = nullptr) {
    const bool needs_qnm = (needs_tables & 1) != 0;
    int dev = 0;
    GWF_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices) return fail(GWF_ERR_ARG, "device index out of range");
    std::lock_guard<std::mutex> lock(g_mu);
    DeviceCtx& c = g_ctx[dev];
    if (c.sms == 0) GWF_CUDA(cudaDeviceGetAttribute(&c.sms, cudaDevAttrMultiProcessorCount, dev));
    if (needs_qnm) {
        if (g_qnm_n == 0) return fail(GWF_ERR_ARG, "QNM tables not set (gwf_set_qnm_tables)");
        if (c.qnm_version != g_qnm_version) {
            if (c.qnm.a) cudaFree(const_cast<double*>(c.qnm.a));
            double* buf = nullptr;
            GWF_CUDA(cudaMalloc(&buf, sizeof(double) * 3 * g_qnm_n));
            GWF_CUDA(cudaMemcpy(buf, g_qnm_host.data(), sizeof(double) * 3 * g_qnm_n, cudaMemcpyHostToDevice));
            c.qnm.a = buf; c.qnm.fring = buf + g_qnm_n; c.qnm.fdamp = buf + 2 * g_qnm_n; c.qnm.n = g_qnm_n;
            c.qnm_version = g_qnm_version;
        }
    }
    if ((needs_tables & 2) && !c.qnm.xitide) {
        // IMRPhenomNSBH: the reference interpolates a 200^3 table of xi_tide that it reads from an HDF5 file or tabulates with numpy.roots
        // (waveforms.py:3286-3373); here the device solves the 8e6 order-10 polynomials itself, once per device (64 MB of HBM)
        double* tab = nullptr;
        const int nodes = kXiRes * kXiRes * kXiRes;
        GWF_CUDA(cudaMalloc(&tab, sizeof(double) * (size_t)nodes));
        xitide_table_kernel<<<(nodes + 127) / 128, 128>>>(tab);
        GWF_CUDA(cudaGetLastError());
        GWF_CUDA(cudaDeviceSynchronize());
        c.qnm.xitide = tab;
    }
    *out = &c;
    if (dev_out) *dev_out = dev;
    return GWF_OK;
}
// raise a kernel's dynamic shared memory limit (only when this device has not granted at least `bytes` to it yet)
template <class Kern> static int ensure_smem(DeviceCtx& c, Kern kern, size_t bytes) {
    std::lock_guard<std::mutex> lock(g_mu);
    size_t& have = c.smem_set[reinterpret_cast<const void*>(kern)];
    if (bytes > have) {
        GWF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        have = bytes;
    }
    return GWF_OK;
}

static int collect_psds(const gwf_psd* const* psds, int npsd, PsdDev* out) {
    if (npsd < 1 || npsd > kMaxPsd) return fail(GWF_ERR_ARG, "number of PSD tables must be in [1, 8]");
    int dev = 0;
    GWF_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices) return fail(GWF_ERR_ARG, "device index out of range");
    for (int i = 0; i < npsd; ++i) {
        if (!psds[i]) return fail(GWF_ERR_ARG, "null PSD handle");
        PsdHost* h = const_cast<PsdHost*>(reinterpret_cast<const PsdHost*>(psds[i]));
        {
            std::lock_guard<std::mutex> lock(g_mu);
            if (!h->tab[dev]) {           // first use on this device
                GWF_CUDA(cudaMalloc(&h->tab[dev], sizeof(double4) * h->tab_h.size()));
                GWF_CUDA(cudaMalloc(&h->bucket[dev], sizeof(int) * h->bucket_h.size()));
                GWF_CUDA(cudaMemcpy(h->tab[dev], h->tab_h.data(), sizeof(double4) * h->tab_h.size(), cudaMemcpyHostToDevice));
                GWF_CUDA(cudaMemcpy(h->bucket[dev], h->bucket_h.data(), sizeof(int) * h->bucket_h.size(), cudaMemcpyHostToDevice));
            }
        }
        out[i] = h->dev;
        out[i].tab = h->tab[dev];
        out[i].bucket = h->bucket[dev];
    }
    return GWF_OK;
}

// which part of a Fisher call to run, on which events: the workspace (records + EventAux) is laid out for n_total events; the
// prologue (phase bit 1) always covers all of them, the kernels (phase bit 2) the events [lo, lo + m) -- outputs then refer to that range
struct FisherRange {
    long long lo, m;
    int phases;
};
template <int MODEL, int NT>
static int run_fisher(const gwf_model* model, const gwf_detector* dets, int ndet, const gwf_psd* const* psds, int npsd, const EventsDev& ev_all,
                      long long n_total, const FisherRange& range, const gwf_opts* opts, const gwf_fisher_out* outp, void* ws, size_t ws_bytes,
                      cudaStream_t st) {
    typedef typename ModelTraits<MODEL, NT>::Rec Rec;
    constexpr int NP = NT + 7, NPACK = NP * (NP + 1) / 2;
    double* fisher = outp->fisher_packed;
    double* snr2 = outp->snr2;
    double* snr_derivs = outp->snr_derivs;
    const size_t rec_bytes = (sizeof(Rec) * (size_t)n_total + 15) & ~(size_t)15;
    if (ws_bytes < rec_bytes + sizeof(EventAux) * (size_t)n_total) return fail(GWF_ERR_WORKSPACE, "workspace too small");
    Rec* recs_all = reinterpret_cast<Rec*>(ws);
    EventAux* aux_all = reinterpret_cast<EventAux*>(reinterpret_cast<char*>(ws) + rec_bytes);
    // the kernels' view: the events of the range
    const long long n = range.m;
    Rec* recs = recs_all + range.lo;
    EventAux* aux = aux_all + range.lo;
    EventsDev ev;
    for (int i = 0; i < GWF_NPARAM_IN; ++i) ev.p[i] = ev_all.p[i] ? ev_all.p[i] + range.lo : nullptr;
    ModelCfg cfg = {model->id, model->flags, model->fcutPar, model->fRef};
    DeviceCtx* ctx = nullptr;
    if (int rc0 = device_ctx(model_tables<MODEL>(), &ctx)) return rc0;
    NetworkDev net;
    PsdDev pd[kMaxPsd];
    int rc = collect_psds(psds, npsd, pd);
    if (rc) return rc;
    rc = build_network(dets, ndet, pd, npsd, -1, false, net);
    if (rc) return rc;
    GroupInfo gi = group_info(net);
    // One event per pair of warps (each warp takes every other block of 32 samples): the work unit of the persistent loop is
    // half an event, which shortens its last round (1e4 events on 1184 warps: 8.45 -> 9 rounds of whole events, 16.9 -> 17 of
    // halves; measured 1.36 -> 1.31 ms).  The mapping is fixed per model -- never chosen from n -- so that an event's result
    // does not depend on the size of the batch it is computed in.  TaylorF2 events are too cheap for the repeated staging
    // (measured 4 % slower), so they keep one warp per event.
    int pair = (MODEL != kTaylorF2 && !(opts->flags & GWF_OPT_ONE_WARP_PER_EVENT)) ? 1 : 0;
    // IMRPhenomHM pairs split the modes and the packed entries of the SAME samples (fisher_hm_split_kernel) instead of taking
    // alternate blocks of samples
    constexpr bool kSplitPair = MODEL == kPhenomHM;
    // lock step of the two halves: a named barrier per event (bit 1), for IMRPhenomHM per block of samples (bit 2) -- see the
    // comment at the end of fisher_kernel's event loop for the measurements behind the choice
    if (pair) pair |= (MODEL == kPhenomHM) ? 6 : 2;
    const int lin = (opts->flags & GWF_OPT_LIN_GRID) ? 1 : 0;
    const int pb = 128;
    if (range.phases & 1) {
        if (!(opts->flags & GWF_OPT_REUSE_WORKSPACE)) {
            // the prologue also leaves the per-event geometry and grids (EventAux) for the Fisher kernel
            AuxPlan ap;
            ap.out = aux_all;
            for (int g = 0; g < kMaxGroups; ++g) ap.fmax[g] = net.group_fmax[g];
            ap.res = opts->res; ap.lin = lin; ap.stride = (pair && !(kSplitPair && !(opts->flags & GWF_OPT_HM_BLOCK_PAIRS))) ? 64 : 32;
            prologue_kernel<MODEL, NT><<<dim3((unsigned)((n_total + pb - 1) / pb), (MODEL == kPhenomD || MODEL == kNRTidalv2) ? 2 : 1), pb, 0, st>>>(
                ev_all, n_total, cfg, opts->flags, ctx->qnm, gi, recs_all, nullptr, ap, outp->status);
            GWF_CUDA(cudaGetLastError());
        } else if (outp->status) GWF_CUDA(cudaMemsetAsync(outp->status, 0, sizeof(int) * (size_t)n_total, st));   // the input bits are the prologue's
    }
    if (!(range.phases & 2) || n == 0) return GWF_OK;
    const int sms = ctx->sms;
    // dynamic shared memory: per-warp staging blocks, then the PSD windows (as many tables as fit in 227 KB)
    const size_t ws_bytes_smem = sizeof(FisherSmem<Rec, typename PointFns<MODEL, NT>::Extra>) * kWarpsPerCta + (pair ? kXferBytes : 0);
    size_t shmem_pass = plan_psd_cache(net, ws_bytes_smem, kSmemLimit);
    // the unrolled, shape-specialised form (measured per 1e4 events: IMRPhenomD ET+2CE 1.55 -> 1.06 ms, NRTidalv2 3.73 -> 2.26 ms,
    // TaylorF2 ETSL 0.66 -> 0.55 ms; without the compile-time shape TaylorF2's unrolled loop spilled and lost 7-20 %)
    constexpr bool kHasFast = PointFns<MODEL, NT>::kHasFast;
    typedef void (*Kern)(const Rec*, const EventAux*, EventsDev, long long, int, int, ModelCfg, const NetworkDev, const FisherOut, int);
    // IMRPhenomHM needs extra accumulators for the SNR derivatives (its own instantiation); the other models rebuild them
    // from the compact Gram whenever the output pointer is given
    constexpr bool kSdKernel = MODEL == kPhenomHM;
    const Kern k0 = (kSdKernel && snr_derivs) ? fisher_kernel<MODEL, NT, 0, kSdKernel> : fisher_kernel<MODEL, NT, 0, false>;
    // kernel for a fast plan and shape (instantiated shapes only; anything else runs the run-time-bounds form)
    auto pick = [&](int fast, int shape) -> Kern {
        if (!kHasFast || fast == 0) return k0;
#define GWF_PICK(S) if (shape == S) return fast == 2 ? (Kern)fisher_kernel<MODEL, NT, kHasFast ? 2 : 0, false, kHasFast ? S : 0> \
                                                     : (Kern)fisher_kernel<MODEL, NT, kHasFast ? 1 : 0, false, kHasFast ? S : 0>;
        GWF_FISHER_SHAPES(GWF_PICK)
#undef GWF_PICK
        return fast == 2 ? (Kern)fisher_kernel<MODEL, NT, kHasFast ? 2 : 0, false> : (Kern)fisher_kernel<MODEL, NT, kHasFast ? 1 : 0, false>;
    };
    const bool allow_fast = kHasFast && !(opts->flags & GWF_OPT_GENERIC_LOOP);
    int fast = allow_fast ? plan_fast(net) : 0;
    const int per_cta = pair ? kWarpsPerCta / 2 : kWarpsPerCta;
    const long long want = (n + per_cta - 1) / per_cta;
    const unsigned grid = (unsigned)std::min<long long>(want, (long long)sms);
    const int npass = opts->per_arm ? gwf_num_arms(dets, ndet) : 1;
    for (int pass = 0; pass < npass; ++pass) {
        if (opts->per_arm) {
            rc = build_network(dets, ndet, pd, npsd, pass, false, net);
            if (rc) return rc;
            // a table that found no room next to the others in the all-arms plan can fit on its own in a per-arm pass:
            // the launch is sized for THIS pass's plan
            shmem_pass = plan_psd_cache(net, ws_bytes_smem, kSmemLimit);
            fast = allow_fast ? plan_fast(net) : 0;
        }
        FisherOut fo;
        fo.fisher = fisher + (size_t)pass * n * NPACK;
        fo.snr2 = snr2 ? snr2 + (size_t)pass * n : nullptr;
        fo.snr2_integ = outp->snr2_integ ? outp->snr2_integ + (size_t)pass * n : nullptr;
        fo.snr_derivs = snr_derivs ? snr_derivs + (size_t)pass * n * NP : nullptr;
        fo.status = outp->status;
        fo.peers.n = 0;
        for (int q = 0; q < kMaxPeers; ++q) fo.peers.p[q] = nullptr;
        if (outp->npeers > 0 && !opts->per_arm) {
            if (outp->npeers > kMaxPeers || !outp->peer_fisher) return fail(GWF_ERR_ARG, "gwf_fisher_out: at most 8 peer slots");
            fo.peers.n = outp->npeers;
            for (int q = 0; q < outp->npeers; ++q) {
                if (!outp->peer_fisher[q]) return fail(GWF_ERR_ARG, "gwf_fisher_out: null peer slot");
                fo.peers.p[q] = outp->peer_fisher[q];
            }
        }
        if constexpr (kSplitPair) {
            if (pair && !(opts->flags & GWF_OPT_HM_BLOCK_PAIRS)) {
                typedef WarpSmem<Rec, HMExtra> WSs;
                const size_t base = sizeof(WSs) * (kWarpsPerCta / 2) + sizeof(double) * (kWarpsPerCta / 2) * 2 * HMStrainSink<NT>::kWords * 32;
                const size_t shmem_split = plan_psd_cache(net, base, kSmemLimit);
                auto ks = snr_derivs ? fisher_hm_split_kernel<NT, true> : fisher_hm_split_kernel<NT, false>;
                if (int rcs = ensure_smem(*ctx, ks, shmem_split)) return rcs;
                ks<<<grid, kFisherThreads, shmem_split, st>>>(recs, aux, ev, n, opts->res, lin, cfg, net, fo);
                GWF_CUDA(cudaGetLastError());
                continue;
            }
        }
        const Kern kern = pick(fast, fast_shape(net));
        if (int rcs = ensure_smem(*ctx, kern, shmem_pass)) return rcs;
        kern<<<grid, kFisherThreads, shmem_pass, st>>>(recs, aux, ev, n, opts->res, lin, cfg, net, fo, pair);
        GWF_CUDA(cudaGetLastError());
    }
    return GWF_OK;
}

template <int MODEL, int NT>
static int run_derivs(const gwf_model* model, const gwf_detector* dets, int ndet, const gwf_psd* const* psds, int npsd, const EventsDev& ev,
                      long long n, const gwf_opts* opts, double* derivs, void* ws, size_t ws_bytes, cudaStream_t st) {
    typedef typename ModelTraits<MODEL, NT>::Rec Rec;
    constexpr int NP = NT + 7;
    if (ws_bytes < sizeof(Rec) * (size_t)n) return fail(GWF_ERR_WORKSPACE, "workspace too small");
    Rec* recs = reinterpret_cast<Rec*>(ws);
    ModelCfg cfg = {model->id, model->flags, model->fcutPar, model->fRef};
    DeviceCtx* ctx = nullptr;
    if (int rc0 = device_ctx(model_tables<MODEL>(), &ctx)) return rc0;
    NetworkDev net;
    PsdDev pd[kMaxPsd];
    int rc = collect_psds(psds, npsd, pd);
    if (rc) return rc;
    rc = build_network(dets, ndet, pd, npsd, -1, false, net);
    if (rc) return rc;
    GroupInfo gi = group_info(net);
    const int pb = 128;
    prologue_kernel<MODEL, NT><<<(unsigned)((n + pb - 1) / pb), pb, 0, st>>>(ev, n, cfg, opts->flags, ctx->qnm, gi, recs);
    GWF_CUDA(cudaGetLastError());
    const int sms = ctx->sms;
    const size_t shmem = sizeof(WarpSmem<Rec, typename PointFns<MODEL, NT>::Extra>) * kWarpsPerCta;
    auto kern = derivs_kernel<MODEL, NT>;
    if (int rcs = ensure_smem(*ctx, kern, shmem)) return rcs;
    const long long want = (n + kWarpsPerCta - 1) / kWarpsPerCta;
    const unsigned grid = (unsigned)std::min<long long>(want, (long long)sms * 2);
    const int lin = (opts->flags & GWF_OPT_LIN_GRID) ? 1 : 0;
    const int npass = gwf_num_arms(dets, ndet);
    for (int pass = 0; pass < npass; ++pass) {
        rc = build_network(dets, ndet, pd, npsd, pass, false, net);
        if (rc) return rc;
        for (int i = 0; i < net.npsd; ++i) net.psd[i].c_off = -1;
        kern<<<grid, kFisherThreads, shmem, st>>>(recs, ev, n, opts->res, lin, cfg, net,
                                                  reinterpret_cast<double2*>(derivs) + (size_t)pass * NP * n * opts->res);
        GWF_CUDA(cudaGetLastError());
    }
    return GWF_OK;
}

template <int MODEL>
static int run_strain(const gwf_model* model, const gwf_detector* dets, int ndet, const gwf_psd* const* psds, int npsd, const EventsDev& ev,
                      long long n, const gwf_opts* opts, double* strain, void* ws, size_t ws_bytes, cudaStream_t st) {
    typedef typename ModelTraits<MODEL, 4>::Rec Rec;
    if (ws_bytes < sizeof(Rec) * (size_t)n) return fail(GWF_ERR_WORKSPACE, "workspace too small");
    Rec* recs = reinterpret_cast<Rec*>(ws);
    ModelCfg cfg = {model->id, model->flags, model->fcutPar, model->fRef};
    DeviceCtx* ctx = nullptr;
    if (int rc0 = device_ctx(model_tables<MODEL>(), &ctx)) return rc0;
    NetworkDev net;
    PsdDev pd[kMaxPsd];
    int rc = collect_psds(psds, npsd, pd);
    if (rc) return rc;
    rc = build_network(dets, ndet, pd, npsd, -1, false, net);
    if (rc) return rc;
    GroupInfo gi = group_info(net);
    const int pb = 128;
    // GWstrain is handed the dict entries as they are (no Fisher re-parametrisation, signal.py:1871)
    prologue_kernel<MODEL, 4><<<(unsigned)((n + pb - 1) / pb), pb, 0, st>>>(ev, n, cfg, 0, ctx->qnm, gi, recs);
    GWF_CUDA(cudaGetLastError());
    const int sms = ctx->sms;
    const size_t shmem = sizeof(WarpSmem<Rec, typename PointFns<MODEL, 4>::Extra>) * kWarpsPerCta;
    auto kern = strain_kernel<MODEL>;
    if (int rcs = ensure_smem(*ctx, kern, shmem)) return rcs;
    const long long want = (n + kWarpsPerCta - 1) / kWarpsPerCta;
    const unsigned grid = (unsigned)std::min<long long>(want, (long long)sms * 2);
    const int lin = (opts->flags & GWF_OPT_LIN_GRID) ? 1 : 0;
    const int npass = gwf_num_arms(dets, ndet);
    for (int pass = 0; pass < npass; ++pass) {
        rc = build_network(dets, ndet, pd, npsd, pass, false, net);
        if (rc) return rc;
        for (int i = 0; i < net.npsd; ++i) net.psd[i].c_off = -1;
        kern<<<grid, kFisherThreads, shmem, st>>>(recs, ev, n, opts->res, lin, cfg, net, reinterpret_cast<double2*>(strain) + (size_t)pass * n * opts->res);
        GWF_CUDA(cudaGetLastError());
    }
    return GWF_OK;
}

template <int MODEL>
static int run_snr(const gwf_model* model, const gwf_detector* dets, int ndet, const gwf_psd* const* psds, int npsd, const EventsDev& ev,
                   long long n, const gwf_opts* opts, double* snr2_arm, void* ws, size_t ws_bytes, cudaStream_t st) {
    typedef typename ModelTraits<MODEL, 4>::Rec Rec;
    const size_t rec_bytes = (sizeof(Rec) * (size_t)n + 15) & ~(size_t)15;
    if (ws_bytes < rec_bytes + sizeof(EventAux) * (size_t)n) return fail(GWF_ERR_WORKSPACE, "workspace too small");
    Rec* recs = reinterpret_cast<Rec*>(ws);
    EventAux* aux = reinterpret_cast<EventAux*>(reinterpret_cast<char*>(ws) + rec_bytes);
    ModelCfg cfg = {model->id, model->flags, model->fcutPar, model->fRef};
    DeviceCtx* ctx = nullptr;
    if (int rc0 = device_ctx(model_tables<MODEL>(), &ctx)) return rc0;
    NetworkDev net;
    PsdDev pd[kMaxPsd];
    int rc = collect_psds(psds, npsd, pd);
    if (rc) return rc;
    rc = build_network(dets, ndet, pd, npsd, -1, true, net);
    if (rc) return rc;
    GroupInfo gi = group_info(net);
    const int pb = 128;
    // SNRInteg hands the dict entries straight to the waveform: no Fisher re-parametrisation (signal.py:715-726)
    const int lin = (opts->flags & GWF_OPT_LIN_GRID) ? 1 : 0;
    AuxPlan ap;
    ap.out = aux;
    for (int g = 0; g < kMaxGroups; ++g) ap.fmax[g] = net.group_fmax[g];
    constexpr int kSnrSplit = SnrMap<MODEL>::kSplit, kSnrGroups = SnrMap<MODEL>::kGroups;
    ap.res = opts->res; ap.lin = lin; ap.stride = 32 * kSnrSplit;
    prologue_kernel<MODEL, 4><<<dim3((unsigned)((n + pb - 1) / pb), (MODEL == kPhenomD || MODEL == kNRTidalv2) ? 2 : 1), pb, 0, st>>>(ev, n, cfg, 0, ctx->qnm, gi, recs, nullptr, ap);
    GWF_CUDA(cudaGetLastError());
    const int sms = ctx->sms;
    const size_t base = sizeof(WarpSmem<Rec, typename PointFns<MODEL, 4>::Extra>) * kSnrGroups + sizeof(double) * net.narms * (32 * kSnrWarps + kSnrWarps);
    const size_t shmem = plan_psd_cache(net, base, kSmemLimit);
    constexpr bool kHasFast = PointFns<MODEL, 4>::kHasFast;
    typedef void (*Kern)(const Rec*, const EventAux*, EventsDev, long long, int, int, ModelCfg, const NetworkDev, int, double*);
    const int fast = (kHasFast && !(opts->flags & GWF_OPT_GENERIC_LOOP)) ? plan_fast(net) : 0;
    auto pick = [&](int shape) -> Kern {
        if (!kHasFast || fast == 0) return snr_kernel<MODEL, 0>;
#define GWF_PICK(S) if (shape == S) return fast == 2 ? (Kern)snr_kernel<MODEL, kHasFast ? 2 : 0, kHasFast ? S : 0> : (Kern)snr_kernel<MODEL, kHasFast ? 1 : 0, kHasFast ? S : 0>;
        GWF_SNR_SHAPES(GWF_PICK)
#undef GWF_PICK
        return fast == 2 ? (Kern)snr_kernel<MODEL, kHasFast ? 2 : 0> : (Kern)snr_kernel<MODEL, kHasFast ? 1 : 0>;
    };
    const Kern kern = pick(fast_shape(net));
    if (int rcs = ensure_smem(*ctx, kern, shmem)) return rcs;
    const long long want = (n + kSnrGroups - 1) / kSnrGroups;
    const unsigned grid = (unsigned)std::min<long long>(want, (long long)sms);
    kern<<<grid, kSnrThreads, shmem, st>>>(recs, aux, ev, n, opts->res, lin, cfg, net, net.narms, snr2_arm);
    GWF_CUDA(cudaGetLastError());
    return GWF_OK;
}

// DetDev / ArmDev of one gwf_detector at the arm orientation xax + rot_deg (signal.py:342-387: rot in degrees)
static void single_arm(const gwf_detector& d, double rot_deg, DetDev& o, ArmDev& a) {
    std::memset(&o, 0, sizeof(o));
    o.sl = std::sin(d.lat_rad); o.cl = std::cos(d.lat_rad);
    o.s2l = std::sin(2.0 * d.lat_rad); o.c2l = std::cos(2.0 * d.lat_rad);
    o.slon = std::sin(d.long_rad); o.clon = std::cos(d.long_rad);
    o.fmin = d.fmin; o.fmax = d.fmax > 0.0 ? d.fmax : 0.0;
    o.no_motion = d.no_motion != 0;
    o.use_rot = (d.use_earth_motion != 0) && !o.no_motion;
    const double sarm = d.shape == 0 ? 1.0 : std::sin(kPi / 3.);
    const double x = d.xax_rad + rot_deg * kPi / 180.;
    a.S2 = sarm * std::sin(2.0 * x); a.C2 = sarm * std::cos(2.0 * x); a.weight = 1.0; a.out = 0; a.pad = 0;
}

template <int MODEL>
static int run_signal_grid(const gwf_model* model, const gwf_detector* det, double rot_deg, const EventsDev& ev, long long n, const double* f, int res,
                           int f2d, const SignalOut& out, void* ws, size_t ws_bytes, cudaStream_t st) {
    typedef typename ModelTraits<MODEL, 4>::Rec Rec;
    const size_t rec_bytes = (sizeof(Rec) * (size_t)n + 15) & ~(size_t)15;
    if (ws_bytes < rec_bytes + 2 * sizeof(double) * (size_t)n) return fail(GWF_ERR_WORKSPACE, "workspace too small");
    Rec* recs = reinterpret_cast<Rec*>(ws);
    double* fmin_ev = reinterpret_cast<double*>(reinterpret_cast<char*>(ws) + rec_bytes);
    ModelCfg cfg = {model->id, model->flags, model->fcutPar, model->fRef};
    DeviceCtx* ctx = nullptr;
    if (int rc0 = device_ctx(model_tables<MODEL>(), &ctx)) return rc0;
    GroupInfo gi = group_info_user_grid();
    const int pb = 128;
    const unsigned pg = (unsigned)((n + pb - 1) / pb);
    grid_min_kernel<<<pg, pb, 0, st>>>(f, res, n, f2d, fmin_ev);                 // fRef = min of the user's grid (waveforms.py:1139)
    GWF_CUDA(cudaGetLastError());
    prologue_kernel<MODEL, 4><<<pg, pb, 0, st>>>(ev, n, cfg, 0, ctx->qnm, gi, recs, fmin_ev);
    GWF_CUDA(cudaGetLastError());
    DetDev d;
    ArmDev a;
    single_arm(*det, rot_deg, d, a);
    const size_t shmem = sizeof(WaveBlk<Rec>) * kWarpsPerCta;
    auto kern = signal_grid_kernel<MODEL>;
    if (int rcs = ensure_smem(*ctx, kern, shmem)) return rcs;
    const long long want = (n + kWarpsPerCta - 1) / kWarpsPerCta;
    const unsigned grid = (unsigned)std::min<long long>(want, (long long)ctx->sms * 4);
    kern<<<grid, kFisherThreads, shmem, st>>>(recs, ev, n, f, res, f2d, cfg, d, a, out);
    GWF_CUDA(cudaGetLastError());
    return GWF_OK;
}

template <int MODEL>
static int run_waveform(const gwf_model* model, const EventsDev& ev, long long n, const double* f, int res, int f2d, double* phi, double* ampl,
                        double* tau, double* hphc, double* fcut, void* ws, size_t ws_bytes, cudaStream_t st) {
    typedef typename ModelTraits<MODEL, 4>::Rec Rec;
    const size_t rec_bytes = (sizeof(Rec) * (size_t)n + 15) & ~(size_t)15;
    if (ws_bytes < rec_bytes + 2 * sizeof(double) * (size_t)n) return fail(GWF_ERR_WORKSPACE, "workspace too small");
    Rec* recs = reinterpret_cast<Rec*>(ws);
    double* fmin_ev = reinterpret_cast<double*>(reinterpret_cast<char*>(ws) + rec_bytes);
    ModelCfg cfg = {model->id, model->flags, model->fcutPar, model->fRef};
    DeviceCtx* ctx = nullptr;
    if (int rc0 = device_ctx(model_tables<MODEL>(), &ctx)) return rc0;
    GroupInfo gi = group_info_user_grid();
    const int pb = 128;
    const unsigned pg = (unsigned)((n + pb - 1) / pb);
    const double* fmin_arg = nullptr;
    if (res > 0) {
        grid_min_kernel<<<pg, pb, 0, st>>>(f, res, n, f2d, fmin_ev);
        GWF_CUDA(cudaGetLastError());
        fmin_arg = fmin_ev;
    }
    prologue_kernel<MODEL, 4><<<pg, pb, 0, st>>>(ev, n, cfg, 0, ctx->qnm, gi, recs, fmin_arg);
    GWF_CUDA(cudaGetLastError());
    const int sms = ctx->sms;
    const size_t shmem = sizeof(WaveBlk<Rec>) * kWarpsPerCta;
    auto kern = waveform_kernel<MODEL>;
    if (int rcs = ensure_smem(*ctx, kern, shmem)) return rcs;
    const long long want = (n + kWarpsPerCta - 1) / kWarpsPerCta;
    const unsigned grid = (unsigned)std::min<long long>(want, (long long)sms * 4);
    kern<<<grid, kFisherThreads, shmem, st>>>(recs, ev, n, f, res, f2d, cfg, phi, ampl, tau, hphc, fcut);
    GWF_CUDA(cudaGetLastError());
    return GWF_OK;
}

}  // namespace gwf

using namespace gwf;

extern "C" {

int gwf_version(void) { return GWF_VERSION; }
const char* gwf_last_error(void) { return g_err.c_str(); }

int gwf_num_params(const gwf_model* model) {
    if (!model) return GWF_ERR_ARG;
    const int nt = model_nt(*model);
    return nt < 0 ? GWF_ERR_ARG : nt + 7;
}

int gwf_num_arms(const gwf_detector* dets, int32_t ndet) {
    int n = 0;
    for (int i = 0; i < ndet; ++i) n += dets[i].shape == 0 ? 1 : 3;
    return n;
}

size_t gwf_workspace_bytes(const gwf_model* model, int64_t n) {
    if (!model || n < 0) return 0;
    size_t rec = 0;
    switch (model->id) {
        case GWF_TAYLORF2: rec = sizeof(TF2Rec<7>); break;
        case GWF_IMRPHENOMD: rec = sizeof(PhenomDRec<4>); break;
        case GWF_IMRPHENOMD_NRTIDALV2: rec = std::max(sizeof(NRTidalRec<4>), sizeof(NRTidalRec<6>)); break;
        case GWF_IMRPHENOMHM: rec = sizeof(HMRec<4>); break;
        case GWF_IMRPHENOMNSBH: rec = std::max(sizeof(NSBHRec<4>), sizeof(NSBHRec<6>)); break;
        default: rec = 0;
    }
    // records, then the larger of the waveform path's per-event grid minimum and the Fisher path's EventAux
    return ((rec * (size_t)n + 15) & ~(size_t)15) + std::max(2 * sizeof(double), sizeof(EventAux)) * (size_t)n;
}

int gwf_psd_create(const double* f, const double* S, int32_t n, gwf_psd** out) {
    if (!out) return fail(GWF_ERR_ARG, "gwf_psd_create: null output");
    PsdHost* h = new PsdHost;
    int rc = build_psd_tables(f, S, n, h->tab_h, h->bucket_h, h->dev);
    if (rc) { delete h; return rc; }
    *out = reinterpret_cast<gwf_psd*>(h);     // device copies are made by the first call that uses the handle on a device
    return GWF_OK;
}

void gwf_psd_destroy(gwf_psd* psd) {
    if (!psd) return;
    PsdHost* h = reinterpret_cast<PsdHost*>(psd);
    int cur = 0;
    cudaGetDevice(&cur);
    for (int d = 0; d < kMaxDevices; ++d)
        if (h->tab[d]) {
            cudaSetDevice(d);
            cudaFree(h->tab[d]);
            cudaFree(h->bucket[d]);
        }
    cudaSetDevice(cur);
    delete h;
}

int gwf_xitide_table(double* table_host) {
    DeviceCtx* ctx = nullptr;
    if (int rc0 = device_ctx(2, &ctx)) return rc0;
    if (table_host)
        GWF_CUDA(cudaMemcpy(table_host, ctx->qnm.xitide, sizeof(double) * (size_t)kXiRes * kXiRes * kXiRes, cudaMemcpyDeviceToHost));
    return GWF_OK;
}

int gwf_set_qnm_tables(const double* a, const double* fring, const double* fdamp, int32_t n) {
    if (!a || !fring || !fdamp || n < 2) return fail(GWF_ERR_ARG, "gwf_set_qnm_tables: bad arguments");
    std::lock_guard<std::mutex> lock(g_mu);
    g_qnm_host.resize(3 * (size_t)n);
    std::memcpy(g_qnm_host.data(), a, sizeof(double) * n);
    std::memcpy(g_qnm_host.data() + n, fring, sizeof(double) * n);
    std::memcpy(g_qnm_host.data() + 2 * n, fdamp, sizeof(double) * n);
    g_qnm_n = n;
    ++g_qnm_version;                           // every device re-uploads on its next call
    return GWF_OK;
}

static int check_common(const gwf_model* model, const gwf_detector* dets, const gwf_psd* const* psds, const gwf_events* events, int64_t n,
                        const gwf_opts* opts) {
    if (!model || !dets || !psds || !events || !opts) return fail(GWF_ERR_ARG, "null argument");
    if (n < 0) return fail(GWF_ERR_ARG, "negative event count");
    if (opts->res < 2) return fail(GWF_ERR_ARG, "res must be at least 2");
    for (int i = 0; i < 11; ++i)
        if (!events->p[i] && n > 0) return fail(GWF_ERR_ARG, "missing event parameter array");
    if ((model->id == GWF_IMRPHENOMD || model->id == GWF_IMRPHENOMD_NRTIDALV2 || model->id == GWF_IMRPHENOMNSBH) && g_qnm_n == 0)
        return fail(GWF_ERR_ARG, "QNM tables not set (gwf_set_qnm_tables)");
    return GWF_OK;
}

static int fisher_dispatch(const gwf_model* model, const gwf_detector* dets, int32_t ndet, const gwf_psd* const* psds, int32_t npsd, const gwf_events* events,
                           int64_t n, const FisherRange& range, const gwf_opts* opts, const gwf_fisher_out* out, void* workspace, size_t workspace_bytes,
                           void* stream);

int gwf_fisher_ex(const gwf_model* model, const gwf_detector* dets, int32_t ndet, const gwf_psd* const* psds, int32_t npsd, const gwf_events* events,
                  int64_t n, const gwf_opts* opts, const gwf_fisher_out* out, void* workspace, size_t workspace_bytes, void* stream) {
    const FisherRange all = {0, n, 3};
    if (!out || !out->fisher_packed) return fail(GWF_ERR_ARG, "null output");
    return fisher_dispatch(model, dets, ndet, psds, npsd, events, n, all, opts, out, workspace, workspace_bytes, stream);
}

int gwf_fisher_range(const gwf_model* model, const gwf_detector* dets, int32_t ndet, const gwf_psd* const* psds, int32_t npsd, const gwf_events* events,
                     int64_t n_total, int64_t lo, int64_t m, int32_t phases, const gwf_opts* opts, const gwf_fisher_out* out, void* workspace,
                     size_t workspace_bytes, void* stream) {
    if (lo < 0 || m < 0 || lo + m > n_total || !(phases & 3)) return fail(GWF_ERR_ARG, "gwf_fisher_range: bad range or phases");
    if (!out || ((phases & 2) && !out->fisher_packed)) return fail(GWF_ERR_ARG, "null output");
    const FisherRange r = {lo, m, phases & 3};
    return fisher_dispatch(model, dets, ndet, psds, npsd, events, n_total, r, opts, out, workspace, workspace_bytes, stream);
}

static int fisher_dispatch(const gwf_model* model, const gwf_detector* dets, int32_t ndet, const gwf_psd* const* psds, int32_t npsd, const gwf_events* events,
                           int64_t n, const FisherRange& range, const gwf_opts* opts, const gwf_fisher_out* out, void* workspace, size_t workspace_bytes,
                           void* stream) {
    int rc = check_common(model, dets, psds, events, n, opts);
    if (rc) return rc;
    if (n == 0) return GWF_OK;
    EventsDev ev;
    for (int i = 0; i < GWF_NPARAM_IN; ++i) ev.p[i] = events->p[i];
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    switch (model->id) {
        case GWF_TAYLORF2: {
            const bool ecc = (model->flags & GWF_MODEL_ECCENTRIC) != 0;
            if (ecc && !ev.p[15]) return fail(GWF_ERR_ARG, "eccentric model needs ecc");
            if (model->flags & GWF_MODEL_TIDAL) {
                if (!ev.p[11] || !ev.p[12]) return fail(GWF_ERR_ARG, "tidal model needs Lambda1, Lambda2");
                if (ecc) return run_fisher<kTaylorF2, 7>(model, dets, ndet, psds, npsd, ev, n, range, opts, out, workspace, workspace_bytes, st);
                return run_fisher<kTaylorF2, 6>(model, dets, ndet, psds, npsd, ev, n, range, opts, out, workspace, workspace_bytes, st);
            }
            if (ecc) return run_fisher<kTaylorF2, 5>(model, dets, ndet, psds, npsd, ev, n, range, opts, out, workspace, workspace_bytes, st);
            return run_fisher<kTaylorF2, 4>(model, dets, ndet, psds, npsd, ev, n, range, opts, out, workspace, workspace_bytes, st);
        }
        case GWF_IMRPHENOMD:
            return run_fisher<kPhenomD, 4>(model, dets, ndet, psds, npsd, ev, n, range, opts, out, workspace, workspace_bytes, st);
        case GWF_IMRPHENOMD_NRTIDALV2:
            if (!ev.p[11] || !ev.p[12]) return fail(GWF_ERR_ARG, "tidal model needs Lambda1, Lambda2");
            return run_fisher<kNRTidalv2, 6>(model, dets, ndet, psds, npsd, ev, n, range, opts, out, workspace, workspace_bytes, st);
        case GWF_IMRPHENOMHM:
            return run_fisher<kPhenomHM, 4>(model, dets, ndet, psds, npsd, ev, n, range, opts, out, workspace, workspace_bytes, st);
        case GWF_IMRPHENOMNSBH:
            if (!ev.p[11] || !ev.p[12]) return fail(GWF_ERR_ARG, "tidal model needs Lambda1, Lambda2");
            return run_fisher<kNSBH, 6>(model, dets, ndet, psds, npsd, ev, n, range, opts, out, workspace, workspace_bytes, st);
        default:
            return fail(GWF_ERR_UNSUPPORTED, "model not built yet");
    }
}

int64_t gwf_round_events(const gwf_model* model) {
    if (!model) return 0;
    DeviceCtx* ctx = nullptr;
    if (device_ctx(false, &ctx)) return 0;
    return (int64_t)ctx->sms * (model->id == GWF_TAYLORF2 ? kWarpsPerCta : kWarpsPerCta / 2);
}

int gwf_fisher(const gwf_model* model, const gwf_detector* dets, int32_t ndet, const gwf_psd* const* psds, int32_t npsd, const gwf_events* events,
               int64_t n, const gwf_opts* opts, double* fisher_packed, double* snr2, void* workspace, size_t workspace_bytes, void* stream) {
    gwf_fisher_out out;
    std::memset(&out, 0, sizeof(out));
    out.fisher_packed = fisher_packed;
    out.snr2 = snr2;
    return gwf_fisher_ex(model, dets, ndet, psds, npsd, events, n, opts, &out, workspace, workspace_bytes, stream);
}

int gwf_strain_derivs(const gwf_model* model, const gwf_detector* dets, int32_t ndet, const gwf_psd* const* psds, int32_t npsd, const gwf_events* events,
                      int64_t n, const gwf_opts* opts, double* derivs, void* workspace, size_t workspace_bytes, void* stream) {
    int rc = check_common(model, dets, psds, events, n, opts);
    if (rc) return rc;
    if (!derivs) return fail(GWF_ERR_ARG, "null output");
    if (n == 0) return GWF_OK;
    EventsDev ev;
    for (int i = 0; i < GWF_NPARAM_IN; ++i) ev.p[i] = events->p[i];
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    switch (model->id) {
        case GWF_TAYLORF2: {
            const bool ecc = (model->flags & GWF_MODEL_ECCENTRIC) != 0;
            if (ecc && !ev.p[15]) return fail(GWF_ERR_ARG, "eccentric model needs ecc");
            if (model->flags & GWF_MODEL_TIDAL) {
                if (!ev.p[11] || !ev.p[12]) return fail(GWF_ERR_ARG, "tidal model needs Lambda1, Lambda2");
                if (ecc) return run_derivs<kTaylorF2, 7>(model, dets, ndet, psds, npsd, ev, n, opts, derivs, workspace, workspace_bytes, st);
                return run_derivs<kTaylorF2, 6>(model, dets, ndet, psds, npsd, ev, n, opts, derivs, workspace, workspace_bytes, st);
            }
            if (ecc) return run_derivs<kTaylorF2, 5>(model, dets, ndet, psds, npsd, ev, n, opts, derivs, workspace, workspace_bytes, st);
            return run_derivs<kTaylorF2, 4>(model, dets, ndet, psds, npsd, ev, n, opts, derivs, workspace, workspace_bytes, st);
        }
        case GWF_IMRPHENOMD:
            return run_derivs<kPhenomD, 4>(model, dets, ndet, psds, npsd, ev, n, opts, derivs, workspace, workspace_bytes, st);
        case GWF_IMRPHENOMD_NRTIDALV2:
            if (!ev.p[11] || !ev.p[12]) return fail(GWF_ERR_ARG, "tidal model needs Lambda1, Lambda2");
            return run_derivs<kNRTidalv2, 6>(model, dets, ndet, psds, npsd, ev, n, opts, derivs, workspace, workspace_bytes, st);
        case GWF_IMRPHENOMHM:
            return run_derivs<kPhenomHM, 4>(model, dets, ndet, psds, npsd, ev, n, opts, derivs, workspace, workspace_bytes, st);
        case GWF_IMRPHENOMNSBH:
            if (!ev.p[11] || !ev.p[12]) return fail(GWF_ERR_ARG, "tidal model needs Lambda1, Lambda2");
            return run_derivs<kNSBH, 6>(model, dets, ndet, psds, npsd, ev, n, opts, derivs, workspace, workspace_bytes, st);
        default:
            return fail(GWF_ERR_UNSUPPORTED, "gwf_strain_derivs: model not built");
    }
}

int gwf_snr(const gwf_model* model, const gwf_detector* dets, int32_t ndet, const gwf_psd* const* psds, int32_t npsd, const gwf_events* events,
            int64_t n, const gwf_opts* opts, double* snr2_arm, void* workspace, size_t workspace_bytes, void* stream) {
    int rc = check_common(model, dets, psds, events, n, opts);
    if (rc) return rc;
    if (!snr2_arm) return fail(GWF_ERR_ARG, "null output");
    if (n == 0) return GWF_OK;
    EventsDev ev;
    for (int i = 0; i < GWF_NPARAM_IN; ++i) ev.p[i] = events->p[i];
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    switch (model->id) {
        case GWF_TAYLORF2:
            return run_snr<kTaylorF2>(model, dets, ndet, psds, npsd, ev, n, opts, snr2_arm, workspace, workspace_bytes, st);
        case GWF_IMRPHENOMD:
            return run_snr<kPhenomD>(model, dets, ndet, psds, npsd, ev, n, opts, snr2_arm, workspace, workspace_bytes, st);
        case GWF_IMRPHENOMD_NRTIDALV2:
            if (!ev.p[11] || !ev.p[12]) return fail(GWF_ERR_ARG, "tidal model needs Lambda1, Lambda2");
            return run_snr<kNRTidalv2>(model, dets, ndet, psds, npsd, ev, n, opts, snr2_arm, workspace, workspace_bytes, st);
        case GWF_IMRPHENOMHM:
            return run_snr<kPhenomHM>(model, dets, ndet, psds, npsd, ev, n, opts, snr2_arm, workspace, workspace_bytes, st);
        case GWF_IMRPHENOMNSBH:
            return run_snr<kNSBH>(model, dets, ndet, psds, npsd, ev, n, opts, snr2_arm, workspace, workspace_bytes, st);
        default:
            return fail(GWF_ERR_UNSUPPORTED, "model not built yet");
    }
}

int gwf_strain(const gwf_model* model, const gwf_detector* dets, int32_t ndet, const gwf_psd* const* psds, int32_t npsd, const gwf_events* events,
               int64_t n, const gwf_opts* opts, double* strain, void* workspace, size_t workspace_bytes, void* stream) {
    int rc = check_common(model, dets, psds, events, n, opts);
    if (rc) return rc;
    if (!strain) return fail(GWF_ERR_ARG, "null output");
    if (n == 0) return GWF_OK;
    EventsDev ev;
    for (int i = 0; i < GWF_NPARAM_IN; ++i) ev.p[i] = events->p[i];
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    switch (model->id) {
        case GWF_TAYLORF2:
            if ((model->flags & GWF_MODEL_ECCENTRIC) && !ev.p[15]) return fail(GWF_ERR_ARG, "eccentric model needs ecc");
            if ((model->flags & GWF_MODEL_TIDAL) && (!ev.p[11] || !ev.p[12])) return fail(GWF_ERR_ARG, "tidal model needs Lambda1, Lambda2");
            return run_strain<kTaylorF2>(model, dets, ndet, psds, npsd, ev, n, opts, strain, workspace, workspace_bytes, st);
        case GWF_IMRPHENOMD:
            return run_strain<kPhenomD>(model, dets, ndet, psds, npsd, ev, n, opts, strain, workspace, workspace_bytes, st);
        case GWF_IMRPHENOMD_NRTIDALV2:
            if (!ev.p[11] || !ev.p[12]) return fail(GWF_ERR_ARG, "tidal model needs Lambda1, Lambda2");
            return run_strain<kNRTidalv2>(model, dets, ndet, psds, npsd, ev, n, opts, strain, workspace, workspace_bytes, st);
        case GWF_IMRPHENOMHM:
            return run_strain<kPhenomHM>(model, dets, ndet, psds, npsd, ev, n, opts, strain, workspace, workspace_bytes, st);
        case GWF_IMRPHENOMNSBH:
            return run_strain<kNSBH>(model, dets, ndet, psds, npsd, ev, n, opts, strain, workspace, workspace_bytes, st);
        default:
            return fail(GWF_ERR_UNSUPPORTED, "gwf_strain: model not built");
    }
}

int gwf_overlap(const double* h1, const double* h2, const double* fcut, int64_t n, int32_t res, double fmin, const gwf_psd* psd, double* overlap,
                double* snr2_1, double* snr2_2, void* stream) {
    if (!h1 || !h2 || !fcut || !psd || !overlap || !snr2_1 || !snr2_2) return fail(GWF_ERR_ARG, "gwf_overlap: null argument");
    if (n < 0 || res < 2 || !(fmin > 0.0)) return fail(GWF_ERR_ARG, "gwf_overlap: bad grid");
    if (n == 0) return GWF_OK;
    PsdDev pd;                                  // the handle's table on the CURRENT device (uploaded on first use)
    if (int rc = collect_psds(&psd, 1, &pd)) return rc;
    pd.c_off = -1;
    const int tb = 256;
    const long long threads = (long long)n * 32;
    overlap_kernel<<<(unsigned)((threads + tb - 1) / tb), tb, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const double2*>(h1), reinterpret_cast<const double2*>(h2), fcut, n, res, fmin, pd, overlap, snr2_1, snr2_2);
    GWF_CUDA(cudaGetLastError());
    return GWF_OK;
}

int gwf_unpack_fisher_ld(const double* packed, int64_t n, int32_t nP, double* full, int64_t ld, void* stream) {
    if (!packed || !full || nP < 1 || nP > 14 || ld < n) return fail(GWF_ERR_ARG, "gwf_unpack_fisher: bad arguments");
    if (n == 0) return GWF_OK;
    unpack_kernel<<<(unsigned)((n + 31) / 32), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(packed, n, nP, full, ld);
    GWF_CUDA(cudaGetLastError());
    return GWF_OK;
}

int gwf_peer_alloc(size_t bytes, void** ptr_out, unsigned char* handle_out) {
    if (!ptr_out || !handle_out || bytes == 0) return fail(GWF_ERR_ARG, "gwf_peer_alloc: bad arguments");
    void* p = nullptr;
    GWF_CUDA(cudaMalloc(&p, bytes));                 // its own allocation: the IPC handle names exactly this buffer
    GWF_CUDA(cudaMemset(p, 0, bytes));
    cudaIpcMemHandle_t h;
    const cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); return fail(GWF_ERR_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e)); }
    static_assert(sizeof(h) == GWF_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t is 64 bytes");
    std::memcpy(handle_out, &h, sizeof(h));
    *ptr_out = p;
    return GWF_OK;
}

int gwf_peer_open(const unsigned char* handle, void** ptr_out) {
    if (!handle || !ptr_out) return fail(GWF_ERR_ARG, "gwf_peer_open: bad arguments");
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof(h));
    // opened with THIS rank's device current: the lazy-peer-access flag maps the exporter's memory for this device's kernels
    GWF_CUDA(cudaIpcOpenMemHandle(ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
    return GWF_OK;
}

int gwf_peer_close(void* ptr) {
    if (ptr) GWF_CUDA(cudaIpcCloseMemHandle(ptr));
    return GWF_OK;
}

int gwf_peer_free(void* ptr) {
    if (ptr) GWF_CUDA(cudaFree(ptr));
    return GWF_OK;
}

int gwf_unpack_gather(const double* packed, int64_t n, int32_t nP, double* full, int64_t ld, double* const* peer_slots, int32_t npeers, void* stream) {
    if (!packed || nP < 1 || nP > 14 || (full && ld < n)) return fail(GWF_ERR_ARG, "gwf_unpack_gather: bad arguments");
    if (npeers < 0 || npeers > kMaxPeers || (npeers > 0 && !peer_slots)) return fail(GWF_ERR_ARG, "gwf_unpack_gather: at most 8 peer slots");
    if (n == 0) return GWF_OK;
    PeerSlots ps;
    ps.n = npeers;
    for (int i = 0; i < kMaxPeers; ++i) ps.p[i] = i < npeers ? peer_slots[i] : nullptr;
    for (int i = 0; i < npeers; ++i)
        if (!ps.p[i]) return fail(GWF_ERR_ARG, "gwf_unpack_gather: null peer slot");
    unpack_gather_kernel<<<(unsigned)((n + 31) / 32), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(packed, n, nP, full, ld, ps);
    GWF_CUDA(cudaGetLastError());
    return GWF_OK;
}

int gwf_unpack_fisher(const double* packed, int64_t n, int32_t nP, double* full, void* stream) {
    return gwf_unpack_fisher_ld(packed, n, nP, full, n, stream);
}

int gwf_copy_2d(void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t width_bytes, size_t height, void* stream) {
    if (!dst || !src) return fail(GWF_ERR_ARG, "gwf_copy_2d: null pointer");
    if (width_bytes == 0 || height == 0) return GWF_OK;
    GWF_CUDA(cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, width_bytes, height, cudaMemcpyDefault, reinterpret_cast<cudaStream_t>(stream)));
    return GWF_OK;
}

int gwf_fp64_peak(double ms, double* tflops_out, void* stream) {
    if (!tflops_out) return fail(GWF_ERR_ARG, "gwf_fp64_peak: null output");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int dev = 0, sms = 0;
    GWF_CUDA(cudaGetDevice(&dev));
    GWF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    double* out = nullptr;
    GWF_CUDA(cudaMalloc(&out, sizeof(double)));
    cudaEvent_t e0, e1;
    GWF_CUDA(cudaEventCreate(&e0));
    GWF_CUDA(cudaEventCreate(&e1));
    const int blocks = sms * 8, threads = 256;
    int iters = 2000;
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        GWF_CUDA(cudaEventRecord(e0, st));
        dfma_kernel<<<blocks, threads, 0, st>>>(out, iters, 1.0000001, 1e-9);
        GWF_CUDA(cudaEventRecord(e1, st));
        GWF_CUDA(cudaEventSynchronize(e1));
        float t = 0.f;
        GWF_CUDA(cudaEventElapsedTime(&t, e0, e1));
        const double flops = 2.0 * 64.0 * iters * (double)blocks * threads;
        if (rep > 0) best = std::max(best, flops / (t * 1e-3) * 1e-12);
        if (t > 0.f) iters = (int)std::min(4.0e6, std::max(1000.0, iters * ms / t));
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    *tflops_out = best;
    return GWF_OK;
}

int gwf_waveform(const gwf_model* model, const gwf_events* events, int64_t n, const double* f, int32_t res, int32_t f_is_2d, double* phi_out,
                 double* ampl_out, double* tau_out, double* hphc_out, double* fcut_out, void* workspace, size_t workspace_bytes, void* stream) {
    if (!model || !events) return fail(GWF_ERR_ARG, "null argument");
    if (n < 0 || res < 0 || (res > 0 && !f)) return fail(GWF_ERR_ARG, "gwf_waveform: bad grid");
    for (int i = 0; i < 11; ++i)
        if (!events->p[i] && n > 0) return fail(GWF_ERR_ARG, "missing event parameter array");
    if ((model->id == GWF_IMRPHENOMD || model->id == GWF_IMRPHENOMD_NRTIDALV2 || model->id == GWF_IMRPHENOMNSBH) && g_qnm_n == 0) return fail(GWF_ERR_ARG, "QNM tables not set (gwf_set_qnm_tables)");
    if (hphc_out && model->id != GWF_IMRPHENOMHM) return fail(GWF_ERR_UNSUPPORTED, "hphc is only defined for IMRPhenomHM");
    if (n == 0) return GWF_OK;
    EventsDev ev;
    for (int i = 0; i < GWF_NPARAM_IN; ++i) ev.p[i] = events->p[i];
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    switch (model->id) {
        case GWF_TAYLORF2: return run_waveform<kTaylorF2>(model, ev, n, f, res, f_is_2d, phi_out, ampl_out, tau_out, hphc_out, fcut_out, workspace, workspace_bytes, st);
        case GWF_IMRPHENOMD: return run_waveform<kPhenomD>(model, ev, n, f, res, f_is_2d, phi_out, ampl_out, tau_out, hphc_out, fcut_out, workspace, workspace_bytes, st);
        case GWF_IMRPHENOMD_NRTIDALV2:
            if (!ev.p[11] || !ev.p[12]) return fail(GWF_ERR_ARG, "tidal model needs Lambda1, Lambda2");
            return run_waveform<kNRTidalv2>(model, ev, n, f, res, f_is_2d, phi_out, ampl_out, tau_out, hphc_out, fcut_out, workspace, workspace_bytes, st);
        case GWF_IMRPHENOMHM: return run_waveform<kPhenomHM>(model, ev, n, f, res, f_is_2d, phi_out, ampl_out, tau_out, hphc_out, fcut_out, workspace, workspace_bytes, st);
        case GWF_IMRPHENOMNSBH: return run_waveform<kNSBH>(model, ev, n, f, res, f_is_2d, phi_out, ampl_out, tau_out, hphc_out, fcut_out, workspace, workspace_bytes, st);
        default: return fail(GWF_ERR_ARG, "unknown model");
    }
}

int gwf_signal_grid(const gwf_model* model, const gwf_detector* det, double rot_deg, const gwf_events* events, int64_t n, const double* f, int32_t res,
                    int32_t f_is_2d, const gwf_signal_out* outp, void* workspace, size_t workspace_bytes, void* stream) {
    if (!model || !det || !events || !outp || !f) return fail(GWF_ERR_ARG, "gwf_signal_grid: null argument");
    if (n < 0 || res < 1) return fail(GWF_ERR_ARG, "gwf_signal_grid: bad grid");
    if (det->shape != 0 && det->shape != 1) return fail(GWF_ERR_ARG, "Enter valid detector configuration");
    for (int i = 0; i < 11; ++i)
        if (!events->p[i] && n > 0) return fail(GWF_ERR_ARG, "missing event parameter array");
    if ((model->id == GWF_IMRPHENOMD || model->id == GWF_IMRPHENOMD_NRTIDALV2 || model->id == GWF_IMRPHENOMNSBH) && g_qnm_n == 0) return fail(GWF_ERR_ARG, "QNM tables not set (gwf_set_qnm_tables)");
    if (outp->psi && model->id == GWF_IMRPHENOMHM) return fail(GWF_ERR_UNSUPPORTED, "GWPhase is not defined for IMRPhenomHM (its Phi is per mode)");
    if (n == 0) return GWF_OK;
    EventsDev ev;
    for (int i = 0; i < GWF_NPARAM_IN; ++i) ev.p[i] = events->p[i];
    const SignalOut out = {outp->Ap, outp->Ac, outp->psi, outp->strain, outp->Fp, outp->Fc, outp->dt};
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    switch (model->id) {
        case GWF_TAYLORF2:
            if ((model->flags & GWF_MODEL_ECCENTRIC) && !ev.p[15]) return fail(GWF_ERR_ARG, "eccentric model needs ecc");
            if ((model->flags & GWF_MODEL_TIDAL) && (!ev.p[11] || !ev.p[12])) return fail(GWF_ERR_ARG, "tidal model needs Lambda1, Lambda2");
            return run_signal_grid<kTaylorF2>(model, det, rot_deg, ev, n, f, res, f_is_2d, out, workspace, workspace_bytes, st);
        case GWF_IMRPHENOMD: return run_signal_grid<kPhenomD>(model, det, rot_deg, ev, n, f, res, f_is_2d, out, workspace, workspace_bytes, st);
        case GWF_IMRPHENOMD_NRTIDALV2:
            if (!ev.p[11] || !ev.p[12]) return fail(GWF_ERR_ARG, "tidal model needs Lambda1, Lambda2");
            return run_signal_grid<kNRTidalv2>(model, det, rot_deg, ev, n, f, res, f_is_2d, out, workspace, workspace_bytes, st);
        case GWF_IMRPHENOMHM: return run_signal_grid<kPhenomHM>(model, det, rot_deg, ev, n, f, res, f_is_2d, out, workspace, workspace_bytes, st);
        case GWF_IMRPHENOMNSBH: return run_signal_grid<kNSBH>(model, det, rot_deg, ev, n, f, res, f_is_2d, out, workspace, workspace_bytes, st);
        default: return fail(GWF_ERR_ARG, "unknown model");
    }
}

int gwf_pattern(const gwf_detector* det, double rot_deg, const double* theta, const double* phi, const double* t, const double* psi, int64_t m,
                double* Fp, double* Fc, double* dt, void* stream) {
    if (!det || !theta || !phi || !t) return fail(GWF_ERR_ARG, "gwf_pattern: null argument");
    if ((Fp || Fc) && !psi) return fail(GWF_ERR_ARG, "gwf_pattern: the pattern functions need psi");
    if (det->shape != 0 && det->shape != 1) return fail(GWF_ERR_ARG, "Enter valid detector configuration");
    if (m < 0) return fail(GWF_ERR_ARG, "negative count");
    if (m == 0) return GWF_OK;
    DetDev d;
    ArmDev a;
    single_arm(*det, rot_deg, d, a);
    pattern_kernel<<<(unsigned)((m + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(d, a, theta, phi, t, psi, m, Fp, Fc, dt);
    GWF_CUDA(cudaGetLastError());
    return GWF_OK;
}

int gwf_covariance(const double* fisher, int64_t n, int32_t nP, int32_t method, double thresh, double* cov, double* inv_err, int32_t* status,
                   void* stream) {
    if (!fisher || !cov || !inv_err) return fail(GWF_ERR_ARG, "gwf_covariance: null argument");
    if (nP < 1 || nP > kCovMaxP) return fail(GWF_ERR_ARG, "gwf_covariance: nP must be in [1, 16]");
    if (method < 0 || method > 3) return fail(GWF_ERR_ARG, "gwf_covariance: unknown method");
    if (n < 0) return fail(GWF_ERR_ARG, "negative event count");
    if (n == 0) return GWF_OK;
    const int tb = 64;
    covariance_kernel<<<(unsigned)((n + tb - 1) / tb), tb, 0, reinterpret_cast<cudaStream_t>(stream)>>>(fisher, n, nP, method, thresh, cov, inv_err, status);
    GWF_CUDA(cudaGetLastError());
    return GWF_OK;
}

int gwf_eigen(const double* fisher, int64_t n, int32_t nP, double* evals, double* evecs, double* cond, void* stream) {
    if (!fisher || !evals || !cond) return fail(GWF_ERR_ARG, "gwf_eigen: null argument");
    if (nP < 1 || nP > kCovMaxP) return fail(GWF_ERR_ARG, "gwf_eigen: nP must be in [1, 16]");
    if (n < 0) return fail(GWF_ERR_ARG, "negative event count");
    if (n == 0) return GWF_OK;
    const int tb = 64;
    eigen_kernel<<<(unsigned)((n + tb - 1) / tb), tb, 0, reinterpret_cast<cudaStream_t>(stream)>>>(fisher, n, nP, evals, evecs, cond);
    GWF_CUDA(cudaGetLastError());
    return GWF_OK;
}

int gwf_inversion_error(const double* fisher, const double* cov, int64_t n, int32_t nP, double* err, void* stream) {
    if (!fisher || !cov || !err || nP < 1) return fail(GWF_ERR_ARG, "gwf_inversion_error: bad arguments");
    if (n == 0) return GWF_OK;
    const int tb = 128;
    inversion_error_kernel<<<(unsigned)((n + tb - 1) / tb), tb, 0, reinterpret_cast<cudaStream_t>(stream)>>>(fisher, cov, n, nP, err);
    GWF_CUDA(cudaGetLastError());
    return GWF_OK;
}

}  // extern "C"
