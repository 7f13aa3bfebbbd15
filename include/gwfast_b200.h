/* gwfast_b200 -- C ABI of the B200-native Fisher/SNR engine (libgwfast_b200.so).
 *
 * The reference (CosmoStatGW/gwfast) has no FFI layer: the hot path sits behind Python methods.  Each entry
 * point below names the reference method whose numerical work it replaces; the Python classes in
 * gwfast_b200/{waveforms,signal,network}.py keep the reference's names and call these through ctypes
 * (INTEGRATION.md shows the stub).  Conventions:
 *   - plain C types only; every array pointer is a DEVICE pointer unless the name ends in _host;
 *   - all arithmetic is FP64; events are SoA (one array of length N per parameter);
 *   - calls are stream-ordered on `stream` (a cudaStream_t passed as void*), never synchronise, and allocate
 *     nothing: the caller passes the workspace (gwf_workspace_bytes);
 *   - return 0 on success, a negative gwf_status otherwise; gwf_last_error() gives the text.
 */
#ifndef GWFAST_B200_H
#define GWFAST_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GWF_VERSION 100 /* 0.1.0 */

typedef enum {
    GWF_OK = 0,
    GWF_ERR_ARG = -1,        /* invalid argument (null pointer, bad id, too many detectors ...) */
    GWF_ERR_UNSUPPORTED = -2, /* combination not built (e.g. eccentric TaylorF2) */
    GWF_ERR_CUDA = -3,       /* a CUDA runtime call failed */
    GWF_ERR_WORKSPACE = -4   /* workspace too small */
} gwf_status;

/* waveform models: gwfast/waveforms.py classes at :697, :959, :1339, :1838, :2752 */
typedef enum { GWF_TAYLORF2 = 0, GWF_IMRPHENOMD = 1, GWF_IMRPHENOMD_NRTIDALV2 = 2, GWF_IMRPHENOMHM = 3, GWF_IMRPHENOMNSBH = 4 } gwf_model_id;

/* gwf_model.flags -- constructor options of the reference classes that change the arithmetic */
#define GWF_MODEL_TIDAL 1          /* TaylorF2_RestrictedPN(is_tidal=True)            waveforms.py:850-855 */
#define GWF_MODEL_3P5PN_SPINHO 2   /* TaylorF2_RestrictedPN(use_3p5PN_SpinHO=True)    waveforms.py:808-812 */
#define GWF_MODEL_PHIREF_VLSO 4    /* TaylorF2_RestrictedPN(phiref_vlso=True)         waveforms.py:796-802 */
#define GWF_MODEL_QUADMON_TID 8    /* TaylorF2_RestrictedPN(use_QuadMonTid=True)      waveforms.py:774-782 */
#define GWF_MODEL_KERR_ISCO 16     /* TaylorF2_RestrictedPN(which_ISCO='Kerr')        waveforms.py:922-953 */
#define GWF_MODEL_NO_FCUT 32       /* WaveFormModel(apply_fcut=False)                 waveforms.py:1142-1153 */
#define GWF_MODEL_HAS_FREF 64      /* IMRPhenomD(fRef=...)                            waveforms.py:1139-1141 */
#define GWF_MODEL_LAMBDA_GIVEN 128 /* events carried Lambda1/Lambda2 when fcut() ran  signal.py:884 (SURVEY A-20) */
#define GWF_MODEL_ECCENTRIC 512    /* TaylorF2_RestrictedPN(is_eccentric=True): parameter ecc appended (nP = 12, or 14 with tidal);
                                      fRef_ecc in gwf_model.fRef when GWF_MODEL_HAS_FREF, else v0ecc = v(fmin)   waveforms.py:814-845 */
#define GWF_MODEL_NEWTONIAN 256    /* NewtInspiral (with id GWF_TAYLORF2): leading-order phase, base-class tau_star; the engine
                                      still returns the 11-parameter layout, whose eta/chi rows the 8-parameter NewtInspiral
                                      contract drops on the host (waveforms.py:94-96, 205-260; signal.py:1143-1151) */

typedef struct {
    int32_t id;      /* gwf_model_id */
    int32_t flags;   /* GWF_MODEL_* */
    double fcutPar;  /* WaveFormModel.fcutPar (waveforms.py:75) */
    double fRef;     /* Hz, used if GWF_MODEL_HAS_FREF */
} gwf_model;

/* one GWSignal (gwfast/signal.py:68-160) */
typedef struct {
    double lat_rad, long_rad, xax_rad; /* det_lat_rad, det_long_rad, det_xax_rad (signal.py:113-116) */
    int32_t shape;                     /* 0 = 'L', 1 = 'T' (signal.py:144-147) */
    int32_t use_earth_motion;          /* useEarthMotion (signal.py:136) */
    int32_t no_motion;                 /* noMotion (signal.py:137) */
    int32_t psd;                       /* index into the psds[] array passed with the call */
    double fmin, fmax;                 /* Hz; fmax <= 0 means None (signal.py:141-142) */
} gwf_detector;

/* PSD table on the device: built once per GWSignal from strainFreq/noiseCurve (signal.py:122-130) */
typedef struct gwf_psd gwf_psd;
int gwf_psd_create(const double* freq_host, const double* psd_host, int32_t n, gwf_psd** out);
void gwf_psd_destroy(gwf_psd* psd);

/* QNM ringdown tables WFfiles/QNMData_{a,fring,fdamp}.txt (waveforms.py:988-990); host arrays, copied once */
int gwf_set_qnm_tables(const double* a_host, const double* fring_host, const double* fdamp_host, int32_t n);

/* IMRPhenomNSBH's xi_tide table (waveforms.py:3286-3373: WFfiles/xiTide_Table_200.h5, or IMRPhenomNSBH._tabulate_xiTide with
 * numpy.roots): 200^3 doubles over (compactness in [0.1, 0.5], q in [1, 100], chi_BH in [-1, 1]), row-major.  The current device
 * computes its own copy the first time the model runs there; this call makes sure it exists and, if table_host is not NULL,
 * copies it out (what IMRPhenomNSBH.xiTide_interp tabulates). */
int gwf_xitide_table(double* table_host);

/* the events dict as device SoA; order of p[]:
 * Mc eta dL theta phi iota psi tcoal Phicoal chi1z chi2z Lambda1 Lambda2 fcut Mtot_sec ecc
 * p[11], p[12] may be NULL for BBH; p[15] (orbital eccentricity e0) is read by the eccentric TaylorF2 only.  p[13] (wf_model.fcut(**events) in Hz, signal.py:715/884) and p[14]
 * (M*GMsun_over_c3 in seconds, waveforms.py:1026) are optional: when the host passes the values it computed with the
 * reference's own expressions, the last grid sample lands on Mf = fcutPar with the reference's rounding, which decides
 * whether that sample is inside the waveform cut (waveforms.py:1145-1147); when NULL both are computed on the device. */
#define GWF_NPARAM_IN 16
typedef struct {
    const double* p[GWF_NPARAM_IN];
} gwf_events;

/* gwf_opts.flags -- keyword arguments of GWSignal.FisherMatr (signal.py:782-786) */
#define GWF_OPT_M1M2 1       /* use_m1m2=True      */
#define GWF_OPT_CHIS_CHIA 2  /* use_chi1chi2=False */
#define GWF_OPT_LIN_GRID 4   /* spacing='lin'      */
#define GWF_OPT_REUSE_WORKSPACE 8 /* the workspace still holds the coefficient records of the previous gwf_fisher call on the
                                     same events/model/options: skip the prologue kernel (used to time the main kernel alone) */
#define GWF_OPT_GENERIC_LOOP 16   /* run the general detector loop even when the network qualifies for the unrolled single-group
                                     form with all PSD windows in shared memory (tests compare the two kernels) */
#define GWF_OPT_ONE_WARP_PER_EVENT 32 /* never split an event over two warps (the launcher does that when it shortens the
                                     persistent loop of a small catalog; tests compare the two mappings) */
#define GWF_OPT_HM_BLOCK_PAIRS 64 /* IMRPhenomHM: the two warps of a pair take alternate blocks of 32 samples (the mapping of the other models)
                                     instead of splitting the modes and the packed entries of the same samples (tests compare the two) */
typedef struct {
    int32_t res;    /* frequency samples per event (res=1000) */
    int32_t flags;  /* GWF_OPT_* */
    int32_t per_arm; /* 0: one Fisher summed over the whole network (DetNet.FisherMatr default, network.py:118);
                        1: one Fisher per arm, triangle arms in the order 0, 60 deg, -(1+2) (return_all=True) */
    int32_t reserved;
} gwf_opts;

int gwf_version(void);
const char* gwf_last_error(void);

/* number of Fisher parameters (WaveFormModel.nParams, waveforms.py:87-124) and of arms (1 per L, 3 per T) */
int gwf_num_params(const gwf_model* model);
int gwf_num_arms(const gwf_detector* dets, int32_t ndet);

/* bytes of device workspace needed by gwf_fisher / gwf_snr for n events */
size_t gwf_workspace_bytes(const gwf_model* model, int64_t n);

/* Replaces DetNet.FisherMatr / GWSignal.FisherMatr (network.py:84-123, signal.py:782-1098) including
 * _SignalDerivatives + _AnalyticalDerivatives (signal.py:1102-1584).
 *   fisher_packed: per_arm=0 -> [n][nP(nP+1)/2], per_arm=1 -> [n_arms][n][nP(nP+1)/2]; lower triangle, row-major,
 *                  rows/cols in ParNums order (waveforms.py:78-147)
 *   snr2:          same leading shape, [n] per block: 4 int |h|^2/Sn df of the arms in the block
 *                  (= SNR^2 for the non-HM models; may be NULL) */
int gwf_fisher(const gwf_model* model, const gwf_detector* dets, int32_t ndet, const gwf_psd* const* psds, int32_t npsd,
               const gwf_events* events, int64_t n, const gwf_opts* opts, double* fisher_packed, double* snr2,
               void* workspace, size_t workspace_bytes, void* stream);

/* Per-event status word (SURVEY.md 5 "failure detection"): a bad event gives NaN rows in the reference; here it is also flagged. */
#define GWF_EV_NONFINITE_INPUT 1   /* a parameter of the event is NaN or infinite */
#define GWF_EV_OUT_OF_DOMAIN 2     /* Mc <= 0, dL <= 0 or eta outside (0, 0.25] */
#define GWF_EV_NONFINITE_OUTPUT 4  /* a Fisher element (or SNR^2) of the event came out NaN or infinite */
#define GWF_EV_EMPTY_GRID 8        /* fcut <= fmin for some detector: the frequency grid of the event is degenerate */

/* Outputs of gwf_fisher_ex; every pointer except fisher_packed may be NULL.  "blocks" = 1 (per_arm=0) or n_arms (per_arm=1). */
typedef struct {
    double* fisher_packed; /* [blocks][n][nP(nP+1)/2], as gwf_fisher */
    double* snr2;          /* [blocks][n]: 4 int |h|^2/Sn df of the arms in the block */
    double* snr2_integ;    /* [blocks][n]: the integral GWSignal.SNRInteg forms for the same arms (signal.py:725-728), so that one launch
                              serves DetNet.SNR and DetNet.FisherMatr of the same events: equal to snr2 except for IMRPhenomHM, whose
                              SNR drops the +/x cross term (signal.py:457-460) */
    double* snr_derivs;    /* [blocks][n][nP]: 4 Re int conj(d_i h) h / Sn df, rows in ParNums order, the tcoal row per second like the
                              Fisher (return_SNR_derivatives, signal.py:938-945; the reference divides the sum over arms by the network
                              SNR on the host, network.py:143) */
    int32_t* status;       /* [n]: GWF_EV_* bits, 0 = clean */
    double* const* peer_fisher; /* multi-GPU (per_arm = 0 only): npeers device pointers [n][nP(nP+1)/2], this rank's slot in the gathered
                              buffer of every rank of the box (peers mapped with gwf_peer_open); the kernel stores the finished packed
                              row of every event there as soon as the event is done, so the all-gather of the Fisher matrices
                               overlaps the computation.  The ranks synchronise before reading.  NULL / 0: none */
    int32_t npeers;
    int32_t reserved;
} gwf_fisher_out;

/* gwf_fisher with all optional outputs (return_SNR_derivatives of GWSignal.FisherMatr, the SNR of the same launch, status words). */
int gwf_fisher_ex(const gwf_model* model, const gwf_detector* dets, int32_t ndet, const gwf_psd* const* psds, int32_t npsd,
                  const gwf_events* events, int64_t n, const gwf_opts* opts, const gwf_fisher_out* out,
                  void* workspace, size_t workspace_bytes, void* stream);

/* gwf_fisher_ex in two parts, so that a caller can overlap the device->host copy of one group of events with the kernels of the next
 * without paying one prologue launch and one ragged last round of the persistent kernel per group: the workspace is laid out for
 * n_total events;  phases & 1: the prologue over ALL n_total events (out->status: [n_total], other outputs unused);
 * phases & 2: the Fisher kernels on the events [lo, lo + m) -- every pointer of `out` then refers to that range ([blocks][m][...],
 * status + lo).  Group boundaries at multiples of gwf_round_events() keep every CTA of the persistent grid equally loaded. */
int gwf_fisher_range(const gwf_model* model, const gwf_detector* dets, int32_t ndet, const gwf_psd* const* psds, int32_t npsd,
                     const gwf_events* events, int64_t n_total, int64_t lo, int64_t m, int32_t phases, const gwf_opts* opts,
                     const gwf_fisher_out* out, void* workspace, size_t workspace_bytes, void* stream);
/* events one round of the persistent Fisher grid takes on the current device (SMs x events per CTA for this model) */
int64_t gwf_round_events(const gwf_model* model);

/* The return_derivatives output of GWSignal.FisherMatr (signal.py:917-945, network.py:124-141): the derivative strain itself,
 *   derivs: complex128 (re, im) [n_arms][nP][n][res], one block per arm (triangle arms 0, 60 deg, -(1+2)), rows in ParNums order, the
 *           tcoal row per second (signal.py:920); samples beyond the waveform cut are 0.  TaylorF2, IMRPhenomD, IMRPhenomD_NRTidalv2.
 * This is the array gwf_fisher never writes (nP * res * 16 B per event and arm); use it only when the strain derivatives themselves
 * are wanted. */
int gwf_strain_derivs(const gwf_model* model, const gwf_detector* dets, int32_t ndet, const gwf_psd* const* psds, int32_t npsd,
                      const gwf_events* events, int64_t n, const gwf_opts* opts, double* derivs,
                      void* workspace, size_t workspace_bytes, void* stream);

/* GWSignal.GWstrain on the engine's grid (signal.py:484-655), the building block of GWSignal.WFOverlap (signal.py:1759-1930):
 *   strain: complex128 (re, im) [n_arms][n][res], one block per arm (triangle arms 0, 60 deg, -(1+2)); the dict entries go straight to
 *           the waveform (no Fisher re-parametrisation); the grid is geomspace(fmin, events->p[13] if given else the model's fcut, res);
 *           samples beyond the waveform cut are 0.  TaylorF2 (all options), IMRPhenomD, IMRPhenomD_NRTidalv2. */
int gwf_strain(const gwf_model* model, const gwf_detector* dets, int32_t ndet, const gwf_psd* const* psds, int32_t npsd,
               const gwf_events* events, int64_t n, const gwf_opts* opts, double* strain,
               void* workspace, size_t workspace_bytes, void* stream);

/* The three integrals of GWSignal.WFOverlap for ONE arm (signal.py:1876-1925) from two gwf_strain outputs on the same grid:
 *   h1, h2: complex128 [n][res];  fcut: [n] upper grid frequency (the larger of the two waveforms' cuts, signal.py:1853-1858)
 *   overlap[n] = 4 int Re(h1 conj h2)/Sn df,  snr2_1[n] = 4 int |h1|^2/Sn df,  snr2_2[n] = 4 int |h2|^2/Sn df */
int gwf_overlap(const double* h1, const double* h2, const double* fcut, int64_t n, int32_t res, double fmin, const gwf_psd* psd,
                double* overlap, double* snr2_1, double* snr2_2, void* stream);

/* Replaces DetNet.SNR / GWSignal.SNRInteg (network.py:53-81, signal.py:658-777).
 *   snr2_arm: [n_arms][n], the per-arm integrals 4 int (Ap^2+Ac^2)/Sn df (SNR_arm = sqrt of it) */
int gwf_snr(const gwf_model* model, const gwf_detector* dets, int32_t ndet, const gwf_psd* const* psds, int32_t npsd,
            const gwf_events* events, int64_t n, const gwf_opts* opts, double* snr2_arm,
            void* workspace, size_t workspace_bytes, void* stream);

/* packed lower triangle [n][nP(nP+1)/2] -> the reference's (nP, nP, N) layout, event axis fastest (signal.py:924) */
int gwf_unpack_fisher(const double* packed, int64_t n, int32_t nP, double* full, void* stream);
/* the same into a slice of a larger (nP, nP, ld) array: plane (i, j) of the chunk starts at full + (i nP + j) ld; ld >= n.  Lets the
 * chunks of a catalog land in the one array DetNet.FisherMatr returns (network.py:118) */
int gwf_unpack_fisher_ld(const double* packed, int64_t n, int32_t nP, double* full, int64_t ld, void* stream);
/* Multi-GPU (events sharded over the GPUs of one box, one final all-gather of the Fisher matrices; the reference's
 * counterpart is the result files its batch pool writes, run/calculate_forecasts_from_catalog.py:900-1024): gwf_unpack_fisher_ld
 * fused with the write side of an all-gather over NVLink peer memory.  peer_slots[q] (q < npeers <= 8) points at THIS rank's slot
 * [n][nP(nP+1)/2] inside peer q's gathered buffer (a device pointer of another GPU mapped into this process with CUDA IPC, or of
 * this GPU); the kernel stores the packed rows there while it transposes them into `full` (which may be NULL).  The ranks must
 * synchronise (any barrier) before the gathered buffers are read. */
/* the gathered buffer of a rank and its peers' views of it: gwf_peer_alloc = cudaMalloc (zero-filled) + cudaIpcGetMemHandle; the 64-byte
 * handle travels to the other ranks by any means (the process group), which map it with gwf_peer_open (cudaIpcOpenMemHandle with lazy peer
 * access, opened on the importing rank's own device so that its kernels can store into it).  gwf_peer_close / gwf_peer_free undo them. */
#define GWF_IPC_HANDLE_BYTES 64
int gwf_peer_alloc(size_t bytes, void** ptr_out, unsigned char* handle_out);
int gwf_peer_open(const unsigned char* handle, void** ptr_out);
int gwf_peer_close(void* ptr);
int gwf_peer_free(void* ptr);
int gwf_unpack_gather(const double* packed, int64_t n, int32_t nP, double* full, int64_t ld, double* const* peer_slots, int32_t npeers,
                      void* stream);
/* stream-ordered 2-D copy (cudaMemcpy2DAsync, direction from the pointers): moves a chunk's (nP nP) x n planes between the device
 * array and the pinned host array the caller returns, while the next chunk computes */
int gwf_copy_2d(void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t width_bytes, size_t height, void* stream);

/* Replaces WaveFormModel.Phi / Ampl / tau_star / fcut (and IMRPhenomHM.hphc) evaluated on a user grid (waveforms.py:149-199,
 * 2256-2616), with the dict entries handed straight to the waveform (no Fisher re-parametrisation).
 *   f:        [res][n] if f_is_2d else [res] (shared by all events); the PhenomD-family reference frequency is min_k f (waveforms.py:1139)
 *   phi_out:  [nm][res][n], ampl_out: [nm][res][n] with nm = 1, or 6 for IMRPhenomHM in the order 21, 22, 32, 33, 43, 44 (waveforms.py:2075)
 *   tau_out:  [res][n];  hphc_out: [4][res][n] = Re hp, Im hp, Re hc, Im hc (IMRPhenomHM only);  fcut_out: [n]
 * Any output may be NULL; res may be 0 (f NULL) if only fcut_out is wanted. */
int gwf_waveform(const gwf_model* model, const gwf_events* events, int64_t n, const double* f, int32_t res, int32_t f_is_2d,
                 double* phi_out, double* ampl_out, double* tau_out, double* hphc_out, double* fcut_out,
                 void* workspace, size_t workspace_bytes, void* stream);

/* Replaces GWSignal.GWAmplitudes / GWPhase / GWstrain (gwfast/signal.py:425-655) on a USER frequency grid, for one detector at the arm
 * orientation xax + rot_deg (the `rot` argument, degrees; the triangle's arms are rot = 0, 60 and -(1+2)).  The dict entries go
 * straight to the waveform (GWstrain's own re-parametrisation flags -- is_m1m2, is_chi1chi2, is_Lam1Lam2 -- are applied by the caller).
 *   f: [res][n] if f_is_2d else [res];  every output is [res][n] and may be NULL:
 *   Ap, Ac   GWAmplitudes (A Fp (1+cos^2 iota)/2, A Fc cos iota; IMRPhenomHM: |hp| Fp, |hc| Fc, signal.py:457-464)
 *   psi      GWPhase = 2 pi f tcoal 86400 - Phicoal - Phi(f)   (signal.py:484; not for IMRPhenomHM)
 *   strain   complex128 (re, im): (Ap + i Ac) exp(i (psi + 2 pi f Delta t))   (IMRPhenomHM: (hp Fp + hc Fc) exp(i(...)), signal.py:584-607)
 *   Fp, Fc   the pattern functions at the time the response uses (t = tcoal - tau(f)/86400 with useEarthMotion, + Delta t/86400)
 *   dt       Delta t_loc in seconds at that time (signal.py:401-423) */
typedef struct {
    double *Ap, *Ac, *psi, *strain, *Fp, *Fc, *dt;
} gwf_signal_out;
int gwf_signal_grid(const gwf_model* model, const gwf_detector* det, double rot_deg, const gwf_events* events, int64_t n, const double* f,
                    int32_t res, int32_t f_is_2d, const gwf_signal_out* out, void* workspace, size_t workspace_bytes, void* stream);

/* Replaces GWSignal._PatternFunction(theta, phi, t, psi, rot) and GWSignal._DeltLoc(theta, phi, t) (signal.py:342-423), element-wise over
 * m points (device arrays): the pattern functions at exactly the time t (GMST, days) and the Earth-centre -> site delay in seconds.
 * Fp, Fc, dt may be NULL; psi may be NULL if only dt is wanted. */
int gwf_pattern(const gwf_detector* det, double rot_deg, const double* theta, const double* phi, const double* t, const double* psi, int64_t m,
                double* Fp, double* Fc, double* dt, void* stream);

/* Replaces fisherTools.CovMatr (gwfast/fisherTools.py:32-196): per event, normalise by the diagonal (ws F ws), invert, symmetrise,
 * undo the normalisation; double-double arithmetic on the device instead of the reference's per-event mpmath loop.
 *   fisher, cov: device (nP, nP, N) arrays, event axis fastest (the layout FisherMatr returns);  inv_err: [N] = max|cov F - 1|
 *   method: 0 Cholesky ('cho'; also 'inv'/'lu'), falling back to the symmetric eigen-decomposition when the matrix is not positive
 *           definite (the reference's alt_method='svd'); 1 eigen ('svd'); 2 eigen with singular values below thresh*max raised to it
 *           (truncate=True, svals_thresh); 3 eigen with singular values <= thresh excluded ('svd_reg')
 *   status: [N] or NULL: 0 Cholesky, 1 eigen route, 2 all-NaN input (NaN output, fisherTools.py:64-67), 3 non-finite result,
 *           4 zero on the diagonal (normalisation skipped, fisherTools.py:97-99) */
int gwf_covariance(const double* fisher, int64_t n, int32_t nP, int32_t method, double thresh, double* cov, double* inv_err, int32_t* status,
                   void* stream);

/* Replaces fisherTools.CheckFisher (fisherTools.py:216-279): eigenvalues ascending [nP][N], eigenvectors [nP][nP][N] (component i of
 * vector k at (i, k); may be NULL), condition number max|lambda|/min|lambda| [N]. */
int gwf_eigen(const double* fisher, int64_t n, int32_t nP, double* evals, double* evecs, double* cond, void* stream);

/* Replaces fisherTools.compute_inversion_error (fisherTools.py:199-211): err[N] = max |cov F - 1|. */
int gwf_inversion_error(const double* fisher, const double* cov, int64_t n, int32_t nP, double* err, void* stream);

/* Diagnostics (no reference counterpart): sustained FP64 FMA throughput of the current device in TFLOP/s, measured with a
 * register-resident DFMA chain kernel over ~`ms` milliseconds; used by bench.py as the measured FP64 roofline denominator.
 * Synchronises the stream. */
int gwf_fp64_peak(double ms, double* tflops_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GWFAST_B200_H */
