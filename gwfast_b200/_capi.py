"""ctypes mirror of ``include/gwfast_b200.h`` and loader of ``lib/libgwfast_b200.so``.

The library is the product: there is no Python/numpy fallback.  If it cannot be loaded (not built, wrong
arch) every entry point of the package raises :class:`EngineUnavailable`.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('GWFAST_B200_LIB', os.path.join(_HERE, 'lib', 'libgwfast_b200.so'))   # override: kernel experiments

GWF_TAYLORF2, GWF_IMRPHENOMD, GWF_IMRPHENOMD_NRTIDALV2, GWF_IMRPHENOMHM, GWF_IMRPHENOMNSBH = 0, 1, 2, 3, 4
GWF_MODEL_TIDAL, GWF_MODEL_3P5PN_SPINHO, GWF_MODEL_PHIREF_VLSO, GWF_MODEL_QUADMON_TID = 1, 2, 4, 8
GWF_MODEL_KERR_ISCO, GWF_MODEL_NO_FCUT, GWF_MODEL_HAS_FREF, GWF_MODEL_LAMBDA_GIVEN, GWF_MODEL_NEWTONIAN, GWF_MODEL_ECCENTRIC = 16, 32, 64, 128, 256, 512
GWF_OPT_M1M2, GWF_OPT_CHIS_CHIA, GWF_OPT_LIN_GRID, GWF_OPT_REUSE_WORKSPACE, GWF_OPT_GENERIC_LOOP, GWF_OPT_ONE_WARP_PER_EVENT, GWF_OPT_HM_BLOCK_PAIRS = 1, 2, 4, 8, 16, 32, 64
GWF_NPARAM_IN = 16
# order of gwf_events.p[]
EVENT_KEYS = ('Mc', 'eta', 'dL', 'theta', 'phi', 'iota', 'psi', 'tcoal', 'Phicoal', 'chi1z', 'chi2z', 'Lambda1', 'Lambda2',
              '_fcut', '_Mtot_sec', 'ecc')


class gwf_model(C.Structure):
    _fields_ = [('id', C.c_int32), ('flags', C.c_int32), ('fcutPar', C.c_double), ('fRef', C.c_double)]


class gwf_detector(C.Structure):
    _fields_ = [('lat_rad', C.c_double), ('long_rad', C.c_double), ('xax_rad', C.c_double), ('shape', C.c_int32),
                ('use_earth_motion', C.c_int32), ('no_motion', C.c_int32), ('psd', C.c_int32), ('fmin', C.c_double),
                ('fmax', C.c_double)]


class gwf_events(C.Structure):
    _fields_ = [('p', C.c_void_p * GWF_NPARAM_IN)]


class gwf_opts(C.Structure):
    _fields_ = [('res', C.c_int32), ('flags', C.c_int32), ('per_arm', C.c_int32), ('reserved', C.c_int32)]


class gwf_fisher_out(C.Structure):
    _fields_ = [('fisher_packed', C.c_void_p), ('snr2', C.c_void_p), ('snr2_integ', C.c_void_p), ('snr_derivs', C.c_void_p), ('status', C.c_void_p),
                ('peer_fisher', C.c_void_p), ('npeers', C.c_int32), ('reserved', C.c_int32)]


class gwf_signal_out(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ('Ap', 'Ac', 'psi', 'strain', 'Fp', 'Fc', 'dt')]


# per-event status bits (gwf_fisher_out.status)
GWF_EV_NONFINITE_INPUT, GWF_EV_OUT_OF_DOMAIN, GWF_EV_NONFINITE_OUTPUT, GWF_EV_EMPTY_GRID = 1, 2, 4, 8


class EngineUnavailable(RuntimeError):
    pass


class EngineError(RuntimeError):
    pass


# every symbol include/gwfast_b200.h declares (tests check that the library exports all of them)
SYMBOLS = ('gwf_version', 'gwf_last_error', 'gwf_num_params', 'gwf_num_arms', 'gwf_workspace_bytes', 'gwf_psd_create',
           'gwf_psd_destroy', 'gwf_set_qnm_tables', 'gwf_xitide_table', 'gwf_fisher', 'gwf_fisher_ex', 'gwf_fisher_range', 'gwf_round_events', 'gwf_strain_derivs', 'gwf_strain', 'gwf_overlap', 'gwf_snr', 'gwf_unpack_fisher', 'gwf_unpack_fisher_ld', 'gwf_unpack_gather', 'gwf_peer_alloc', 'gwf_peer_open', 'gwf_peer_close', 'gwf_peer_free', 'gwf_copy_2d', 'gwf_waveform', 'gwf_signal_grid', 'gwf_pattern', 'gwf_fp64_peak', 'gwf_covariance', 'gwf_eigen',
           'gwf_inversion_error')

_lib = None


def load():
    """Load the shared library (once) and declare the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineUnavailable('%s not built: run `python -c "import __graft_entry__ as g; g.build()"` '
                                '(there is no CPU fallback)' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    P = C.POINTER
    lib.gwf_version.restype = C.c_int
    lib.gwf_last_error.restype = C.c_char_p
    lib.gwf_num_params.argtypes = [P(gwf_model)]
    lib.gwf_num_arms.argtypes = [P(gwf_detector), i32]
    lib.gwf_workspace_bytes.argtypes = [P(gwf_model), i64]
    lib.gwf_workspace_bytes.restype = C.c_size_t
    lib.gwf_psd_create.argtypes = [P(dbl), P(dbl), i32, P(vp)]
    lib.gwf_psd_destroy.argtypes = [vp]
    lib.gwf_psd_destroy.restype = None
    lib.gwf_set_qnm_tables.argtypes = [P(dbl), P(dbl), P(dbl), i32]
    lib.gwf_xitide_table.argtypes = [vp]
    common = [P(gwf_model), P(gwf_detector), i32, P(vp), i32, P(gwf_events), i64, P(gwf_opts)]
    lib.gwf_fisher.argtypes = common + [vp, vp, vp, C.c_size_t, vp]
    lib.gwf_fisher_ex.argtypes = common + [P(gwf_fisher_out), vp, C.c_size_t, vp]
    lib.gwf_fisher_range.argtypes = [P(gwf_model), P(gwf_detector), i32, P(vp), i32, P(gwf_events), i64, i64, i64, i32, P(gwf_opts), P(gwf_fisher_out), vp, C.c_size_t, vp]
    lib.gwf_round_events.argtypes = [P(gwf_model)]
    lib.gwf_round_events.restype = i64
    lib.gwf_snr.argtypes = common + [vp, vp, C.c_size_t, vp]
    lib.gwf_strain_derivs.argtypes = common + [vp, vp, C.c_size_t, vp]
    lib.gwf_strain.argtypes = common + [vp, vp, C.c_size_t, vp]
    lib.gwf_overlap.argtypes = [vp, vp, vp, i64, i32, dbl, vp, vp, vp, vp, vp]
    lib.gwf_unpack_fisher.argtypes = [vp, i64, i32, vp, vp]
    lib.gwf_unpack_fisher_ld.argtypes = [vp, i64, i32, vp, i64, vp]
    lib.gwf_peer_alloc.argtypes = [C.c_size_t, P(vp), C.c_char_p]
    lib.gwf_peer_open.argtypes = [C.c_char_p, P(vp)]
    lib.gwf_peer_close.argtypes = [vp]
    lib.gwf_peer_free.argtypes = [vp]
    lib.gwf_unpack_gather.argtypes = [vp, i64, i32, vp, i64, P(vp), i32, vp]
    lib.gwf_copy_2d.argtypes = [vp, C.c_size_t, vp, C.c_size_t, C.c_size_t, C.c_size_t, vp]
    lib.gwf_waveform.argtypes = [P(gwf_model), P(gwf_events), i64, vp, i32, i32, vp, vp, vp, vp, vp, vp, C.c_size_t, vp]
    lib.gwf_signal_grid.argtypes = [P(gwf_model), P(gwf_detector), dbl, P(gwf_events), i64, vp, i32, i32, P(gwf_signal_out), vp, C.c_size_t, vp]
    lib.gwf_pattern.argtypes = [P(gwf_detector), dbl, vp, vp, vp, vp, i64, vp, vp, vp, vp]
    lib.gwf_fp64_peak.argtypes = [dbl, P(dbl), vp]
    lib.gwf_covariance.argtypes = [vp, i64, i32, i32, dbl, vp, vp, vp, vp]
    lib.gwf_eigen.argtypes = [vp, i64, i32, vp, vp, vp, vp]
    lib.gwf_inversion_error.argtypes = [vp, vp, i64, i32, vp, vp]
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        raise EngineError('%s failed (%d): %s' % (what, rc, load().gwf_last_error().decode()))
