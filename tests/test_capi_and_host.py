"""C ABI surface and host-side contract, no GPU: the library loads and exports every symbol the header declares,
the ctypes mirrors match the C structs, and the gwfast-style classes keep the reference's bookkeeping and errors."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def _built():
    import __graft_entry__ as g
    g.build()
    from gwfast_b200 import _capi
    return _capi


def test_library_exports_every_declared_symbol():
    K = _built()
    header = open(os.path.join(ROOT, 'include', 'gwfast_b200.h')).read()
    declared = set(re.findall(r'\b(gwf_[a-z_0-9]+)\s*\(', header))
    assert declared == set(K.SYMBOLS)
    lib = K.load()
    for s in declared:
        assert hasattr(lib, s), s
    assert lib.gwf_version() == 100


def test_struct_layouts_and_pure_host_entry_points():
    K = _built()
    lib = K.load()
    assert C.sizeof(K.gwf_model) == 24 and C.sizeof(K.gwf_detector) == 56 and C.sizeof(K.gwf_opts) == 16 and C.sizeof(K.gwf_events) == 16 * 8
    for mid, flags, nP in ((0, 0, 11), (0, K.GWF_MODEL_TIDAL, 13), (1, 0, 11), (2, 0, 13), (3, 0, 11)):
        assert lib.gwf_num_params(C.byref(K.gwf_model(mid, flags, 0.2, 0.))) == nP
    dets = (K.gwf_detector * 3)(K.gwf_detector(0, 0, 0, 1, 0, 0, 0, 2., 0.), K.gwf_detector(0, 0, 0, 0, 0, 0, 0, 2., 0.), K.gwf_detector(0, 0, 0, 0, 0, 0, 0, 2., 0.))
    assert lib.gwf_num_arms(dets, 3) == 5
    assert lib.gwf_workspace_bytes(C.byref(K.gwf_model(1, 0, 0.2, 0.)), 10) > 10 * 1500
    # argument checking happens before any CUDA call
    opts = K.gwf_opts(1000, 0, 0, 0)
    rc = lib.gwf_fisher(None, dets, 3, None, 0, None, 0, C.byref(opts), None, None, None, 0, None)
    assert rc == -1 and b'null' in lib.gwf_last_error()


def test_no_cpu_fallback_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    from gwfast_b200 import waveforms, signal, synthetic
    from gwfast_b200._capi import EngineUnavailable
    s = synthetic.build_network(signal.GWSignal, waveforms.IMRPhenomD(), 'ETSL')['ETSL']
    with pytest.raises(EngineUnavailable):
        s.SNRInteg(synthetic.bbh_catalog(2, 1))
    with pytest.raises(EngineUnavailable):
        s.FisherMatr(synthetic.bbh_catalog(2, 1))


def test_parameter_ordering_contract():
    """waveforms.py:78-147."""
    from gwfast_b200 import waveforms as W
    base = ['Mc', 'eta', 'dL', 'theta', 'phi', 'iota', 'psi', 'tcoal', 'Phicoal']
    assert list(W.IMRPhenomD().ParNums) == base + ['chi1z', 'chi2z'] and W.IMRPhenomD().nParams == 11
    assert list(W.IMRPhenomD(is_chi1chi2=False).ParNums) == base + ['chiS', 'chiA']
    m = W.IMRPhenomD_NRTidalv2()
    assert list(m.ParNums) == base + ['chi1z', 'chi2z', 'LambdaTilde', 'deltaLambda'] and m.nParams == 13 and m.is_tidal and m.objType == 'BNS'
    t = W.TaylorF2_RestrictedPN(is_tidal=True)
    assert t.nParams == 13 and t.is_holomorphic and t.objType == 'BNS'
    assert W.IMRPhenomHM().is_HigherModes and W.IMRPhenomHM().nParams == 11
    assert abs(W.TaylorF2_RestrictedPN().fcutPar - 1. / (6. * np.pi * np.sqrt(6.) * 4.925491025543576e-06)) < 1e-9
    ev = dict(Mc=np.array([1.19752182]), eta=np.array([0.24786618]))
    assert abs(W.TaylorF2_RestrictedPN().fcut(**ev)[0] - 1590.08612662) < 1e-6          # notebook known answer
    assert abs(W.IMRPhenomD().fcut(**ev)[0] - 0.2 / (ev['Mc'][0] * 4.925491025543576e-06 / ev['eta'][0] ** 0.6)) < 1e-9


def test_events_dict_side_effects_and_errors():
    """gwfastUtils.py:965-1019, signal.py:86-90, 694-713."""
    from gwfast_b200 import gwfastUtils as U, signal, waveforms as W
    ev = dict(m1=np.array([30.]), m2=np.array([20.]), ra=np.array([1.]), dec=np.array([0.2]), thetaJN=np.array([0.3]), tcoal=np.array([0.1]))
    U.check_evparams(ev)
    assert {'Mc', 'eta', 'theta', 'phi', 'iota'} <= set(ev)
    assert abs(ev['eta'][0] - 0.24) < 1e-15 and abs(ev['theta'][0] - (np.pi / 2 - 0.2)) < 1e-15
    with pytest.raises(ValueError, match='tGPS and tcoal'):
        U.check_evparams(dict(Mc=np.array([1.])))
    with pytest.raises(ValueError, match='valid detector configuration'):
        signal.GWSignal(W.IMRPhenomD(), psd_path='x', detector_shape='X')
    with pytest.raises(ValueError, match='valid PSD or ASD path'):
        signal.GWSignal(W.IMRPhenomD())
    from gwfast_b200 import synthetic
    s = synthetic.build_network(signal.GWSignal, W.IMRPhenomD(), 'ETSL')['ETSL']
    bad = synthetic.bbh_catalog(2, 1)
    del bad['chi1z']
    with pytest.raises(ValueError, match='chi1z, chi2z and chiS, chiA'):
        s.SNRInteg(bad)
    assert s.strainFreq.shape == s.noiseCurve.shape and s.angbtwArms == 0.5 * np.pi and s.fmin == 2.
    L1, L2 = np.array([400.]), np.array([700.])
    lt, dl = U.Lamt_delLam_from_Lam12(L1, L2, np.array([0.245]))
    r1, r2 = U.Lam12_from_Lamt_delLam(lt, dl, np.array([0.245]))
    assert abs(r1[0] / 400. - 1) < 1e-9 and abs(r2[0] / 700. - 1) < 1e-9


def test_round_two_entry_points_check_their_arguments_before_any_cuda_call():
    K = _built()
    lib = K.load()
    m = K.gwf_model(1, 0, 0.2, 0.)
    det = K.gwf_detector(0.7, 0.1, 0., 1, 1, 0, 0, 2., 0.)
    opts = K.gwf_opts(1000, 0, 0, 0)
    fo = K.gwf_fisher_out(None, None, None, None, None)
    # gwf_fisher_range: the range must lie inside the workspace's events and a phase must be selected
    assert lib.gwf_fisher_range(C.byref(m), None, 0, None, 0, None, 10, 8, 4, 2, C.byref(opts), C.byref(fo), None, 0, None) == -1
    assert b'range' in lib.gwf_last_error()
    assert lib.gwf_fisher_range(C.byref(m), None, 0, None, 0, None, 10, 0, 10, 0, C.byref(opts), C.byref(fo), None, 0, None) == -1
    assert lib.gwf_unpack_gather(None, 4, 11, None, 4, None, 0, None) == -1
    slots = (C.c_void_p * 9)()
    assert lib.gwf_unpack_gather(C.c_void_p(8), 4, 11, None, 4, slots, 9, None) == -1 and b'8 peer' in lib.gwf_last_error()
    assert lib.gwf_pattern(C.byref(det), 0., None, None, None, None, 4, None, None, None, None) == -1
    so = K.gwf_signal_out(None, None, None, None, None, None, None)
    assert lib.gwf_signal_grid(C.byref(m), C.byref(det), 0., None, 4, None, 10, 0, C.byref(so), None, 0, None) == -1
    assert lib.gwf_peer_open(None, None) == -1 and lib.gwf_peer_alloc(0, None, None) == -1
    assert C.sizeof(K.gwf_fisher_out) == 7 * 8 and C.sizeof(K.gwf_signal_out) == 7 * 8


def test_launch_groups_are_cut_at_multiples_of_the_persistent_round():
    from gwfast_b200 import _engine
    for m in (1, 591, 592, 1000, 2500, 10000, 65536):
        for rnd in (592, 1184):
            g = _engine._round_groups(m, rnd)
            assert g[0][0] == 0 and sum(x[1] for x in g) == m and all(a[0] + a[1] == b[0] for a, b in zip(g, g[1:]))
            assert all(lo % rnd == 0 for lo, _ in g) and len(g) <= _engine.MAX_GROUPS
            assert all(x[1] >= 2 * rnd or len(g) == 1 or x is g[-1] for x in g)        # every group but the last holds whole rounds, at least two
    assert _engine._round_groups(10000, 592) == [(0, 2960), (2960, 2368), (5328, 2368), (7696, 2304)]
    assert _engine._round_groups(1000, 592) == [(0, 1000)]
