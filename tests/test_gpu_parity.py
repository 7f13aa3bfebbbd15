"""Parity tests proper: the CUDA path, called through the gwfast-style API -> ctypes -> C ABI, against
(i) the reference's own outputs (tests/golden, generated from the unmodified reference),
(ii) the numpy port oracle on fresh seeded inputs,
(iii) size-independent properties at BASELINE.json's full configuration size.
Tolerances are the north star's: SNR 1e-9 relative, Fisher 1e-6 relative to sqrt(F_ii F_jj); FP64 throughout."""
import numpy as np
import pytest

from conftest import load_golden, make_network, copy_events, fisher_err, snr_err, SNR_RTOL, FISHER_TOL

pytestmark = pytest.mark.gpu

GOLDEN = ['c1_tf2_bns_etsl', 'c1b_tf2tidal_et', 'c1c_tf2_options_etsl', 'c2_phenomd_et2ce', 'var_m1m2_chisa', 'var_lin_res400_fmax',
          'var_fref_nocut', 'var_tf2_m1m2']


@pytest.fixture(scope='module', autouse=True)
def _engine_loaded():
    import torch
    assert torch.cuda.is_available(), 'these tests need the B200'
    from gwfast_b200 import _engine
    _engine.state()          # raises if libgwfast_b200.so is missing: no silent fallback


@pytest.mark.parametrize('name', GOLDEN)
def test_engine_matches_reference_golden(name):
    cfg, ev, out = load_golden(name)
    net = make_network('engine', cfg)
    res = cfg.get('res', 1000)
    assert snr_err(net.SNR(copy_events(ev), res=res), out['snr']) < SNR_RTOL
    F = net.FisherMatr(copy_events(ev), res=res, **cfg.get('fisher_kw', {}))
    assert F.shape == out['fisher'].shape
    assert fisher_err(F, out['fisher']) < FISHER_TOL


@pytest.mark.parametrize('name', ['c2_phenomd_et2ce', 'c1b_tf2tidal_et', 'var_tf2_m1m2'])
def test_engine_return_all_matches_reference(name):
    cfg, ev, out = load_golden(name)
    net = make_network('engine', cfg)
    fa = net.FisherMatr(copy_events(ev), return_all=True, **cfg.get('fisher_kw', {}))
    sa = net.SNR(copy_events(ev), return_all=True)
    keys = [k[8:] for k in out if k.startswith('fisher__')]
    assert set(fa) == set(keys) and set(sa) == set(keys)
    for k in keys:
        assert fisher_err(fa[k], out['fisher__' + k]) < FISHER_TOL, k
        assert snr_err(sa[k], out['snr__' + k]) < SNR_RTOL, k


def test_single_detector_api_on_reference_warmup_event():
    """GWSignal.SNRInteg / FisherMatr (L and T, rotation on/off) on gwfast/signal.py:181-203's event."""
    from gwfast_b200 import waveforms, signal, gwfastGlobals as glob
    import os
    cfg, ev, out = load_golden('init_event')
    psd = os.path.join(glob.detPath, cfg['psd'])
    for cls in ('TaylorF2_RestrictedPN', 'IMRPhenomD'):
        for shape in 'LT':
            for rot in (0, 1):
                s = signal.GWSignal(getattr(waveforms, cls)(), psd_path=psd, detector_shape=shape, det_lat=cfg['det_lat'], det_long=cfg['det_long'],
                                    det_xax=cfg['det_xax'], verbose=False, useEarthMotion=bool(rot), fmin=cfg['fmin'])
                key = '%s__%s__%d' % (cls, shape, rot)
                e = copy_events(ev)
                assert snr_err(s.SNRInteg(e), out['snr__' + key]) < SNR_RTOL, key
                assert fisher_err(s.FisherMatr(e), out['fisher__' + key]) < FISHER_TOL, key
                if shape == 'T':
                    per_arm = s.SNRInteg(copy_events(ev), return_all=True)
                    assert per_arm.shape == (3, 1)
                    assert abs(np.sqrt((per_arm ** 2).sum()) / out['snr__' + key][0] - 1) < SNR_RTOL
                    Fl = s.FisherMatr(copy_events(ev), return_all=True)
                    assert isinstance(Fl, list) and len(Fl) == 3 and fisher_err(sum(Fl), out['fisher__' + key]) < FISHER_TOL


@pytest.mark.parametrize('model,cat,netname,rot', [('IMRPhenomD', 'bbh', 'ET+2CE', True), ('IMRPhenomD', 'bbh', 'LVK-O4', False),
                                                   ('TaylorF2_RestrictedPN', 'bns', 'ET', True)])
def test_engine_matches_port_on_seeded_catalog(model, cat, netname, rot):
    from gwfast_b200 import synthetic
    ev = synthetic.bbh_catalog(48, 4242) if cat == 'bbh' else synthetic.bns_catalog(48, 4243)
    cfg = dict(model=dict(cls=model), network=netname, rot=rot, fmin=10. if netname == 'LVK-O4' else 2.)
    eng, port = make_network('engine', cfg), make_network('port', cfg)
    assert snr_err(eng.SNR(copy_events(ev)), port.SNR(copy_events(ev))) < SNR_RTOL
    assert fisher_err(eng.FisherMatr(copy_events(ev)), port.FisherMatr(copy_events(ev))) < FISHER_TOL


def test_properties_at_full_config_size():
    """BASELINE.json configs[1] at full size (10^4 events, ET+2CE, IMRPhenomD, Earth rotation on)."""
    from gwfast_b200 import synthetic
    cfg = dict(model=dict(cls='IMRPhenomD'), network='ET+2CE', rot=True, fmin=2.)
    net = make_network('engine', cfg)
    ev = synthetic.bbh_catalog(10000, synthetic.SEEDS['C2'])
    snr = net.SNR(copy_events(ev))
    F = net.FisherMatr(copy_events(ev))
    assert snr.shape == (10000,) and F.shape == (11, 11, 10000)
    assert np.all(np.isfinite(snr)) and np.all(np.isfinite(F))
    # the reference's own self-check, signal.py:1576: d h/d dL = -h/dL  =>  F[dL,dL] dL^2 = SNR^2
    assert np.max(np.abs(F[2, 2] * ev['dL'] ** 2 / snr ** 2 - 1)) < 1e-12
    # d h/d Phicoal = -i h  =>  F[Phicoal,Phicoal] = SNR^2 and F[dL,Phicoal] = 0
    assert np.max(np.abs(F[8, 8] / snr ** 2 - 1)) < 1e-12
    assert np.max(np.abs(F[2, 8]) / np.sqrt(F[2, 2] * F[8, 8])) < 1e-12
    assert np.array_equal(F, F.transpose(1, 0, 2))
    # Gram matrices are positive semi-definite: all 2x2 principal minors non-negative
    dg = np.einsum('iin->in', F)
    assert np.all(dg > 0)
    assert np.all(F ** 2 <= dg[:, None, :] * dg[None, :, :] * (1 + 1e-12))
    # the first 64 events are the golden fixture's
    _, _, out = load_golden('c2_phenomd_et2ce')
    assert snr_err(snr[:64], out['snr']) < SNR_RTOL and fisher_err(F[..., :64], out['fisher']) < FISHER_TOL
    # events are independent: results do not depend on the batch they are computed in (bitwise)
    sub = {k: v[1234:1300] for k, v in ev.items()}
    assert np.array_equal(net.FisherMatr(copy_events(sub)), F[..., 1234:1300])
    assert np.array_equal(net.SNR(copy_events(sub)), snr[1234:1300])


def test_network_sum_rules():
    """sum over arms/detectors of per-arm Fishers = network Fisher; triangle 2-Gram form = explicit 3 arms."""
    from gwfast_b200 import synthetic
    cfg = dict(model=dict(cls='IMRPhenomD'), network='ET+2CE', rot=True, fmin=2.)
    net = make_network('engine', cfg)
    ev = synthetic.bbh_catalog(200, 77)
    F = net.FisherMatr(copy_events(ev))
    fa = net.FisherMatr(copy_events(ev), return_all=True)
    total = sum(v for k, v in fa.items() if k != 'net')
    assert fisher_err(total, F) < 1e-12 and fisher_err(fa['net'], F) < 1e-12
    sa = net.SNR(copy_events(ev), return_all=True)
    assert snr_err(np.sqrt(sum(v ** 2 for k, v in sa.items() if k != 'net')), net.SNR(copy_events(ev))) < 1e-13
    # each detector alone, through GWSignal, equals its entry of the fused network launch
    for d, s in net.signals.items():
        Fd = s.FisherMatr(copy_events(ev))
        want = fa[d] if s.detector_shape == 'L' else fa[d + '_0'] + fa[d + '_1'] + fa[d + '_2']
        assert fisher_err(Fd, want) < 1e-12


def test_reparametrisation_is_a_congruence():
    """Fisher in (m1, m2, chiS, chiA) = J^T F J with the analytic Jacobian of (Mc, eta, chi1z, chi2z)."""
    from gwfast_b200 import synthetic
    cfg = dict(model=dict(cls='IMRPhenomD'), network='ET', rot=True, fmin=2.)
    net = make_network('engine', cfg)
    ev = synthetic.bbh_catalog(32, 5)
    F = net.FisherMatr(copy_events(ev))
    F2 = net.FisherMatr(copy_events(ev), use_m1m2=True, use_chi1chi2=False)
    Mc, eta = ev['Mc'], ev['eta']
    M = Mc / eta ** 0.6
    sq = np.sqrt(1 - 4 * eta)
    m1, m2 = 0.5 * M * (1 + sq), 0.5 * M * (1 - sq)
    J = np.zeros((11, 11, len(Mc)))
    for i in range(11):
        J[i, i] = 1.
    J[0, 0], J[0, 1] = Mc * (0.6 / m1 - 0.2 / M), Mc * (0.6 / m2 - 0.2 / M)          # dMc/dm1, dMc/dm2
    J[1, 0], J[1, 1] = eta * (1 / m1 - 2 / M), eta * (1 / m2 - 2 / M)                  # deta/dm1, deta/dm2
    J[9, 9], J[9, 10], J[10, 9], J[10, 10] = 1., 1., 1., -1.                          # chi1 = chiS + chiA, chi2 = chiS - chiA
    want = np.einsum('ian,ijn,jbn->abn', J, F, J)
    assert fisher_err(F2, want) < 1e-9


def test_edge_cases():
    from gwfast_b200 import synthetic, waveforms, signal, network
    cfg = dict(model=dict(cls='TaylorF2_RestrictedPN'), network='ETSL', rot=True, fmin=2.)
    eng, port = make_network('engine', cfg), make_network('port', cfg)
    one = synthetic.bns_catalog(1, 3)
    assert eng.SNR(copy_events(one)).shape == (1,) and eng.FisherMatr(copy_events(one)).shape == (11, 11, 1)
    # ragged: resolutions that are not a multiple of the warp width, minimum grid, non-contiguous inputs
    ev = synthetic.bns_catalog(33, 4)
    for res in (2, 31, 33, 257):
        assert snr_err(eng.SNR(copy_events(ev), res=res), port.SNR(copy_events(ev), res=res)) < SNR_RTOL, res
        assert fisher_err(eng.FisherMatr(copy_events(ev), res=res), port.FisherMatr(copy_events(ev), res=res)) < FISHER_TOL, res
    big = synthetic.bns_catalog(66, 4)
    strided = {k: v[::2] for k, v in big.items()}
    dense = {k: np.ascontiguousarray(v) for k, v in strided.items()}
    assert np.array_equal(eng.FisherMatr(strided), eng.FisherMatr(dense))
    # PSD = 1 outside the table (signal.py:723): CE tables start at 5 Hz, fmin = 2 Hz -- covered by ET+2CE goldens; fmax below fmin of table
    # df instead of res (signal.py:890-892)
    s = synthetic.build_network(signal.GWSignal, waveforms.TaylorF2_RestrictedPN(), 'ETSL', fmin=10.)['ETSL']
    from oracle.port import waveforms as PW, detector as PD
    p = synthetic.build_network(PD.Detector, PW.TaylorF2_RestrictedPN(), 'ETSL', fmin=10.)['ETSL']
    fcut = PW.TaylorF2_RestrictedPN().fcut(**ev)
    res_df = int(np.amax(np.floor(1 + (fcut - 10.) / 0.5)))
    assert fisher_err(s.FisherMatr(copy_events(ev), res=None, df=0.5, spacing='lin'), p.FisherMatr(copy_events(ev), res=res_df, spacing='lin')) < FISHER_TOL
    with pytest.raises(ValueError, match='resolution in frequency or step size'):
        s.FisherMatr(copy_events(ev), res=None)
    # the events dict is filled in place like the reference does
    e2 = copy_events(ev)
    e2['chiS'], e2['chiA'] = 0.5 * (e2['chi1z'] + e2['chi2z']), 0.5 * (e2['chi1z'] - e2['chi2z'])
    del e2['chi1z'], e2['chi2z']
    s.SNRInteg(e2)
    assert 'chi1z' in e2 and 'chi2z' in e2


def test_duty_factor_masks_follow_numpy_rng():
    """signal.py:672-673, 729-731: Bernoulli masks from the global numpy RNG seeded with seedUse, one draw per arm."""
    from gwfast_b200 import synthetic, waveforms, signal
    ev = synthetic.bbh_catalog(64, 8)
    s = synthetic.build_network(signal.GWSignal, waveforms.IMRPhenomD(), 'ET', DutyFactor=0.6)['ET']
    s._update_seed(seed=1234)
    full = synthetic.build_network(signal.GWSignal, waveforms.IMRPhenomD(), 'ET')['ET']
    arms = full.SNRInteg(copy_events(ev), return_all=True)
    np.random.seed(1234)
    masks = np.array([np.random.choice([0, 1], 64, p=[0.4, 0.6]) for _ in range(3)])
    assert np.allclose(s.SNRInteg(copy_events(ev)), np.sqrt(((arms * masks) ** 2).sum(axis=0)), rtol=1e-13)
    Fl = full.FisherMatr(copy_events(ev), return_all=True)
    want = sum(Fl[i] * masks[i] for i in range(3))
    assert np.allclose(s.FisherMatr(copy_events(ev)), want, rtol=1e-12, atol=0.)
