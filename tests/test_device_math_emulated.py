"""The device formulas (gwfast_b200/csrc/*.cuh are __host__ __device__) driven by a CPU loop (tests/emu) and compared
with the reference outputs in tests/golden.  This checks the math without a GPU; thread mapping, shared-memory staging
and warp reductions are covered by the -m gpu tests.  The emulation library is test infrastructure: the package never
loads it."""
import shutil

import numpy as np
import pytest

from conftest import load_golden, fisher_err, snr_err, SNR_RTOL, FISHER_TOL

pytestmark = pytest.mark.skipif(shutil.which('nvcc') is None, reason='nvcc needed to build the emulation harness')


def _emu_inputs(cfg):
    from gwfast_b200 import synthetic, waveforms, signal, _capi as K
    kw = {}
    if cfg.get('fmax') is not None:
        kw['fmax'] = cfg['fmax']
    model = getattr(waveforms, cfg['model']['cls'])(**cfg['model'].get('kw', {}))
    sigs = synthetic.build_network(signal.GWSignal, model, cfg['network'], useEarthMotion=cfg['rot'], fmin=cfg['fmin'], **kw)
    dets = [s._detector_struct(i) for i, s in enumerate(sigs.values())]
    psds = [(s.strainFreq, s.noiseCurve) for s in sigs.values()]
    return model, dets, psds


@pytest.mark.parametrize('name', ['c1_tf2_bns_etsl', 'c1b_tf2tidal_et', 'c1c_tf2_options_etsl', 'c2_phenomd_et2ce', 'var_m1m2_chisa',
                                  'var_lin_res400_fmax', 'var_fref_nocut', 'var_tf2_m1m2', 'tf2ecc_etsl', 'tf2ecc_fref_et', 'tf2ecc_tidal_lvk'])
def test_emulated_device_math_matches_reference(name):
    import emu_driver as E
    from gwfast_b200 import _capi as K
    cfg, ev, out = load_golden(name)
    model, dets, psds = _emu_inputs(cfg)
    n = min(len(ev['Mc']), 24)
    sub = {k: v[:n] for k, v in ev.items()}
    fkw = cfg.get('fisher_kw', {})
    flags = (K.GWF_OPT_M1M2 if fkw.get('use_m1m2') else 0) | (0 if fkw.get('use_chi1chi2', True) else K.GWF_OPT_CHIS_CHIA) | \
            (K.GWF_OPT_LIN_GRID if fkw.get('spacing') == 'lin' else 0)
    res = cfg.get('res', 1000)
    from gwfast_b200 import signal
    packed, s2 = E.run(model._descriptor(sub), dets, psds, signal._engine_events(model, sub, None, bool(fkw.get('use_m1m2'))), res=res, flags=flags)
    F = E.unpack(packed, model.nParams)[0]
    assert fisher_err(F, out['fisher'][..., :n]) < FISHER_TOL
    arms, _ = E.run(model._descriptor(sub), dets, psds, signal._engine_events(model, sub), res=res, snr_mode=True)
    assert snr_err(np.sqrt(arms.sum(axis=0)), out['snr'][:n]) < SNR_RTOL
    # tighter than the gate: what the formulation actually achieves
    assert fisher_err(F, out['fisher'][..., :n]) < 5e-9


def test_emulated_nrtidal_matches_masked_reference():
    """IMRPhenomD_NRTidalv2 (13 parameters, custom-JVP taper): against the reference with its rounding-fragile last grid
    sample forced to zero, and within the size of that artefact against the raw reference."""
    import emu_driver as E
    cfg, ev, out = load_golden('c3_nrtidal_et2ce')
    model, dets, psds = _emu_inputs(cfg)
    from gwfast_b200 import signal
    ev = signal._engine_events(model, ev)
    packed, _ = E.run(model._descriptor(ev), dets, psds, ev)
    F = E.unpack(packed, 13)[0]
    arms, _ = E.run(model._descriptor(ev), dets, psds, ev, snr_mode=True)
    snr = np.sqrt(arms.sum(axis=0))
    assert snr_err(snr, out['snr_masked']) < SNR_RTOL and fisher_err(F, out['fisher_masked']) < FISHER_TOL
    assert snr_err(snr, out['snr']) < 1e-5 and fisher_err(F, out['fisher']) < 5e-3


@pytest.mark.parametrize('name', ['newt_et', 'newt_lvk_m1m2'])
def test_emulated_newtinspiral_matches_reference(name):
    """NewtInspiral = the TaylorF2 point functions with GWF_MODEL_NEWTONIAN; the engine's 11-parameter layout reduced to the model's 8."""
    import emu_driver as E
    from gwfast_b200 import signal, _capi as K
    cfg, ev, out = load_golden(name)
    model, dets, psds = _emu_inputs(cfg)
    flags = K.GWF_OPT_M1M2 if cfg.get('fisher_kw', {}).get('use_m1m2') else 0
    e2 = signal._engine_events(model, ev, None, bool(flags))
    packed, _ = E.run(model._descriptor(ev), dets, psds, e2, flags=flags)
    rows = model._engine_rows
    F = E.unpack(packed, 11)[0][rows][:, rows]
    assert fisher_err(F, out['fisher']) < FISHER_TOL
    # the dropped eta / spin rows vanish up to the rounding of M = Mc/eta^(3/5) cancelling against Mc (1e-14 of the Mc row)
    # (with use_m1m2 slot 1 is m2, which does enter through Mc; the reference differentiates w.r.t. m1 only, signal.py:1147)
    dropped = [9, 10] if flags else [1, 9, 10]
    assert np.max(np.abs(E.unpack(packed, 11)[0][dropped])) < 1e-10 * np.max(np.abs(F))
    arms, _ = E.run(model._descriptor(ev), dets, psds, signal._engine_events(model, ev), snr_mode=True)
    assert snr_err(np.sqrt(arms.sum(axis=0)), out['snr']) < SNR_RTOL


def test_emulated_nrtidal_taper_tangent_does_not_overflow():
    """Event 5487 of the C3 catalog has a grid sample 5e-6 (in Mf) above f_merger: exp(...) ~ 1e300 in the Planck taper tangent
    (waveforms.py:1714); the reference's inf/inf -> nan_to_num -> 0 must come out as a finite ~0 here, not inf*0 = NaN."""
    import emu_driver as E
    from conftest import make_network, copy_events
    from gwfast_b200 import synthetic, signal
    cfg = dict(model=dict(cls='IMRPhenomD_NRTidalv2'), network='ET+2CE', rot=True, fmin=2.)
    ev = synthetic.bns_catalog(10000, synthetic.SEEDS['C3'], tidal=True)
    sub = {k: v[[5487]] for k, v in ev.items()}
    model, dets, psds = _emu_inputs(cfg)
    e2 = signal._engine_events(model, sub)
    packed, _ = E.run(model._descriptor(e2), dets, psds, e2)
    F = E.unpack(packed, 13)[0]
    assert np.all(np.isfinite(F))
    port = make_network('port', dict(cfg, model=dict(cls='IMRPhenomD_NRTidalv2', kw=dict(taper_end_zero=True))))
    assert fisher_err(F, port.FisherMatr(copy_events(sub))) < FISHER_TOL


@pytest.mark.parametrize('name', ['c4_phenomhm_lvk', 'c4b_phenomhm_et'])
def test_emulated_phenomhm_matches_reference(name):
    """IMRPhenomHM: six modes, complex mode sum, iota differentiated through the spin-weighted harmonics; the SNR keeps the
    reference's definition (|hp| Fp, |hc| Fc without the cross term)."""
    import emu_driver as E
    cfg, ev, out = load_golden(name)
    model, dets, psds = _emu_inputs(cfg)
    from gwfast_b200 import signal
    ev = signal._engine_events(model, ev)
    packed, _ = E.run(model._descriptor(ev), dets, psds, ev)
    F = E.unpack(packed, 11)[0]
    arms, _ = E.run(model._descriptor(ev), dets, psds, ev, snr_mode=True)
    assert snr_err(np.sqrt(arms.sum(axis=0)), out['snr']) < SNR_RTOL
    assert fisher_err(F, out['fisher']) < FISHER_TOL
    # the HM SNR is NOT (h|h)^(1/2): F[dL,dL] dL^2 / SNR^2 deviates from 1 at the 1e-3 level by design
    r = F[2, 2] * ev['dL'] ** 2 / out['snr'] ** 2
    assert np.all(np.abs(r - 1) < 0.05) and np.any(np.abs(r - 1) > 1e-6)


def test_emulated_waveform_values_match_reference(has_reference):
    """WaveFormModel.Phi / Ampl / tau_star (+ IMRPhenomHM.hphc) on a user grid against the reference's own methods (container only).
    The very last sample (Mf == fcutPar up to rounding) is excluded: the reference itself returns 0 or the MRD value there
    depending on the last bit of M*GMsun_over_c3*f."""
    if not has_reference:
        pytest.skip('reference tree not mounted')
    import warnings
    warnings.filterwarnings('ignore')
    import emu_driver as E
    from oracle import reference
    from gwfast_b200 import waveforms as W, synthetic
    wf, sig, net, utils, glob = reference.load()
    for cls, cat in (('TaylorF2_RestrictedPN', synthetic.bbh_catalog(6, 3)), ('IMRPhenomD', synthetic.bbh_catalog(6, 3)),
                     ('IMRPhenomD_NRTidalv2', synthetic.bns_catalog(6, 3, tidal=True)), ('IMRPhenomHM', synthetic.bbh_catalog(6, 3))):
        rm, m = getattr(wf, cls)(), getattr(W, cls)()
        fg = np.geomspace(np.full(6, 5.), 0.97 * rm.fcut(**cat), 200)
        ev = dict(cat)
        if cls != 'TaylorF2_RestrictedPN':
            ev['_Mtot_sec'] = (cat['Mc'] / (cat['eta'] ** (3. / 5.))) * glob.GMsun_over_c3
        out = E.waveform(m._descriptor(cat), ev, fg, want=('phi', 'ampl', 'tau') + (('hphc',) if cls == 'IMRPhenomHM' else ()))
        P, A, T = rm.Phi(fg, **cat), rm.Ampl(fg, **cat), rm.tau_star(fg, **cat)
        assert np.max(np.abs(out['tau'] / T - 1)) < 1e-9
        assert np.max(np.abs(out['fcut'] / rm.fcut(**cat) - 1)) < 1e-14
        if cls == 'IMRPhenomHM':
            for i, k in enumerate(('21', '22', '32', '33', '43', '44')):
                assert np.max(np.abs(out['phi'][i] - P[k])) < 1e-8, k
                assert np.max(np.abs(out['ampl'][i] - A[k]) / np.max(np.abs(A[k]), axis=0)) < 1e-11, k
            hp, hc = rm.hphc(fg, **cat)
            assert np.max(np.abs(out['hphc'][0] + 1j * out['hphc'][1] - hp) / np.max(np.abs(hp), axis=0)) < 1e-10
            assert np.max(np.abs(out['hphc'][2] + 1j * out['hphc'][3] - hc) / np.max(np.abs(hc), axis=0)) < 1e-10
        else:
            assert np.max(np.abs(out['phi'][0] - P)) < 1e-8, cls
            assert np.max(np.abs(out['ampl'][0] - A) / np.max(np.abs(A), axis=0)) < 1e-11, cls


@pytest.mark.parametrize('name', ['nsbh_et2ce', 'nsbh_lvk_lin_fmax', 'nsbh_et_m1m2_chisa_fref_nocut', 'edge_nsbh_et2ce'])
def test_emulated_nsbh_matches_reference(name):
    """IMRPhenomNSBH (13 parameters): PhenomD phase with the NSBH remnant, t0 at the last grid sample of every group, Pade tidal phase,
    the IMRPhenomC-style amplitude differentiated per sample in dual arithmetic, and xi_tide from the engine's own root finder
    (no table: the emulation evaluates the eight nodes of an event's cell on the spot) -- against the unmodified reference, whose
    table was tabulated by its own numpy.roots loop (oracle/nsbh_table.py)."""
    import emu_driver as E
    from gwfast_b200 import signal, _capi as K
    cfg, ev, out = load_golden(name)
    model, dets, psds = _emu_inputs(cfg)
    n = min(len(ev['Mc']), 20)
    sub = {k: v[:n] for k, v in ev.items()}
    fkw = cfg.get('fisher_kw', {})
    flags = (K.GWF_OPT_M1M2 if fkw.get('use_m1m2') else 0) | (0 if fkw.get('use_chi1chi2', True) else K.GWF_OPT_CHIS_CHIA) | \
            (K.GWF_OPT_LIN_GRID if fkw.get('spacing') == 'lin' else 0)
    res = cfg.get('res', 1000)
    packed, _ = E.run(model._descriptor(sub), dets, psds, signal._engine_events(model, sub, None, bool(fkw.get('use_m1m2'))), res=res, flags=flags)
    F = E.unpack(packed, 13)[0]
    assert np.all(np.isfinite(F))
    arms, _ = E.run(model._descriptor(sub), dets, psds, signal._engine_events(model, sub), res=res, snr_mode=True)
    assert snr_err(np.sqrt(arms.sum(axis=0)), out['snr'][:n]) < SNR_RTOL
    assert fisher_err(F, out['fisher'][..., :n]) < FISHER_TOL


def test_emulated_nsbh_waveform_values():
    """IMRPhenomNSBH.Phi / Ampl / tau_star / fcut on user grids (inside the cut, and running past it: zeros beyond, t0 at the grid's
    own maximum)."""
    import emu_driver as E
    from gwfast_b200 import waveforms as W
    cfg, ev, out = load_golden('wf_values_nsbh')
    m = W.IMRPhenomNSBH(verbose=False)
    for tag in ('', '_over'):
        fg = out['f' + tag]
        got = E.waveform(m._descriptor(ev), ev, fg, want=('phi', 'ampl', 'tau'))
        inside = np.abs(fg / m.fcut(**ev) - 1.) > 1e-9            # a sample ON the cut goes either way in the reference
        ph, am = out['phi' + tag], out['ampl' + tag]
        assert np.max((np.abs(got['phi'][0] - ph) / (1. + np.abs(ph)))[inside]) < 1e-9
        assert np.max((np.abs(got['ampl'][0] - am) / np.max(np.abs(am), axis=0))[inside]) < 1e-10
        if tag == '':
            assert np.max(np.abs(got['tau'] / out['tau'] - 1)) < 1e-12
    assert np.allclose(m.fcut(**ev), out['fcut'], rtol=1e-14)
    assert np.allclose(got['fcut'], out['fcut'], rtol=1e-13)
