"""Generate ``tests/golden/*.npz`` by running the UNMODIFIED reference (``/root/reference/gwfast``) under the
numpy/dual shim (TEST INFRASTRUCTURE; container only -- the reference tree does not exist on the GPU box).

    python -m oracle.make_golden            # all fixtures
    python -m oracle.make_golden c2 init    # some

Every fixture stores its inputs (events, configuration as a JSON string) next to the reference outputs, so the
tests rebuild the same detector network with the engine (or the port) and compare.
"""
import json
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import reference  # noqa: E402
from gwfast_b200 import synthetic  # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')
REF_PSDS = os.path.join(reference.REFERENCE_ROOT, 'psds')

# the reference's own warm-up event, gwfast/signal.py:181-203
INIT_EVENT = {'Mc': [77.23905294], 'Phicoal': [3.28297867], 'chi1z': [0.2018924], 'chi2z': [-0.68859213], 'dL': [22.68426174],
              'eta': [0.20586622], 'iota': [4.48411048], 'phi': [0.90252645], 'psi': [3.11843169], 'Lambda1': [300.], 'Lambda2': [300.],
              'tcoal': [0.], 'theta': [3.00702251]}
# GW170817 as printed in notebooks/gwfast_tutorial.ipynb cell 10
GW170817 = {'Mc': [1.19752182], 'dL': [0.04374755], 'theta': [1.97888033], 'phi': [3.44616], 'iota': [2.5450656], 'psi': [0.], 'tcoal': [0.43432288],
            'eta': [0.24786618], 'Phicoal': [0.], 'chi1z': [0.00513614], 'chi2z': [0.00323515], 'Lambda1': [368.17802384], 'Lambda2': [586.54870315]}


def _model(wf, spec):
    return getattr(wf, spec['cls'])(**spec.get('kw', {}))


def _copy(ev):
    return {k: np.array(v, dtype=float) for k, v in ev.items()}


def run_network(cfg, ev, want_all=True):
    """cfg: dict(model=dict(cls, kw), network=name, rot, fmin, fmax, fisher_kw, res)."""
    wf, sig, net, utils, glob = reference.load()
    kw = {}
    if cfg.get('fmax') is not None:
        kw['fmax'] = cfg['fmax']
    sigs = synthetic.build_network(sig.GWSignal, _model(wf, cfg['model']), cfg['network'], useEarthMotion=cfg['rot'], fmin=cfg['fmin'],
                                   psd_root=REF_PSDS, **kw)
    N = net.DetNet(sigs, verbose=False)
    res = cfg.get('res', 1000)
    fkw = dict(cfg.get('fisher_kw', {}))
    out = {'snr': N.SNR(_copy(ev), res=res), 'fisher': N.FisherMatr(_copy(ev), res=res, **fkw)}
    if want_all:
        sa = N.SNR(_copy(ev), res=res, return_all=True)
        fa = N.FisherMatr(_copy(ev), res=res, return_all=True, **fkw)
        for k in sa:
            out['snr__' + k] = sa[k]
            out['fisher__' + k] = fa[k]
    return out


def save(name, cfg, ev, out, extra=None):
    os.makedirs(GOLD, exist_ok=True)
    data = {'config': np.array(json.dumps(cfg))}
    data.update({'ev__' + k: np.asarray(v, dtype=float) for k, v in ev.items()})
    data.update(out)
    if extra:
        data.update(extra)
    path = os.path.join(GOLD, name + '.npz')
    np.savez_compressed(path, **data)
    print('wrote %s (%.1f KB)' % (path, os.path.getsize(path) / 1024.))


def take(ev, n):
    return {k: v[:n] for k, v in ev.items()}


def fx_init():
    """the 16-row table of SURVEY.md App. C: 4 models x L/T x rot on the reference's warm-up event."""
    wf, sig, net, utils, glob = reference.load()
    ev = {k: np.array(v) for k, v in INIT_EVENT.items()}
    out = {}
    psd = os.path.join(REF_PSDS, 'ET-0000A-18.txt')
    for cls in ('TaylorF2_RestrictedPN', 'IMRPhenomD', 'IMRPhenomD_NRTidalv2', 'IMRPhenomHM'):
        for shape in 'LT':
            for rot in (0, 1):
                s = sig.GWSignal(getattr(wf, cls)(), psd_path=psd, detector_shape=shape, det_lat=40.44, det_long=9.45, det_xax=0., verbose=False,
                                 useEarthMotion=bool(rot), fmin=2.)
                key = '%s__%s__%d' % (cls, shape, rot)
                out['snr__' + key] = s.SNRInteg(_copy(ev), res=1000)
                out['fisher__' + key] = s.FisherMatr(_copy(ev), res=1000)
    save('init_event', dict(det_lat=40.44, det_long=9.45, det_xax=0., fmin=2., res=1000, psd='ET-0000A-18.txt'), ev, out)


def fx_c1():
    cfg = dict(model=dict(cls='TaylorF2_RestrictedPN'), network='ETSL', rot=True, fmin=2.)
    ev = synthetic.bns_catalog(100, synthetic.SEEDS['C1'])
    save('c1_tf2_bns_etsl', cfg, ev, run_network(cfg, ev))


def fx_c1b():
    cfg = dict(model=dict(cls='TaylorF2_RestrictedPN', kw=dict(is_tidal=True, use_3p5PN_SpinHO=True)), network='ET', rot=True, fmin=2.)
    ev = take(synthetic.bns_catalog(100, synthetic.SEEDS['C1'] + 100, tidal=True), 32)
    save('c1b_tf2tidal_et', cfg, ev, run_network(cfg, ev))
    cfg = dict(model=dict(cls='TaylorF2_RestrictedPN', kw=dict(is_tidal=True, use_QuadMonTid=True, phiref_vlso=True, which_ISCO='Kerr')),
               network='ETSL', rot=True, fmin=5.)
    save('c1c_tf2_options_etsl', cfg, take(ev, 16), run_network(cfg, take(ev, 16)))


def fx_c2():
    cfg = dict(model=dict(cls='IMRPhenomD'), network='ET+2CE', rot=True, fmin=2.)
    ev = take(synthetic.bbh_catalog(10000, synthetic.SEEDS['C2']), 64)
    save('c2_phenomd_et2ce', cfg, ev, run_network(cfg, ev))


def fx_variants():
    ev = take(synthetic.bbh_catalog(10000, synthetic.SEEDS['C2']), 16)
    for tag, cfg in (
        ('m1m2_chisa', dict(model=dict(cls='IMRPhenomD', kw=dict(is_chi1chi2=False)), network='ET', rot=True, fmin=2.,
                            fisher_kw=dict(use_m1m2=True, use_chi1chi2=False))),
        ('lin_res400_fmax', dict(model=dict(cls='IMRPhenomD'), network='ET+2CE', rot=False, fmin=5., fmax=256., res=400, fisher_kw=dict(spacing='lin'))),
        ('fref_nocut', dict(model=dict(cls='IMRPhenomD', kw=dict(fRef=20., apply_fcut=False)), network='ETSL', rot=False, fmin=2.)),
        ('tf2_m1m2', dict(model=dict(cls='TaylorF2_RestrictedPN', kw=dict(use_3p5PN_SpinHO=True)), network='LVK-O4', rot=False, fmin=10.,
                          fisher_kw=dict(use_m1m2=True))),
    ):
        save('var_' + tag, cfg, ev, run_network(cfg, ev))


def newt_reference(cfg, ev, use_m1m2=False):
    """NewtInspiral through the reference's own functions.  SNR: DetNet.SNR unmodified.  Fisher: the reference's FisherMatr cannot
    run for this model without numdifftools -- `derivargs = (1)` is an int and `derivargs[:-2]` raises (signal.py:1147, 1166) --
    so the derivative strain is taken from the reference's UNMODIFIED GWstrain (signal.py:486-655) differentiated w.r.t. its 8
    parameters with the oracle's forward-mode duals (what computeAnalyticalDeriv=False intends, signal.py:1150), and contracted with
    the reference's trapezoid rule (signal.py:917-931), tcoal row per second (:920)."""
    wf, sig, net, utils, glob = reference.load()
    import jax          # the oracle shim (oracle/refshim), on sys.path after reference.load()
    # GWSignal.__init__ runs a warm-up FisherMatr (signal.py:198), which raises for NewtInspiral for the reason above: the detectors
    # are constructed with TaylorF2 and the model is swapped afterwards; every method used below is the reference's own
    sigs = synthetic.build_network(sig.GWSignal, wf.TaylorF2_RestrictedPN(), cfg['network'], useEarthMotion=cfg['rot'], fmin=cfg['fmin'], psd_root=REF_PSDS)
    for s in sigs.values():
        s.wf_model = _model(wf, cfg['model'])
    N = net.DetNet(sigs, verbose=False)
    out = {'snr': N.SNR(_copy(ev)), 'fisher': 0.}
    sa = N.SNR(_copy(ev), return_all=True)
    for k in sa:
        out['snr__' + k] = sa[k]
    n = len(ev['Mc'])
    z = np.zeros(n)
    for name, s in sigs.items():
        e = _copy(ev)
        fcut = s.wf_model.fcut(**e)
        fg = np.geomspace(np.full(fcut.shape, s.fmin), fcut, num=1000)
        Sn = np.interp(fg, s.strainFreq, s.noiseCurve, left=1., right=1.)
        m0, m1 = (utils.m1m2_from_Mceta(e['Mc'], e['eta']) if use_m1m2 else (e['Mc'], e['eta']))
        rots = [0.] if s.detector_shape == 'L' else [0., 60.]
        Ds = []
        for r in rots:
            f = lambda fgr, Mc, eta, dL, theta, phi, iota, psi, tcoal, Phicoal: s.GWstrain(
                fgr, Mc, eta, dL, theta, phi, iota, psi, tcoal, Phicoal, e['chi1z'], e['chi2z'], z, z, z, z, z, z, z, rot=r,
                is_m1m2=use_m1m2, is_chi1chi2=True)
            J = jax.vmap(jax.jacrev(f, argnums=(1, 3, 4, 5, 6, 7, 8, 9)))(fg.T, m0, m1, e['dL'], e['theta'], e['phi'], e['iota'], e['psi'], e['tcoal'], e['Phicoal'])
            D = np.array([np.asarray(j) for j in J])          # (8, N, res)
            D[6] /= 86400.
            Ds.append(D)
        if s.detector_shape == 'T':
            Ds.append(-(Ds[0] + Ds[1]))
        for i, D in enumerate(Ds):
            Fa = np.zeros((8, 8, n))
            for a in range(8):
                for b in range(a, 8):
                    Fa[a, b] = Fa[b, a] = 4. * np.trapezoid((np.conj(D[a]) * D[b]).real.T / Sn, fg, axis=0)
            out['fisher__' + (name if s.detector_shape == 'L' else '%s_%d' % (name, i))] = Fa
            out['fisher'] = out['fisher'] + Fa
    return out


def fx_ecc():
    """Eccentric TaylorF2 (waveforms.py:814-845): v0ecc from the grid, from fRef_ecc, and with tidal terms (14 parameters)."""
    rng = np.random.default_rng(77)
    ev = take(synthetic.bns_catalog(100, synthetic.SEEDS['C1'] + 7, tidal=True), 12)
    ev['ecc'] = rng.uniform(0.005, 0.15, 12)
    nt = {k: v for k, v in ev.items() if not k.startswith('Lambda')}
    wf = reference.load()[0]
    for tag, cfg, e in (
        ('ecc_etsl', dict(model=dict(cls='TaylorF2_RestrictedPN', kw=dict(is_eccentric=True)), network='ETSL', rot=True, fmin=2.), nt),
        ('ecc_fref_et', dict(model=dict(cls='TaylorF2_RestrictedPN', kw=dict(is_eccentric=True, fRef_ecc=10., use_3p5PN_SpinHO=True)), network='ET', rot=True,
                             fmin=5.), nt),
        ('ecc_tidal_lvk', dict(model=dict(cls='TaylorF2_RestrictedPN', kw=dict(is_eccentric=True, is_tidal=True)), network='LVK-O4', rot=False, fmin=10.,
                               fisher_kw=dict(use_m1m2=True)), ev),
    ):
        out = run_network(cfg, e)
        m = _model(wf, cfg['model'])
        fg = np.geomspace(np.full(12, 5.), 0.97 * m.fcut(**e), 100)
        out.update(wf_f=fg, wf_phi=m.Phi(fg, **e))
        save('tf2' + tag, cfg, e, out)


def fx_newt():
    """NewtInspiral (8 parameters): triangle with Earth rotation, and the LVK network with (m1, m2) as mass parameters."""
    ev = take(synthetic.bbh_catalog(10000, synthetic.SEEDS['C2']), 16)
    cfg = dict(model=dict(cls='NewtInspiral', kw=dict(is_chi1chi2=False)), network='ET', rot=True, fmin=2.)
    out = newt_reference(cfg, ev)
    # stand-alone methods of the reference class (these do work): Phi, Ampl, tau_star, fcut on a per-event grid
    wf = reference.load()[0]
    m = wf.NewtInspiral(is_chi1chi2=False)
    fg = np.geomspace(np.full(16, 5.), 0.97 * m.fcut(**ev), 120)
    out.update(wf_f=fg, wf_fcut=m.fcut(**ev), wf_tau=m.tau_star(fg, **ev), wf_phi=m.Phi(fg, **ev), wf_ampl=m.Ampl(fg, **ev))
    save('newt_et', cfg, ev, out)
    cfg = dict(model=dict(cls='NewtInspiral', kw=dict(is_chi1chi2=False)), network='LVK-O4', rot=False, fmin=10., fisher_kw=dict(use_m1m2=True))
    save('newt_lvk_m1m2', cfg, ev, newt_reference(cfg, ev, use_m1m2=True))


def masked_last_sample(cfg, ev):
    """Reference outputs with the LAST grid sample of every arm forced to zero, computed from the reference's own
    GWAmplitudes / derivative arrays (signal.py:425, 1094-1098) and its trapezoid rule.  Used for IMRPhenomD_NRTidalv2, whose
    last sample sits on the end of the Planck taper and is 0 or 1 by last-bit rounding."""
    wf, sig, net, utils, glob = reference.load()
    sigs = synthetic.build_network(sig.GWSignal, _model(wf, cfg['model']), cfg['network'], useEarthMotion=cfg['rot'], fmin=cfg['fmin'], psd_root=REF_PSDS)
    snr2 = 0.
    F = 0.
    for s in sigs.values():
        e = _copy(ev)
        utils.check_evparams(e)
        fcut = s.wf_model.fcut(**e)
        fg = np.geomspace(np.full(fcut.shape, s.fmin), fcut, num=1000)
        Sn = np.interp(fg, s.strainFreq, s.noiseCurve, left=1., right=1.)
        rots = [0.] if s.detector_shape == 'L' else [0., 60.]
        amps = [s.GWAmplitudes(e, fg, rot=r) for r in rots]
        if s.detector_shape == 'T':
            amps.append((-(amps[0][0] + amps[1][0]), -(amps[0][1] + amps[1][1])))
        for Ap, Ac in amps:
            y = (Ap * Ap + Ac * Ac) / Sn
            y[-1] = 0.
            snr2 = snr2 + 4. * np.trapezoid(y, fg, axis=0)
        Fs, Ds = s.FisherMatr(_copy(ev), return_derivatives=True)
        # FisherMatr's own grid (fcut evaluated on the dict as FisherMatr sees it, signal.py:884)
        for D in Ds:
            D = np.array(D)
            D[:, :, -1] = 0.
            nP = D.shape[0]
            Fa = np.zeros((nP, nP, D.shape[1]))
            for a in range(nP):
                for b in range(a, nP):
                    Fa[a, b] = Fa[b, a] = 4. * np.trapezoid((np.conj(D[a]) * D[b]).real.T / Sn, fg, axis=0)
            F = F + Fa
    return {'snr_masked': np.sqrt(snr2), 'fisher_masked': F}


def fx_c3():
    cfg = dict(model=dict(cls='IMRPhenomD_NRTidalv2'), network='ET+2CE', rot=True, fmin=2.)
    ev = take(synthetic.bns_catalog(10000, synthetic.SEEDS['C3'], tidal=True), 32)
    out = run_network(cfg, ev)
    out.update(masked_last_sample(cfg, ev))
    save('c3_nrtidal_et2ce', cfg, ev, out)


def fx_c4():
    cfg = dict(model=dict(cls='IMRPhenomHM'), network='LVK-O4', rot=False, fmin=10.)
    ev = take(synthetic.bbh_catalog(100000, synthetic.SEEDS['C4']), 32)
    save('c4_phenomhm_lvk', cfg, ev, run_network(cfg, ev))
    cfg = dict(model=dict(cls='IMRPhenomHM'), network='ET', rot=True, fmin=2.)
    save('c4b_phenomhm_et', cfg, take(ev, 8), run_network(cfg, take(ev, 8)))


def fx_gw170817():
    """notebook known answers: network SNR 33.16 (cell 12), 1 - F[dL,dL]/(SNR/dL)^2 = 2.2e-16 (cell 15), fcut 1590.1/3433.5 Hz."""
    wf, sig, net, utils, glob = reference.load()
    files = {'L1': 'LVC_O1O2O3/2017-08-06_DCH_C02_L1_O2_Sensitivity_strain_asd.txt', 'H1': 'LVC_O1O2O3/2017-06-10_DCH_C02_H1_O2_Sensitivity_strain_asd.txt',
             'Virgo': 'LVC_O1O2O3/Hrec_hoft_V1O2Repro2A_16384Hz.txt'}
    ev = {k: np.array(v) for k, v in GW170817.items()}
    sigs, extra = {}, {}
    for d, rel in files.items():
        s = glob.detectors[d]
        sigs[d] = sig.GWSignal(wf.IMRPhenomD_NRTidalv2(), psd_path=os.path.join(REF_PSDS, rel), detector_shape=s['shape'], det_lat=s['lat'],
                               det_long=s['long'], det_xax=s['xax'], verbose=False, useEarthMotion=False, fmin=10.)
        extra['psd_f__' + d] = sigs[d].strainFreq
        extra['psd_S__' + d] = sigs[d].noiseCurve
    N = net.DetNet(sigs, verbose=False)
    out = {'snr': N.SNR(_copy(ev)), 'fisher': N.FisherMatr(_copy(ev))}
    out['tf2_fcut_schw'] = wf.TaylorF2_RestrictedPN().fcut(**ev)
    out['tf2_fcut_kerr'] = wf.TaylorF2_RestrictedPN(which_ISCO='Kerr').fcut(**ev)
    print('GW170817: SNR %.8f  1-F[dL,dL]dL^2/SNR^2 = %.2e  fcut %.8f %.8f' % (
        out['snr'][0], 1 - out['fisher'][2, 2, 0] * ev['dL'][0] ** 2 / out['snr'][0] ** 2, out['tf2_fcut_schw'][0], out['tf2_fcut_kerr'][0]))
    save('gw170817', dict(model=dict(cls='IMRPhenomD_NRTidalv2'), detectors=list(files), rot=False, fmin=10.), ev, out, extra)


def fx_wfvalues():
    """WaveFormModel.Phi / Ampl / tau_star / fcut (+ IMRPhenomHM.hphc) of the reference on a per-event grid."""
    wf, sig, net, utils, glob = reference.load()
    out = {}
    evs = {}
    for cls, cat in (('TaylorF2_RestrictedPN', synthetic.bbh_catalog(5, 31)), ('IMRPhenomD', synthetic.bbh_catalog(5, 32)),
                     ('IMRPhenomD_NRTidalv2', synthetic.bns_catalog(5, 33, tidal=True)), ('IMRPhenomHM', synthetic.bbh_catalog(5, 34))):
        m = getattr(wf, cls)()
        fg = np.geomspace(np.full(5, 5.), 0.97 * m.fcut(**cat), 160)
        out[cls + '__f'] = fg
        out[cls + '__fcut'] = m.fcut(**cat)
        out[cls + '__tau'] = m.tau_star(fg, **cat)
        P, A = m.Phi(fg, **cat), m.Ampl(fg, **cat)
        if cls == 'IMRPhenomHM':
            for k in P:
                out[cls + '__phi' + k], out[cls + '__ampl' + k] = P[k], A[k]
            hp, hc = m.hphc(fg, **cat)
            out[cls + '__hp'], out[cls + '__hc'] = hp, hc
        else:
            out[cls + '__phi'], out[cls + '__ampl'] = P, A
        # 1-D grid with scalar parameters (first event)
        one = {k: v[0] for k, v in cat.items()}
        out[cls + '__phi1d'] = np.asarray(m.Phi(fg[:, 0], **one)['22'] if cls == 'IMRPhenomHM' else m.Phi(fg[:, 0], **one))
        for k, v in cat.items():
            evs[cls + '__' + k] = v
    save('wf_values', dict(note='per-model events stored as ev__<cls>__<key>'), evs, out)


def _run_chunk(args):
    """worker of fx_c4big / fx_edge: SNR + Fisher of the unmodified reference on a slice of events"""
    cfg, ev = args
    warnings.filterwarnings('ignore')
    out = run_network(cfg, ev, want_all=False)
    return out['snr'], out['fisher']


def run_network_pool(cfg, ev, chunk=16, procs=None):
    """run_network(..., want_all=False) over slices of `chunk` events in a process pool (the reference under the shim does ~1 IMRPhenomHM
    event per second and core)"""
    import multiprocessing as mp
    n = len(ev['Mc'])
    jobs = [(cfg, {k: v[lo:lo + chunk] for k, v in ev.items()}) for lo in range(0, n, chunk)]
    with mp.get_context('spawn').Pool(procs or min(len(jobs), os.cpu_count() or 1)) as pool:
        parts = pool.map(_run_chunk, jobs)
    return {'snr': np.concatenate([p[0] for p in parts]), 'fisher': np.concatenate([p[1] for p in parts], axis=-1)}


def fx_c4big():
    """IMRPhenomHM on the LVK-O4 network, the first 256 events of the C4 catalog (BASELINE.md 3.3 size)."""
    cfg = dict(model=dict(cls='IMRPhenomHM'), network='LVK-O4', rot=False, fmin=10.)
    ev = take(synthetic.bbh_catalog(100000, synthetic.SEEDS['C4']), 256)
    save('c4_phenomhm_lvk_256', cfg, ev, run_network_pool(cfg, ev))


def fx_edge():
    """SURVEY.md 8(c) edge sets.  Forward-mode convention (what the shim's duals and the engine compute; JAX's reverse mode can
    NaN-poison through the unselected branch of a where):
      * eta = 0.25 exactly: Seta = sqrt(where(eta < 0.25, 1 - 4 eta, 0)) has value 0; its tangent is 0 (oracle/dual.py:_sqrt: what
        jax.jacrev gives -- the select's transpose drops the infinite cotangent of sqrt at 0), so every entry is finite.
      * |chi| in [0.9, 0.99] aligned: gamma2 >= 1, fpeak takes the fabs branch (waveforms.py:1134, 1193).
      * Lambda = 0 and 0 < Lambda < 1: polynomial branch of the spin-induced quadrupole (waveforms.py:1394), kappa2T = 0."""
    fx_edge_eta()
    fx_edge_rest()


def fx_edge_eta():
    bbh = take(synthetic.bbh_catalog(10000, synthetic.SEEDS['C2'] + 50), 16)
    q = _copy(bbh)
    q['eta'][::2] = 0.25
    for tag, cfg in (('phenomd_et2ce', dict(model=dict(cls='IMRPhenomD'), network='ET+2CE', rot=True, fmin=2.)),
                     ('phenomhm_lvk', dict(model=dict(cls='IMRPhenomHM'), network='LVK-O4', rot=False, fmin=10.))):
        save('edge_eta_quarter_' + tag, cfg, q, run_network_pool(cfg, q, chunk=4))
    bns = take(synthetic.bns_catalog(100, synthetic.SEEDS['C1'] + 50, tidal=True), 16)
    qb = {k: v for k, v in _copy(bns).items() if not k.startswith('Lambda')}
    qb['eta'][::2] = 0.25
    cfg = dict(model=dict(cls='TaylorF2_RestrictedPN', kw=dict(use_3p5PN_SpinHO=True)), network='ETSL', rot=True, fmin=2.)
    save('edge_eta_quarter_tf2_etsl', cfg, qb, run_network_pool(cfg, qb, chunk=4))


def fx_edge_rest():
    rng = np.random.default_rng(909)
    hs = take(synthetic.bbh_catalog(10000, synthetic.SEEDS['C2'] + 51), 32)
    hs['chi1z'] = rng.uniform(0.9, 0.99, 32) * np.where(np.arange(32) % 4 == 3, -1., 1.)
    hs['chi2z'] = rng.uniform(0.9, 0.99, 32) * np.where(np.arange(32) % 4 == 2, -1., 1.)
    for tag, cfg in (('phenomd_et2ce', dict(model=dict(cls='IMRPhenomD'), network='ET+2CE', rot=True, fmin=2.)),
                     ('phenomhm_lvk', dict(model=dict(cls='IMRPhenomHM'), network='LVK-O4', rot=False, fmin=10.))):
        save('edge_highspin_' + tag, cfg, hs, run_network_pool(cfg, hs, chunk=4))
    lz = take(synthetic.bns_catalog(100, synthetic.SEEDS['C3'] + 50, tidal=True), 16)
    lz['Lambda1'][0:4] = 0.
    lz['Lambda2'][2:6] = 0.
    lz['Lambda1'][6:9] = rng.uniform(0.05, 0.95, 3)
    lz['Lambda2'][8:11] = rng.uniform(0.05, 0.95, 3)
    cfg = dict(model=dict(cls='IMRPhenomD_NRTidalv2'), network='ET+2CE', rot=True, fmin=2.)
    out = run_network_pool(cfg, lz, chunk=4)
    out.update(masked_last_sample(cfg, lz))
    save('edge_lambda_zero_nrtidal_et2ce', cfg, lz, out)
    cfg = dict(model=dict(cls='TaylorF2_RestrictedPN', kw=dict(is_tidal=True, use_QuadMonTid=True, use_3p5PN_SpinHO=True)), network='ETSL', rot=True, fmin=2.)
    save('edge_lambda_zero_tf2_etsl', cfg, lz, run_network_pool(cfg, lz, chunk=4))


def fx_nsbh():
    """IMRPhenomNSBH (waveforms.py:2752-3374) with the xi_tide table tabulated by the reference's own _tabulate_xiTide
    (oracle/nsbh_table.py): networks + stand-alone Phi / Ampl / tau_star / fcut."""
    wf, sig, net, utils, glob = reference.load()
    ev = synthetic.nsbh_catalog(48, synthetic.SEEDS['NSBH'])
    cfg = dict(model=dict(cls='IMRPhenomNSBH', kw=dict(verbose=False)), network='ET+2CE', rot=True, fmin=2.)
    save('nsbh_et2ce', cfg, ev, run_network_pool(cfg, ev, chunk=6))
    sub = take(ev, 12)
    cfg = dict(model=dict(cls='IMRPhenomNSBH', kw=dict(verbose=False)), network='LVK-O4', rot=False, fmin=10., fmax=512., res=600, fisher_kw=dict(spacing='lin'))
    save('nsbh_lvk_lin_fmax', cfg, sub, run_network(cfg, sub))
    cfg = dict(model=dict(cls='IMRPhenomNSBH', kw=dict(verbose=False, fRef=20., apply_fcut=False, is_chi1chi2=False)), network='ET', rot=True, fmin=5.,
               fisher_kw=dict(use_m1m2=True, use_chi1chi2=False))
    save('nsbh_et_m1m2_chisa_fref_nocut', cfg, sub, run_network(cfg, sub, want_all=False))
    cat = take(ev, 8)
    m = wf.IMRPhenomNSBH(verbose=False)
    fg = np.geomspace(np.full(8, 5.), 0.97 * m.fcut(**cat), 160)
    out = {'f': fg, 'fcut': m.fcut(**cat), 'tau': m.tau_star(fg, **cat), 'phi': m.Phi(fg, **cat), 'ampl': m.Ampl(fg, **cat)}
    fg2 = np.geomspace(np.full(8, 8.), 1.3 * m.fcut(**cat), 120)          # beyond the cut: zeros, and t0 from the grid's own maximum
    out.update({'f_over': fg2, 'phi_over': m.Phi(fg2, **cat), 'ampl_over': m.Ampl(fg2, **cat)})
    save('wf_values_nsbh', dict(model=dict(cls='IMRPhenomNSBH')), cat, out)


def fx_nsbhbig():
    """512 IMRPhenomNSBH events on ET+2CE against the unmodified reference (the size of a parity subset; a second seed)"""
    cfg = dict(model=dict(cls='IMRPhenomNSBH', kw=dict(verbose=False)), network='ET+2CE', rot=True, fmin=2.)
    ev = synthetic.nsbh_catalog(512, synthetic.SEEDS['NSBH'] + 1)
    save('nsbh_et2ce_512', cfg, ev, run_network_pool(cfg, ev, chunk=16))


def fx_nsbhedge():
    """IMRPhenomNSBH edge cases: Lambda_NS = 0 exactly (compactness 1/2, quadrupole branch), 1e-6 (below the 1e-5 switch of the quadrupole
    fit), 1 exactly (switch of the compactness / gamma / delta_2' branches); eta = 0.25; q = 37 .. 93 (towards the upper edge of the
    xi_tide table, q_max = 100); chi_BH -> +1 (last cell of the table, whose chi = 1 nodes are double roots) and -> -1."""
    ev = take(synthetic.nsbh_catalog(64, 777), 20)
    ev = {k: np.array(v) for k, v in ev.items()}
    ev['Lambda2'][0:4] = 0.0
    ev['chi1z'][0:4] = [0.3, -0.2, 0.6, -0.7]
    ev['Lambda2'][4:6] = 1e-6
    ev['Lambda2'][6:8] = 1.0
    ev['eta'][8:10] = 0.25
    m2 = 1.4 * 1.3
    for i, m1 in zip(range(10, 13), (40., 70., 100.)):
        m1 = m1 * 1.3
        ev['Mc'][i] = (m1 * m2) ** 0.6 / (m1 + m2) ** 0.2
        ev['eta'][i] = m1 * m2 / (m1 + m2) ** 2
    ev['chi1z'][13:16] = [0.97, 0.992, 0.9999]
    ev['chi1z'][16:18] = [-0.95, -0.9999]
    cfg = dict(model=dict(cls='IMRPhenomNSBH', kw=dict(verbose=False)), network='ET+2CE', rot=True, fmin=2.)
    save('edge_nsbh_et2ce', cfg, ev, run_network_pool(cfg, ev, chunk=5))


ALL = {'init': fx_init, 'c1': fx_c1, 'c1b': fx_c1b, 'c2': fx_c2, 'var': fx_variants, 'c3': fx_c3, 'c4': fx_c4, 'gw170817': fx_gw170817, 'wf': fx_wfvalues, 'newt': fx_newt, 'ecc': fx_ecc, 'c4big': fx_c4big, 'edge': fx_edge, 'edge_eta': fx_edge_eta, 'nsbh': fx_nsbh, 'nsbhbig': fx_nsbhbig, 'nsbhedge': fx_nsbhedge}

if __name__ == '__main__':
    warnings.filterwarnings('ignore')
    if not reference.available():
        sys.exit('the reference tree is not mounted; fixtures can only be generated in the build container')
    for k in (sys.argv[1:] or list(ALL)):
        ALL[k]()
