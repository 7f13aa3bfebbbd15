"""Strain-level methods of GWSignal (gwfast/signal.py:342-655) served by gwf_signal_grid / gwf_pattern, against the unmodified reference
(tests/golden/signal_methods.npz, oracle/make_golden_signal.py): GWAmplitudes, GWPhase, GWstrain with its re-parametrisation switches and
return_single_comp, _PatternFunction, _DeltLoc, optimal_location."""
import json
import os

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _rel(a, b):
    """max deviation relative to the largest magnitude of each column (amplitudes span decades along a grid)"""
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b) / np.max(np.abs(b), axis=0)))


def _detector(cls, kw, site, rot_on, nomo):
    from gwfast_b200 import waveforms, signal, gwfastGlobals as glob
    s = glob.detectors[site]
    return signal.GWSignal(getattr(waveforms, cls)(**kw), psd_path=os.path.join(glob.detPath, 'ET-0000A-18.txt'), detector_shape=s['shape'], det_lat=s['lat'],
                           det_long=s['long'], det_xax=s['xax'], verbose=False, useEarthMotion=rot_on, noMotion=nomo, fmin=2.)


@pytest.mark.parametrize('case', [0, 1, 2, 3, 4])
def test_amplitudes_phase_and_strain_match_reference(case):
    from gwfast_b200 import gwfastUtils as utils
    cfg, evs, out = load_golden('signal_methods')
    cls, kw, tidal = cfg['cases'][case]
    ev = {k[len(cls) + 2:]: v for k, v in evs.items() if k.startswith(cls + '__')}
    fg = out[cls + '__f']
    n = fg.shape[1]
    z = np.zeros(n)
    for site, rot_on, nomo in cfg['dets']:
        d = _detector(cls, kw, site, rot_on, nomo)
        key = '%s__%s' % (cls, site)
        for rot in (0., 60.):
            Ap, Ac = d.GWAmplitudes(dict(ev), fg, rot=rot)
            assert Ap.shape == fg.shape and _rel(Ap, out['%s__Ap%d' % (key, rot)]) < 1e-9 and _rel(Ac, out['%s__Ac%d' % (key, rot)]) < 1e-9, (key, rot)
        if cls != 'IMRPhenomHM':
            psi = d.GWPhase(dict(ev), fg)
            assert np.max(np.abs(psi - out[key + '__psi']) / (1 + np.abs(out[key + '__psi']))) < 1e-12, key
        else:
            with pytest.raises(TypeError):
                d.GWPhase(dict(ev), fg)
        L1, L2 = ev.get('Lambda1', z), ev.get('Lambda2', z)
        args = (ev['dL'], ev['theta'], ev['phi'], ev['iota'], ev['psi'], ev['tcoal'], ev['Phicoal'])
        h = d.GWstrain(fg, ev['Mc'], ev['eta'], *args, ev['chi1z'], ev['chi2z'], z, z, z, z, L1, L2, z, rot=60., is_chi1chi2=True, is_Lam1Lam2=True)
        assert h.dtype == np.complex128 and h.shape == fg.shape
        # the phase is O(1e5) rad at the low end of a BNS grid: 1e-7 relative to the column's largest |h| is 1e-12 of the phase
        assert _rel(h, out[key + '__strain']) < 1e-7, key
        m1, m2 = utils.m1m2_from_Mceta(ev['Mc'], ev['eta'])
        chiS, chiA = 0.5 * (ev['chi1z'] + ev['chi2z']), 0.5 * (ev['chi1z'] - ev['chi2z'])
        Lt, dLam = utils.Lamt_delLam_from_Lam12(L1, L2, ev['eta']) if tidal else (z, z)
        h2 = d.GWstrain(fg, m1, m2, *args, chiS, chiA, z, z, z, z, Lt, dLam, z, rot=0., is_m1m2=True)
        assert _rel(h2, out[key + '__strain_m1m2']) < 1e-6, key      # the (m1, m2) and (LambdaTilde, deltaLambda) round trips cost a few ulp of Mc, eta, Lambda
        for comp in ('Ap', 'Ac', 'At'):
            got = d.GWstrain(fg, ev['Mc'], ev['eta'], *args, ev['chi1z'], ev['chi2z'], z, z, z, z, L1, L2, z, rot=0., is_chi1chi2=True, is_Lam1Lam2=True,
                             return_single_comp=comp)
            assert _rel(got, out['%s__single_%s' % (key, comp)]) < 1e-9, (key, comp)
        if cls != 'IMRPhenomHM':
            for comp in ('Psip', 'Psit'):
                got = d.GWstrain(fg, ev['Mc'], ev['eta'], *args, ev['chi1z'], ev['chi2z'], z, z, z, z, L1, L2, z, rot=0., is_chi1chi2=True,
                                 is_Lam1Lam2=True, return_single_comp=comp)
                want = out['%s__single_%s' % (key, comp)]
                assert np.max(np.abs(got - want) / (1 + np.abs(want))) < 1e-12, (key, comp)
        with pytest.raises(ValueError):
            d.GWstrain(fg, ev['Mc'], ev['eta'], *args, ev['chi1z'], ev['chi2z'], z, z, z, z, L1, L2, z, return_single_comp='nope')
        one = {k: v[0] for k, v in ev.items()}
        Ap1, Ac1 = d.GWAmplitudes(one, fg[:, 0])
        assert Ap1.shape == (fg.shape[0],) and _rel(Ap1, out[key + '__Ap1d']) < 1e-9 and _rel(Ac1, out[key + '__Ac1d']) < 1e-9


def test_pattern_functions_and_delays_match_reference():
    from gwfast_b200 import waveforms, signal, gwfastGlobals as glob
    cfg, evs, out = load_golden('signal_methods')
    th, ph, t, ps = out['pat_theta'], out['pat_phi'], out['pat_t'], out['pat_psi']
    for name, shape, lat, lon, xax in (('default', 'L', 40.44, 9.45, 0.), ('ETS', 'T', glob.detectors['ETS']['lat'], glob.detectors['ETS']['long'],
                                                                          glob.detectors['ETS']['xax'])):
        d = signal.GWSignal(waveforms.TaylorF2_RestrictedPN(), psd_path=os.path.join(glob.detPath, 'ET-0000A-18.txt'), detector_shape=shape, det_lat=lat,
                            det_long=lon, det_xax=xax, verbose=False, fmin=2.)
        for rot in (0., 60., 120.):
            Fp, Fc = d._PatternFunction(th, ph, t, ps, rot=rot)
            assert np.max(np.abs(Fp - out['pat_%s_Fp%d' % (name, rot)])) < 1e-14 and np.max(np.abs(Fc - out['pat_%s_Fc%d' % (name, rot)])) < 1e-14
        assert np.max(np.abs(d._DeltLoc(th, ph, t) - out['pat_%s_dt' % name])) < 1e-16
        Fp2, _ = d._PatternFunction(th, ph, out['pat_%s_t2' % name], ps)
        assert Fp2.shape == (5, 64) and np.max(np.abs(Fp2 - out['pat_%s_Fp2' % name])) < 1e-14
        # scalars in, scalars out
        f0, c0 = d._PatternFunction(float(th[0]), float(ph[0]), float(t[0]), float(ps[0]))
        assert np.ndim(f0) == 0 and abs(f0 - out['pat_%s_Fp0' % name][0]) < 1e-14
        if name == 'default':
            # notebooks/gwfast_tutorial.ipynb: optimal_location(0.) = [0.8636426, 0.1643505]; at the optimum of an L the response is 1
            loc = d.optimal_location(0.)
            assert np.max(np.abs(loc - np.array([0.8636426, 0.1643505]))) < 5e-3
            assert np.max(np.abs(loc - out['optimal_location_0'])) < 5e-3
            Fp, Fc = d._PatternFunction(loc[0], loc[1], 0., 0.)
            assert abs(np.sqrt(Fp ** 2 + Fc ** 2) - 1.0) < 1e-5
