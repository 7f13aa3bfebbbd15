"""Golden vectors for the strain-level methods of GWSignal (gwfast/signal.py:342-655): GWAmplitudes, GWPhase, GWstrain (with its
re-parametrisation switches and return_single_comp), _PatternFunction, _DeltLoc and optimal_location, from the UNMODIFIED reference
under the numpy shim (TEST INFRASTRUCTURE; container only).

    python -m oracle.make_golden_signal
"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import reference  # noqa: E402
from oracle.make_golden import save, REF_PSDS  # noqa: E402
from gwfast_b200 import synthetic  # noqa: E402

CASES = [('TaylorF2_RestrictedPN', dict(use_3p5PN_SpinHO=True), 'bns', False), ('IMRPhenomD', {}, 'bbh', False),
         ('IMRPhenomD_NRTidalv2', {}, 'bns', True), ('IMRPhenomHM', {}, 'bbh', False), ('IMRPhenomNSBH', dict(verbose=False), 'nsbh', True)]
DETS = [('ETS', True, False), ('CE1Id', False, False), ('ETSL', False, True)]      # (site, useEarthMotion, noMotion)


def main():
    warnings.filterwarnings('ignore')
    wf, sig, net, utils, glob = reference.load()
    out, evs = {}, {}
    n = 6
    for cls, kw, kind, tidal in CASES:
        ev = (synthetic.bbh_catalog(n, 4100) if kind == 'bbh' else synthetic.nsbh_catalog(n, 4102) if kind == 'nsbh' else
              synthetic.bns_catalog(n, 4101, tidal=tidal))
        for k, v in ev.items():
            evs['%s__%s' % (cls, k)] = v
        m = getattr(wf, cls)(**kw)
        fg = np.geomspace(np.full(n, 5.), 0.95 * m.fcut(**ev), 150)
        out[cls + '__f'] = fg
        z = np.zeros(n)
        for site, rot_on, nomo in DETS:
            s = glob.detectors[site]
            d = sig.GWSignal(getattr(wf, cls)(**kw), psd_path=os.path.join(REF_PSDS, 'ET-0000A-18.txt'), detector_shape=s['shape'], det_lat=s['lat'],
                             det_long=s['long'], det_xax=s['xax'], verbose=False, useEarthMotion=rot_on, noMotion=nomo, fmin=2.)
            key = '%s__%s' % (cls, site)
            for rot in (0., 60.):
                Ap, Ac = d.GWAmplitudes(dict(ev), fg, rot=rot)
                out['%s__Ap%d' % (key, rot)], out['%s__Ac%d' % (key, rot)] = np.asarray(Ap), np.asarray(Ac)
            if cls != 'IMRPhenomHM':
                out[key + '__psi'] = np.asarray(d.GWPhase(dict(ev), fg))
            L1 = ev.get('Lambda1', z)
            L2 = ev.get('Lambda2', z)
            args = (ev['dL'], ev['theta'], ev['phi'], ev['iota'], ev['psi'], ev['tcoal'], ev['Phicoal'])
            out[key + '__strain'] = np.asarray(d.GWstrain(fg, ev['Mc'], ev['eta'], *args, ev['chi1z'], ev['chi2z'], z, z, z, z, L1, L2, z, rot=60.,
                                                         is_chi1chi2=True, is_Lam1Lam2=True))
            # the Fisher parametrisation: (m1, m2), (chiS, chiA), (LambdaTilde, deltaLambda)
            m1, m2 = utils.m1m2_from_Mceta(ev['Mc'], ev['eta'])
            chiS, chiA = 0.5 * (ev['chi1z'] + ev['chi2z']), 0.5 * (ev['chi1z'] - ev['chi2z'])
            Lt, dLam = utils.Lamt_delLam_from_Lam12(L1, L2, ev['eta']) if tidal else (z, z)
            out[key + '__strain_m1m2'] = np.asarray(d.GWstrain(fg, m1, m2, *args, chiS, chiA, z, z, z, z, Lt, dLam, z, rot=0., is_m1m2=True))
            for comp in ('Ap', 'Ac', 'Psip', 'At', 'Psit'):
                out['%s__single_%s' % (key, comp)] = np.asarray(d.GWstrain(fg, ev['Mc'], ev['eta'], *args, ev['chi1z'], ev['chi2z'], z, z, z, z, L1, L2, z,
                                                                            rot=0., is_chi1chi2=True, is_Lam1Lam2=True, return_single_comp=comp))
            # 1-D grid with scalar parameters (first event)
            one = {k: v[0] for k, v in ev.items()}
            Ap1, Ac1 = d.GWAmplitudes(one, fg[:, 0])
            out[key + '__Ap1d'], out[key + '__Ac1d'] = np.asarray(Ap1), np.asarray(Ac1)
    # pattern functions and delays on random points, default detector of the notebooks and the ET triangle
    rng = np.random.default_rng(4102)
    th, ph, t, ps = np.arccos(rng.uniform(-1, 1, 64)), rng.uniform(0, 2 * np.pi, 64), rng.uniform(0, 1, 64), rng.uniform(0, np.pi, 64)
    out.update(pat_theta=th, pat_phi=ph, pat_t=t, pat_psi=ps)
    for name, shape, lat, lon, xax in (('default', 'L', 40.44, 9.45, 0.), ('ETS', 'T', glob.detectors['ETS']['lat'], glob.detectors['ETS']['long'],
                                                                          glob.detectors['ETS']['xax'])):
        d = sig.GWSignal(wf.TaylorF2_RestrictedPN(), psd_path=os.path.join(REF_PSDS, 'ET-0000A-18.txt'), detector_shape=shape, det_lat=lat, det_long=lon,
                         det_xax=xax, verbose=False, fmin=2.)
        for rot in (0., 60., 120.):
            Fp, Fc = d._PatternFunction(th, ph, t, ps, rot=rot)
            out['pat_%s_Fp%d' % (name, rot)], out['pat_%s_Fc%d' % (name, rot)] = np.asarray(Fp), np.asarray(Fc)
        out['pat_%s_dt' % name] = np.asarray(d._DeltLoc(th, ph, t))
        # broadcasting: (res, N) times against (N,) angles, as GWAmplitudes calls it
        t2 = rng.uniform(0, 1, (5, 64))
        Fp2, Fc2 = d._PatternFunction(th, ph, t2, ps)
        out['pat_%s_t2' % name], out['pat_%s_Fp2' % name] = t2, np.asarray(Fp2)
        if name == 'default':
            out['optimal_location_0'] = np.asarray(d.optimal_location(0.))       # notebook: [0.8636426, 0.1643505]
            print('optimal_location(0.) =', out['optimal_location_0'])
    save('signal_methods', dict(note='events stored as ev__<cls>__<key>', dets=[list(x) for x in DETS], cases=[[c[0], c[1], c[3]] for c in CASES]), evs, out)


if __name__ == '__main__':
    if not reference.available():
        sys.exit('the reference tree is not mounted; fixtures can only be generated in the build container')
    main()
