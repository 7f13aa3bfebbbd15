"""Golden vectors for WFOverlap: the UNMODIFIED reference's GWSignal.WFOverlap / DetNet.WFOverlap
(signal.py:1759-1930, network.py:198-225) run under the oracle shim.  TEST INFRASTRUCTURE; container only.
Writes tests/golden/wfo_*.npz: the two event dicts (ev1__*, ev2__*), the configuration and the reference outputs."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import reference  # noqa: E402
from oracle.make_golden import _model, _copy, take, REF_PSDS, GOLD  # noqa: E402
from gwfast_b200 import synthetic  # noqa: E402


def perturb(ev, rng, rel=1e-3):
    """a second parameter point close to the first (overlaps of interest are near 1), same sky position and time"""
    out = _copy(ev)
    for k in ('Mc', 'eta', 'chi1z', 'chi2z', 'dL', 'iota', 'psi', 'Phicoal'):
        if k in out:
            out[k] = out[k] * (1. + rel * rng.uniform(-1, 1, out[k].shape))
    if 'eta' in out:
        out['eta'] = np.minimum(out['eta'], 0.2499)
    for k in ('Lambda1', 'Lambda2'):
        if k in out:
            out[k] = out[k] * (1. + 0.05 * rng.uniform(-1, 1, out[k].shape))
    return out


def run(cfg, ev1, ev2):
    wf, sig, net, utils, glob = reference.load()
    kw = {}
    if cfg.get('fmax') is not None:
        kw['fmax'] = cfg['fmax']
    WF1, WF2 = _model(wf, cfg['model1']), _model(wf, cfg['model2'])
    sigs = synthetic.build_network(sig.GWSignal, WF1, cfg['network'], useEarthMotion=cfg['rot'], fmin=cfg['fmin'], psd_root=REF_PSDS, **kw)
    N = net.DetNet(sigs, verbose=False)
    res = cfg.get('res', 1000)
    out = {'overlap__net': np.asarray(N.WFOverlap(WF1, WF2, _copy(ev1), _copy(ev2), res=res), dtype=float)}
    for d, s in sigs.items():
        o, s1, s2 = s.WFOverlap(WF1, WF2, _copy(ev1), _copy(ev2), res=res, return_separate=True)
        out['inner__' + d] = np.asarray(o, dtype=float)
        out['snr1__' + d] = np.asarray(s1, dtype=float)
        out['snr2__' + d] = np.asarray(s2, dtype=float)
        out['overlap__' + d] = np.asarray(s.WFOverlap(WF1, WF2, _copy(ev1), _copy(ev2), res=res), dtype=float)
    return out


def save(name, cfg, ev1, ev2, out):
    data = {'config': np.array(json.dumps(cfg))}
    data.update({'ev1__' + k: np.asarray(v, dtype=float) for k, v in ev1.items()})
    data.update({'ev2__' + k: np.asarray(v, dtype=float) for k, v in ev2.items()})
    data.update(out)
    path = os.path.join(GOLD, name + '.npz')
    np.savez_compressed(path, **data)
    print('wrote %s (%.1f KB)' % (path, os.path.getsize(path) / 1024.), {k: v for k, v in out.items() if k.startswith('overlap')})


def hm_cases():
    """IMRPhenomHM against itself at nearby parameters (LVK), and against IMRPhenomD (ET triangle with rotation): GWstrain's hp Fp + hc Fc"""
    rng = np.random.default_rng(20260097)
    cfg = dict(model1=dict(cls='IMRPhenomHM'), model2=dict(cls='IMRPhenomHM'), network='LVK-O4', rot=False, fmin=10., res=500)
    ev1 = take(synthetic.bbh_catalog(10000, synthetic.SEEDS['C4']), 5)
    ev2 = perturb(ev1, rng)
    save('wfo_hm_lvk', cfg, ev1, ev2, run(cfg, ev1, ev2))
    cfg = dict(model1=dict(cls='IMRPhenomHM'), model2=dict(cls='IMRPhenomD'), network='ET', rot=True, fmin=2., res=400)
    ev1 = take(ev1, 4)
    save('wfo_hm_phenomd_et', cfg, ev1, _copy(ev1), run(cfg, ev1, _copy(ev1)))


def nsbh_cases():
    """IMRPhenomNSBH against itself at nearby parameters (ET + 2 CE with rotation) and against IMRPhenomD_NRTidalv2 at the same ones (LVK)"""
    rng = np.random.default_rng(20260096)
    nsbh = dict(cls='IMRPhenomNSBH', kw=dict(verbose=False))
    cfg = dict(model1=nsbh, model2=nsbh, network='ET+2CE', rot=True, fmin=2., res=500)
    ev1 = take(synthetic.nsbh_catalog(48, synthetic.SEEDS['NSBH']), 5)
    ev2 = perturb(ev1, rng, rel=1e-4)
    save('wfo_nsbh_et2ce', cfg, ev1, ev2, run(cfg, ev1, ev2))
    cfg = dict(model1=nsbh, model2=dict(cls='IMRPhenomD_NRTidalv2'), network='LVK-O4', rot=False, fmin=10., res=400)
    save('wfo_nsbh_nrtidal_lvk', cfg, take(ev1, 4), _copy(take(ev1, 4)), run(cfg, take(ev1, 4), _copy(take(ev1, 4))))


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'hm':
        hm_cases()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'nsbh':
        nsbh_cases()
        sys.exit(0)
    rng = np.random.default_rng(20260099)
    # same model, nearby parameters, ET triangle + 2 CE with Earth rotation
    cfg = dict(model1=dict(cls='IMRPhenomD'), model2=dict(cls='IMRPhenomD'), network='ET+2CE', rot=True, fmin=2., res=600)
    ev1 = take(synthetic.bbh_catalog(10000, synthetic.SEEDS['C2']), 6)
    ev2 = perturb(ev1, rng)
    save("wfo_phenomd_et2ce", cfg, ev1, ev2, run(cfg, ev1, ev2))
    # two different models (different cuts: the grid ends at the larger one), same parameters, LVK without rotation
    cfg = dict(model1=dict(cls='IMRPhenomD'), model2=dict(cls='TaylorF2_RestrictedPN'), network='LVK-O4', rot=False, fmin=10., res=500)
    ev1 = take(synthetic.bbh_catalog(10000, synthetic.SEEDS['C4']), 5)
    save('wfo_phenomd_tf2_lvk', cfg, ev1, _copy(ev1), run(cfg, ev1, _copy(ev1)))
    # tidal models, single L detector, fmax clip (the reference's where(fcutUse > fmax, fmax, fcut1))
    cfg = dict(model1=dict(cls='IMRPhenomD_NRTidalv2'), model2=dict(cls='TaylorF2_RestrictedPN', kw=dict(is_tidal=True)), network='ETSL', rot=True,
               fmin=2., fmax=700., res=400)
    ev1 = take(synthetic.bns_catalog(10000, synthetic.SEEDS['C3'], tidal=True), 4)
    rng = np.random.default_rng(20260098)
    ev2 = perturb(ev1, rng, rel=1e-5)
    save('wfo_tidal_etsl_fmax', cfg, ev1, ev2, run(cfg, ev1, ev2))
