"""Golden vectors for the derivative outputs: the UNMODIFIED reference's DetNet.FisherMatr with
return_derivatives=True / return_SNR_derivatives=True (signal.py:917-945, network.py:124-152), run under the oracle shim.
TEST INFRASTRUCTURE; container only.  Writes tests/golden/deriv_*.npz (inputs + configuration + outputs)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import reference  # noqa: E402
from oracle.make_golden import _model, _copy, save, take, REF_PSDS  # noqa: E402
from gwfast_b200 import synthetic  # noqa: E402


def run(cfg, ev, derivs=True):
    wf, sig, net, utils, glob = reference.load()
    sigs = synthetic.build_network(sig.GWSignal, _model(wf, cfg['model']), cfg['network'], useEarthMotion=cfg['rot'], fmin=cfg['fmin'], psd_root=REF_PSDS)
    N = net.DetNet(sigs, verbose=False)
    res = cfg.get('res', 1000)
    fkw = dict(cfg.get('fisher_kw', {}))
    out = {}
    F, SD = N.FisherMatr(_copy(ev), res=res, return_SNR_derivatives=True, **fkw)
    for k in F:
        out['fisher__' + k] = np.asarray(F[k], dtype=float)
        out['snrderiv__' + k] = np.asarray(SD[k], dtype=float)
    if derivs:
        F2, D = N.FisherMatr(_copy(ev), res=res, return_derivatives=True, **fkw)
        for k in D:
            out['deriv__' + k] = np.asarray(D[k], dtype=complex)
    return out


def hm_strain_derivs():
    """IMRPhenomHM return_derivatives (signal.py:917-945 with the hphc branch of the analytic derivatives, :1378-1380): LVK without and the
    ET triangle with Earth rotation"""
    cfg = dict(model=dict(cls='IMRPhenomHM'), network='LVK-O4', rot=False, fmin=10., res=300)
    ev = take(synthetic.bbh_catalog(10000, synthetic.SEEDS['C4']), 4)
    save('deriv_hm_strain_lvk', cfg, ev, run(cfg, ev))
    cfg = dict(model=dict(cls='IMRPhenomHM'), network='ET', rot=True, fmin=2., res=200)
    save('deriv_hm_strain_et', cfg, take(ev, 3), run(cfg, take(ev, 3)))


def nsbh_derivs():
    """IMRPhenomNSBH: 13 rows, the ET triangle with Earth rotation"""
    cfg = dict(model=dict(cls='IMRPhenomNSBH', kw=dict(verbose=False)), network='ET', rot=True, fmin=2., res=200)
    ev = take(synthetic.nsbh_catalog(48, synthetic.SEEDS['NSBH']), 4)
    save('deriv_nsbh', cfg, ev, run(cfg, ev))


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'hm':
        hm_strain_derivs()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'nsbh':
        nsbh_derivs()
        sys.exit(0)
    cfg = dict(model=dict(cls='IMRPhenomD'), network='ET+2CE', rot=True, fmin=2., res=200)
    ev = take(synthetic.bbh_catalog(10000, synthetic.SEEDS['C2']), 6)
    save('deriv_c2', cfg, ev, run(cfg, ev))
    cfg = dict(model=dict(cls='TaylorF2_RestrictedPN', kw=dict(is_tidal=True, use_3p5PN_SpinHO=True)), network='ETSL', rot=True, fmin=2., res=160)
    ev = take(synthetic.bns_catalog(100, synthetic.SEEDS['C1'] + 100, tidal=True), 4)
    save('deriv_tf2tidal', cfg, ev, run(cfg, ev))
    cfg = dict(model=dict(cls='IMRPhenomD', kw=dict(is_chi1chi2=False)), network='LVK-O4', rot=False, fmin=10., res=128,
               fisher_kw=dict(use_m1m2=True, use_chi1chi2=False))
    ev = take(synthetic.bbh_catalog(10000, synthetic.SEEDS['C2']), 4)
    save('deriv_m1m2_lvk', cfg, ev, run(cfg, ev))
    cfg = dict(model=dict(cls='IMRPhenomHM'), network='LVK-O4', rot=False, fmin=10., res=300)
    ev = take(synthetic.bbh_catalog(10000, synthetic.SEEDS['C4']), 4)
    save('deriv_hm_lvk', cfg, ev, run(cfg, ev, derivs=False))
    hm_strain_derivs()
    cfg = dict(model=dict(cls='IMRPhenomD_NRTidalv2'), network='ET', rot=True, fmin=2., res=200)
    ev = take(synthetic.bns_catalog(10000, synthetic.SEEDS['C3'], tidal=True), 4)
    save('deriv_nrtidal', cfg, ev, run(cfg, ev))
    nsbh_derivs()
