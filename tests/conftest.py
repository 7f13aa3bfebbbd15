import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLD = os.path.join(ROOT, 'tests', 'golden')

# tolerances of BASELINE.json's north star
SNR_RTOL = 1e-9          # |SNR/SNR_ref - 1|
FISHER_TOL = 1e-6        # |dF_ij| / sqrt(F_ii F_jj)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def fisher_err(F, Fref):
    dg = np.sqrt(np.einsum('iin->in', Fref))
    return float(np.max(np.abs(F - Fref) / (dg[:, None, :] * dg[None, :, :])))


def snr_err(s, sref):
    return float(np.max(np.abs(np.asarray(s) / np.asarray(sref) - 1)))


def load_golden(name):
    z = np.load(os.path.join(GOLD, name + '.npz'))
    cfg = json.loads(str(z['config']))
    ev = {k[4:]: z[k] for k in z.files if k.startswith('ev__')}
    out = {k: z[k] for k in z.files if not k.startswith('ev__') and k != 'config'}
    return cfg, ev, out


def make_network(kind, cfg):
    """kind: 'engine' | 'port'.  Build the DetNet/Network described by a golden fixture's config."""
    from gwfast_b200 import synthetic
    kw = {}
    if cfg.get('fmax') is not None:
        kw['fmax'] = cfg['fmax']
    if kind == 'engine':
        from gwfast_b200 import waveforms, signal, network
        model = getattr(waveforms, cfg['model']['cls'])(**cfg['model'].get('kw', {}))
        return network.DetNet(synthetic.build_network(signal.GWSignal, model, cfg['network'], useEarthMotion=cfg['rot'], fmin=cfg['fmin'], **kw), verbose=False)
    from oracle.port import waveforms as PW, detector as PD
    model = getattr(PW, cfg['model']['cls'])(**cfg['model'].get('kw', {}))
    return PD.Network(synthetic.build_network(PD.Detector, model, cfg['network'], useEarthMotion=cfg['rot'], fmin=cfg['fmin'], **kw))


def copy_events(ev):
    return {k: np.array(v, dtype=float) for k, v in ev.items()}


@pytest.fixture(scope='session')
def has_reference():
    from oracle import reference
    return reference.available()
