"""Golden vectors for the catalog driver: the UNMODIFIED reference's compute_errs
(run/calculate_forecasts_from_catalog.py:410-635) on a 200-event catalog, run under the oracle shim (mpmath as installed).
TEST INFRASTRUCTURE; container only.  Writes tests/golden/forecast_*.npz.

    python -m oracle.make_golden_forecast
"""
import argparse
import contextlib
import importlib.util
import io
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from oracle import reference  # noqa: E402
from oracle.make_golden import save, take, _copy, REF_PSDS  # noqa: E402
from gwfast_b200 import synthetic  # noqa: E402


def load_run_script():
    """the reference's run script as a module (its argparse / pool main is under __main__ and is not executed)"""
    wf = reference.load()[0]
    path = os.path.join(reference.REFERENCE_ROOT, 'run', 'calculate_forecasts_from_catalog.py')
    spec = importlib.util.spec_from_file_location('ref_run_script', path)
    mod = importlib.util.module_from_spec(spec)
    # the script builds a dictionary of one instance of every waveform class at import time (:60-70); IMRPhenomNSBH() needs the
    # xiTide table (absent here, and its tabulation path lacks an import): that one constructor is stubbed for the import only --
    # compute_errs never touches it
    real = wf.IMRPhenomNSBH
    wf.IMRPhenomNSBH = lambda *a, **k: None
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            spec.loader.exec_module(mod)
    finally:
        wf.IMRPhenomNSBH = real
    return mod


def run(tag, cfg, ev, snr_th, duty_factor=None, seeds=None, params_fix=()):
    wf, sig, net, utils, glob = reference.load()
    run_script = load_run_script()
    sigs = synthetic.build_network(sig.GWSignal, getattr(wf, cfg['model']['cls'])(**cfg['model'].get('kw', {})), cfg['network'], useEarthMotion=cfg['rot'],
                                   fmin=cfg['fmin'], psd_root=REF_PSDS)
    N = net.DetNet(sigs, verbose=False)
    FLAGS = argparse.Namespace(snr_th=snr_th, duty_factor=duty_factor, seeds=seeds, params_fix=list(params_fix), compute_fisher=1, return_all=1,
                               return_derivatives=0, return_snr_derivatives=0)
    n = len(ev['Mc'])
    with contextlib.redirect_stdout(io.StringIO()):
        snrs_all, Fres, eps, cov, sky, cond, idxs = run_script.compute_errs(_copy(ev), N, FLAGS, 0, n)
    out = {'eps': np.asarray(eps, dtype=float), 'cov': np.asarray(cov, dtype=float), 'sky_area_90': np.asarray(sky, dtype=float),
           'cond_numbers': np.asarray(cond, dtype=float), 'idxs_detected': np.ravel(np.asarray(idxs)), 'errors': np.sqrt(np.einsum('iin->in', np.asarray(cov, dtype=float)))}
    for k, v in snrs_all.items():
        out['snr__' + k] = np.asarray(v, dtype=float)
    for k, v in Fres.items():
        out['fisher__' + k] = np.asarray(v, dtype=float)
    cfg = dict(cfg, snr_th=snr_th, duty_factor=duty_factor, seeds=seeds, params_fix=list(params_fix))
    save('forecast_' + tag, cfg, ev, out)
    print(tag, 'detected', len(out['idxs_detected']), 'of', n, 'median sky area', np.median(out['sky_area_90']))


if __name__ == '__main__':
    warnings.filterwarnings('ignore')
    if not reference.available():
        sys.exit('the reference tree is not mounted; fixtures can only be generated in the build container')
    cfg = dict(model=dict(cls='IMRPhenomD'), network='ET+2CE', rot=True, fmin=2.)
    ev = take(synthetic.bbh_catalog(10000, synthetic.SEEDS['C2']), 200)
    run('c2_200', cfg, ev, snr_th=40.)
    # duty factor (Bernoulli masks per arm, seeded per detector) and fixed parameters, LVK network without rotation
    cfg = dict(model=dict(cls='IMRPhenomD'), network='LVK-O4', rot=False, fmin=10.)
    ev = take(synthetic.bbh_catalog(10000, synthetic.SEEDS['C4']), 96)
    ev['dL'] = ev['dL'] * 0.05
    run('lvk_duty_fix', cfg, ev, snr_th=12., duty_factor=0.7, seeds=[11, 12, 13, 14], params_fix=['iota', 'psi'])
