"""SURVEY.md 8(d) parity subsets, driver-run: the first 1000 events of the C1 / C2 / C3 synthetic catalogs through the engine (B200) and
through the oracle port (host cores, a process pool), the 256-event IMRPhenomHM fixture generated from the unmodified reference
(oracle/make_golden.py:fx_c4big), and the three edge sets of SURVEY.md 8(c) (eta = 0.25, |chi| >= 0.9, Lambda = 0).
Tolerances are the north star's: SNR 1e-9 relative, Fisher 1e-6 relative to sqrt(F_ii F_jj)."""
import multiprocessing as mp
import os

import numpy as np
import pytest

from conftest import load_golden, make_network, copy_events, fisher_err, snr_err, SNR_RTOL, FISHER_TOL

pytestmark = pytest.mark.gpu

N_SUBSET = 1000
CASES = {
    'C1': dict(model=dict(cls='TaylorF2_RestrictedPN'), network='ETSL', rot=True, fmin=2., cat=('bns', 'C1', False)),
    'C2': dict(model=dict(cls='IMRPhenomD'), network='ET+2CE', rot=True, fmin=2., cat=('bbh', 'C2', False)),
    # NRTidalv2: the reference's last grid sample sits exactly on the end of the Planck taper (0 or 1 by last-bit rounding, SURVEY.md
    # App. A-3); the engine defines the taper as 0 there, and so does the port with taper_end_zero=True
    'C3': dict(model=dict(cls='IMRPhenomD_NRTidalv2', kw=dict(taper_end_zero=True)), network='ET+2CE', rot=True, fmin=2., cat=('bns', 'C3', True)),
}


def _catalog(spec, n):
    from gwfast_b200 import synthetic
    kind, seed, tidal = spec
    ev = synthetic.bbh_catalog(10000, synthetic.SEEDS[seed]) if kind == 'bbh' else synthetic.bns_catalog(10000, synthetic.SEEDS[seed], tidal=tidal)
    return {k: v[:n] for k, v in ev.items()}


def _port_chunk(args):
    import warnings
    warnings.filterwarnings('ignore')
    name, lo, hi = args
    cfg = CASES[name]
    ev = {k: v[lo:hi] for k, v in _catalog(cfg['cat'], N_SUBSET).items()}
    port = make_network('port', cfg)
    return port.SNR(copy_events(ev)), port.FisherMatr(copy_events(ev))


@pytest.fixture(scope='module')
def pool():
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    p = mp.get_context('spawn').Pool(min(32, os.cpu_count() or 1))
    yield p
    p.close()
    p.join()


@pytest.mark.timeout(600)
@pytest.mark.parametrize('name', ['C1', 'C2', 'C3'])
def test_first_1000_events_match_the_oracle_port(name, pool):
    cfg = CASES[name]
    ev = _catalog(cfg['cat'], N_SUBSET)
    eng_cfg = dict(cfg, model=dict(cls=cfg['model']['cls']))
    net = make_network('engine', eng_cfg)
    snr, F = net.SNR(copy_events(ev)), net.FisherMatr(copy_events(ev))
    Ff, sf = net.FisherMatr(copy_events(ev), return_SNR=True)
    step = 20
    parts = pool.map(_port_chunk, [(name, lo, min(lo + step, N_SUBSET)) for lo in range(0, N_SUBSET, step)])
    snr_o = np.concatenate([p[0] for p in parts])
    F_o = np.concatenate([p[1] for p in parts], axis=-1)
    assert snr.shape == (N_SUBSET,) and F.shape == F_o.shape
    es, ef = snr_err(snr, snr_o), fisher_err(F, F_o)
    print('%s: first %d events vs oracle port: SNR %.2e, Fisher %.2e' % (name, N_SUBSET, es, ef))
    assert es < SNR_RTOL and ef < FISHER_TOL
    # the fused launch (FisherMatr(return_SNR=True)) is held to the same bar
    assert snr_err(sf, snr_o) < SNR_RTOL and fisher_err(Ff, F_o) < FISHER_TOL


def test_phenomhm_256_events_match_the_reference():
    """BASELINE.json configs[3] at BASELINE.md 3.3's size: the first 256 events of the C4 catalog, IMRPhenomHM on H1/L1/Virgo/KAGRA,
    against the unmodified reference (IMRPhenomHM has no port: parity is pinned on the reference's own outputs)."""
    cfg, ev, out = load_golden('c4_phenomhm_lvk_256')
    net = make_network('engine', cfg)
    snr, F = net.SNR(copy_events(ev)), net.FisherMatr(copy_events(ev))
    es, ef = snr_err(snr, out['snr']), fisher_err(F, out['fisher'])
    print('C4: 256 events vs reference: SNR %.2e, Fisher %.2e' % (es, ef))
    assert es < SNR_RTOL and ef < FISHER_TOL
    Ff, sf = net.FisherMatr(copy_events(ev), return_SNR=True)
    assert snr_err(sf, out['snr']) < SNR_RTOL and np.array_equal(Ff, F)


# ---------------------------------------------------------------------------------------------- edge sets, SURVEY.md 8(c)
@pytest.mark.parametrize('name', ['edge_eta_quarter_phenomd_et2ce', 'edge_eta_quarter_tf2_etsl', 'edge_eta_quarter_phenomhm_lvk'])
def test_edge_eta_exactly_one_quarter(name):
    """eta = 0.25: Seta = sqrt(where(eta < 0.25, 1 - 4 eta, 0)) = 0, and its tangent is 0 -- jax.jacrev's result (the select's
    transpose drops the infinite cotangent of sqrt at 0; oracle/dual.py:_sqrt restates that for the forward-mode shim) and the
    engine's definition.  Every entry of the reference's Fisher is finite and must match, the eta row included."""
    cfg, ev, out = load_golden(name)
    net = make_network('engine', cfg)
    quarter = ev['eta'] == 0.25
    assert quarter.any() and not quarter.all()
    assert np.all(np.isfinite(out['fisher'])) and np.all(np.isfinite(out['snr']))
    snr, F = net.SNR(copy_events(ev)), net.FisherMatr(copy_events(ev))
    assert np.all(np.isfinite(F))
    assert snr_err(snr, out['snr']) < SNR_RTOL and fisher_err(F, out['fisher']) < FISHER_TOL
    assert not net.last_status.any()
    if cfg['model']['cls'] != 'IMRPhenomHM':
        port = make_network('port', cfg)
        assert fisher_err(F, port.FisherMatr(copy_events(ev))) < FISHER_TOL
        assert snr_err(snr, port.SNR(copy_events(ev))) < SNR_RTOL


@pytest.mark.parametrize('name', ['edge_highspin_phenomd_et2ce', 'edge_highspin_phenomhm_lvk'])
def test_edge_high_aligned_spins(name):
    """|chi| in [0.9, 0.99]: gamma2 >= 1, fpeak takes the fabs branch (waveforms.py:1134, 1193); forward-mode values are finite."""
    cfg, ev, out = load_golden(name)
    assert np.all(np.abs(ev['chi1z']) >= 0.9) and np.all(np.abs(ev['chi2z']) >= 0.9)
    net = make_network('engine', cfg)
    snr, F = net.SNR(copy_events(ev)), net.FisherMatr(copy_events(ev))
    assert np.all(np.isfinite(out['fisher']))
    assert snr_err(snr, out['snr']) < SNR_RTOL and fisher_err(F, out['fisher']) < FISHER_TOL


def test_edge_lambda_zero():
    """Lambda = 0 and 0 < Lambda < 1: polynomial branch of the spin-induced quadrupole (waveforms.py:779, 1394), kappa2T -> 0."""
    cfg, ev, out = load_golden('edge_lambda_zero_tf2_etsl')
    assert (ev['Lambda1'] == 0).any() and (ev['Lambda2'] == 0).any()
    net = make_network('engine', cfg)
    assert np.all(np.isfinite(out['fisher']))
    assert snr_err(net.SNR(copy_events(ev)), out['snr']) < SNR_RTOL
    assert fisher_err(net.FisherMatr(copy_events(ev)), out['fisher']) < FISHER_TOL
    cfg, ev, out = load_golden('edge_lambda_zero_nrtidal_et2ce')
    net = make_network('engine', cfg)
    snr, F = net.SNR(copy_events(ev)), net.FisherMatr(copy_events(ev))
    assert np.all(np.isfinite(F)) and np.all(np.isfinite(out['fisher_masked']))
    # NRTidalv2 is compared with the reference's last grid sample masked (taper end), and within the
    # artefact's size with the raw reference
    assert snr_err(snr, out['snr_masked']) < SNR_RTOL and fisher_err(F, out['fisher_masked']) < FISHER_TOL
    assert snr_err(snr, out['snr']) < 1e-5 and fisher_err(F, out['fisher']) < 5e-3
