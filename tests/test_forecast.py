"""Catalog driver: compute_errs / batching / resume of gwfast_b200.forecast, the counterpart of the
reference's run/calculate_forecasts_from_catalog.py (:410-634, :738, :900-1007)."""
import os

import numpy as np
import pytest

from conftest import make_network, copy_events


def test_batch_ranges_and_catalog_io(tmp_path):
    from gwfast_b200 import forecast as fc
    assert fc.batch_ranges(10, 4) == [(0, 4), (4, 8), (8, 10)]
    assert fc.batch_ranges(10, 4, 2, 9) == [(2, 6), (6, 9)]
    assert fc.batch_ranges(0, 4) == []
    ev = {'Mc': np.array([1., 2., 3.]), 'eta': np.array([.2, .21, .22])}
    np.savez(tmp_path / 'c.npz', **ev)
    got = fc.load_catalog(str(tmp_path / 'c.npz'))
    assert set(got) == set(ev) and np.array_equal(got['eta'], ev['eta'])
    with open(tmp_path / 'c.txt', 'w') as fh:
        fh.write('# Mc eta\n1 .2\n2 .21\n3 .22\n')
    got = fc.load_catalog(str(tmp_path / 'c.txt'))
    assert np.array_equal(got['Mc'], ev['Mc']) and np.allclose(got['eta'], ev['eta'])
    with pytest.raises(ValueError):
        fc.load_catalog(str(tmp_path / 'c.csv'))
    assert fc.get_events_subset(ev, np.array([True, False, True]))['Mc'].tolist() == [1., 3.]


@pytest.mark.gpu
def test_compute_errs_is_the_composition_of_the_api_calls():
    from gwfast_b200 import forecast as fc, fisherTools as ft, synthetic
    cfg = dict(model=dict(cls='IMRPhenomD'), network='ET+2CE', rot=True, fmin=2.)
    net = make_network('engine', cfg)
    ev = synthetic.bbh_catalog(300, 99)
    th = 30.
    snrs_all, Fres, eps, cov, sky, cond, idxs = fc.compute_errs(copy_events(ev), net, snr_th=th, i_in=1000)
    snr = net.SNR(copy_events(ev), return_all=True)['net']
    det = snr > th
    assert np.allclose(snr, net.SNR(copy_events(ev)), rtol=1e-14, atol=0)
    assert np.array_equal(snrs_all['net'], snr) and np.array_equal(np.ravel(idxs), 1000 + np.flatnonzero(det))
    sub = {k: v[det] for k, v in ev.items()}
    F = net.FisherMatr(copy_events(sub))
    assert set(Fres) == {'ET_0', 'ET_1', 'ET_2', 'CE1Id', 'CE2NM', 'net'}
    dg = np.sqrt(np.einsum('iin->in', F))
    assert np.max(np.abs(Fres['net'] - F) / (dg[:, None] * dg[None, :])) < 1e-12
    c2, e2 = ft.CovMatr(Fres['net'])
    assert np.array_equal(cov, c2) and np.array_equal(eps, e2)
    wf = net.signals['ET'].wf_model
    assert np.array_equal(sky, ft.compute_localization_region(cov, wf.ParNums, sub['theta']))
    assert cond.shape == (det.sum(),) and np.all(cond > 1)
    # fixed parameters shrink every matrix; nothing detected -> NaN placeholders of the right shape (run script :497-505)
    _, Ff, _, covf, _, _, _ = fc.compute_errs(copy_events(ev), net, snr_th=th, params_fix=['chi1z', 'chi2z'])
    assert Ff['net'].shape == (9, 9, det.sum()) and covf.shape == (9, 9, det.sum())
    out = fc.compute_errs(copy_events(ev), net, snr_th=1e9)
    assert out[1].shape == (11, 11, 0) and len(np.ravel(out[-1])) == 0
    # duty factor: masks per arm from the seeded numpy RNG; the network SNR is rebuilt from the masked arms
    s_all, F_d, *_ = fc.compute_errs(copy_events(ev), net, snr_th=th, duty_factor=0.5, seeds=[1, 2, 3])
    np.random.seed(2)
    m = np.random.choice([0, 1], 300, p=[0.5, 0.5])
    full = net.SNR(copy_events(ev), return_all=True)
    assert np.array_equal(s_all['CE1Id'], full['CE1Id'] * m)
    assert np.allclose(s_all['net'] ** 2, sum(s_all[k] ** 2 for k in s_all if k != 'net'))
    # SNR derivatives ride along
    out = fc.compute_errs(copy_events(ev), net, snr_th=th, return_snr_derivatives=True)
    assert len(out) == 8 and out[2]['net'].shape == (11, det.sum())


@pytest.mark.gpu
def test_run_catalog_batches_resume_and_collect(tmp_path):
    from gwfast_b200 import forecast as fc, synthetic
    cfg = dict(model=dict(cls='TaylorF2_RestrictedPN', kw=dict(use_3p5PN_SpinHO=True)), network='ET', rot=True, fmin=2.)
    net = make_network('engine', cfg)
    ev = synthetic.bns_catalog(250, 5)
    fout = str(tmp_path / 'run')
    files = fc.run_catalog(copy_events(ev), net, fout, batch_size=100, snr_th=8., verbose=False)
    assert [os.path.basename(f) for f in files] == ['batch_0_to_100.npz', 'batch_100_to_200.npz', 'batch_200_to_250.npz']
    one = fc.collect(fout)
    snr = net.SNR(copy_events(ev), return_all=True)['net']
    assert np.array_equal(one['snrs'], snr) and np.array_equal(one['idxs_detected'], np.flatnonzero(snr > 8.))
    assert one['errors'].shape == (11, (snr > 8.).sum()) and np.all(np.isfinite(one['errors']))
    # two "ranks" cover the same batches between them; resume leaves existing files untouched
    fout2 = str(tmp_path / 'run2')
    f0 = fc.run_catalog(copy_events(ev), net, fout2, batch_size=100, snr_th=8., rank=0, world=2, verbose=False)
    f1 = fc.run_catalog(copy_events(ev), net, fout2, batch_size=100, snr_th=8., rank=1, world=2, verbose=False)
    assert len(f0) == 2 and len(f1) == 1
    two = fc.collect(fout2)
    assert np.array_equal(two['snrs'], one['snrs']) and np.array_equal(two['errors'], one['errors'])
    mt = os.path.getmtime(f0[0])
    fc.run_catalog(copy_events(ev), net, fout2, batch_size=100, snr_th=8., resume=True, verbose=False)
    assert os.path.getmtime(f0[0]) == mt
