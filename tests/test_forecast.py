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


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['forecast_c2_200', 'forecast_lvk_duty_fix'])
def test_compute_errs_matches_the_reference_run_script(name):
    """forecast.compute_errs against the UNMODIFIED reference's compute_errs (run/calculate_forecasts_from_catalog.py:410-635, golden from
    oracle/make_golden_forecast.py): SNRs per arm and network, detected indices, per-arm and network Fisher matrices, and -- through
    the covariance of the network Fisher -- parameter errors, 90 % sky areas, condition numbers and inversion errors.

    Tolerances: SNR 1e-9 and Fisher 1e-6 (north star).  The covariance is the inverse of a matrix whose (diagonal-normalised) condition
    number is 1e6-1e12 here, so the Fisher tolerance does not carry over directly: errors and sky areas are held to 1e-5 relative on the events whose normalised
    condition number is below 5e7 (a perturbation bound of 1e-9 * cond < 5 %; the Fisher matrices themselves agree to ~1e-10); the
    rest is too ill-conditioned for a meaningful comparison and is left out (at least half of the events must remain)."""
    from conftest import load_golden, fisher_err
    from gwfast_b200 import forecast as fc
    cfg, ev, out = load_golden(name)
    net = make_network('engine', cfg)
    res = fc.compute_errs(copy_events(ev), net, snr_th=cfg['snr_th'], duty_factor=cfg['duty_factor'], seeds=cfg['seeds'], params_fix=tuple(cfg['params_fix']))
    snrs_all, Fres, eps, cov, sky, cond, idxs = res
    for k in snrs_all:
        ref = out['snr__' + k]
        assert np.array_equal(ref == 0, snrs_all[k] == 0), k                      # duty-factor masks: same draws, same order
        nz = ref != 0
        assert np.max(np.abs(snrs_all[k][nz] / ref[nz] - 1)) < 1e-9, k
    assert np.array_equal(np.ravel(idxs), out['idxs_detected'])
    assert set(Fres) == {k[len('fisher__'):] for k in out if k.startswith('fisher__')}
    for k in Fres:
        ref = out['fisher__' + k]
        assert Fres[k].shape == ref.shape
        live = np.einsum('iin->n', np.abs(ref)) > 0                               # arms switched off by the duty factor are all-zero
        assert np.array_equal(Fres[k][..., ~live], ref[..., ~live])
        if live.any():
            assert fisher_err(Fres[k][..., live], ref[..., live]) < 1e-6, k
    # conditioning of the diagonal-normalised network Fisher, per event
    F = out['fisher__net']
    dg = np.sqrt(np.einsum('iin->in', F))
    with np.errstate(all='ignore'):
        condn = np.array([np.linalg.cond(F[..., i] / np.outer(dg[:, i], dg[:, i])) if np.all(dg[:, i] > 0) else np.inf for i in range(F.shape[-1])])
    tol = 1e-9 * condn
    err = np.sqrt(np.einsum('iin->in', cov))
    ok = np.isfinite(out['errors']).all(axis=0) & np.isfinite(out['sky_area_90']) & (tol < 5e-2)
    assert ok.sum() >= 0.5 * len(ok)                                             # the rest is too ill-conditioned for any comparison
    d_err = np.max(np.abs(err[:, ok] / out['errors'][:, ok] - 1), axis=0)
    d_sky = np.abs(sky[ok] / out['sky_area_90'][ok] - 1)
    print('%s: %d of %d events compared; max |d sigma/sigma| / (1e-9 cond) = %.3g, sky = %.3g; worst |d sigma/sigma| = %.2e' % (
        name, ok.sum(), len(ok), np.max(d_err / tol[ok]), np.max(d_sky / tol[ok]), np.max(d_err)))
    assert np.all(d_err < 1e-5) and np.all(d_sky < 1e-5)                          # measured on B200: 2.5e-8 / 1.2e-7 (errors), 2.5e-8 (sky areas)
    assert np.all(np.abs(np.log10(cond[ok] / out['cond_numbers'][ok])) < 1e-3 + tol[ok])
    assert cov.shape == out['cov'].shape and eps.shape == out['eps'].shape


@pytest.mark.gpu
def test_run_catalog_in_the_reference_file_layout(tmp_path):
    """--reference_files: the per-batch files of the reference script (snrs_<a>_to_<b>.txt, fishers/covs .npy, sky_area, errors,
    inversion_errors, cond_numbers, idxs_det .txt; run script :224-247, :784-788), the name --resume_run looks for (:738-741), and their
    concatenation equal to the .npz run."""
    from gwfast_b200 import forecast as fc, synthetic
    cfg = dict(model=dict(cls='TaylorF2_RestrictedPN', kw=dict(use_3p5PN_SpinHO=True)), network='ET', rot=True, fmin=2.)
    net = make_network('engine', cfg)
    ev = synthetic.bns_catalog(250, 5)
    fout = str(tmp_path / 'ref')
    th = 8.
    files = fc.run_catalog(copy_events(ev), net, fout, batch_size=100, snr_th=th, verbose=False, reference_files=True)
    assert [os.path.basename(f) for f in files] == ['snrs_0_to_100.txt', 'snrs_100_to_200.txt', 'snrs_200_to_250.txt']
    for stem in fc.REFERENCE_FILES:
        assert os.path.exists(os.path.join(fout, '%s_0_to_100.%s' % (stem, 'npy' if stem in ('fishers', 'covs') else 'txt'))), stem
    cat = fc.concatenate_reference_files(fout, fc.batch_ranges(250, 100))
    other = str(tmp_path / 'npz')
    fc.run_catalog(copy_events(ev), net, other, batch_size=100, snr_th=th, verbose=False)
    one = fc.collect(other)
    assert np.allclose(cat['snrs'], one['snrs'], rtol=1e-15) and np.array_equal(cat['idxs_det'].astype(int), one['idxs_detected'])
    assert np.allclose(cat['errors'], one['errors'], rtol=1e-15) and cat['fishers'].shape == (11, 11, len(one['idxs_detected']))
    mt = os.path.getmtime(files[0])
    fc.run_catalog(copy_events(ev), net, fout, batch_size=100, snr_th=th, verbose=False, reference_files=True, resume=True)
    assert os.path.getmtime(files[0]) == mt
