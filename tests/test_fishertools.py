"""fisherTools row: batched covariance / conditioning.

Golden vectors: tests/golden/cov_*.npz, produced by oracle/make_golden_cov.py from the UNMODIFIED reference's
fisherTools.CovMatr / CheckFisher (mpmath) on Fisher matrices of the committed fixtures.  The reference inverts in ~53-bit
arithmetic, so its own result carries an error of order eps * cond(normalised Fisher); the engine works in double-double.
Tolerance (written here): |C - C_ref|_ij <= 2 * 2.2e-16 * cond_norm * sqrt(C_ii C_jj)  -- i.e. within the reference's own
rounding -- plus, independently of the reference, the inverse property checked in extended precision.
CPU tests drive csrc/covariance.cuh on the host (tests/emu); GPU tests call gwfast_b200.fisherTools through the C ABI."""
import shutil

import numpy as np
import pytest

from conftest import load_golden  # noqa: F401  (path setup)
import os

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
EPS = 2.2e-16


def _gold(tag):
    return np.load(os.path.join(GOLD, 'cov_%s.npz' % tag))


def _rel(C, Cref):
    dg = np.sqrt(np.einsum('iin->in', Cref))
    return np.max(np.abs(C - Cref) / (dg[:, None, :] * dg[None, :, :]), axis=(0, 1))


def _inverse_defect(F, C):
    """max |C_n F_n - 1| in the diagonally normalised variables, evaluated in long double."""
    dg = np.sqrt(np.einsum('iin->in', F)).astype(np.longdouble)
    Fn = F.astype(np.longdouble) / (dg[:, None, :] * dg[None, :, :])
    Cn = C.astype(np.longdouble) * (dg[:, None, :] * dg[None, :, :])
    P = np.einsum('ikn,kjn->ijn', Cn, Fn)
    return np.max(np.abs(P - np.eye(F.shape[0])[:, :, None]), axis=(0, 1)).astype(float)


def _check_against_golden(cov_fn, eig_fn, tag):
    z = _gold(tag)
    F = z['fisher']
    cov, err, status = cov_fn(F, 'cho')
    assert np.all(status == 0)                                     # all positive definite -> Cholesky route
    assert np.all(_rel(cov, z['cov_cho']) <= 2 * EPS * z['cond_norm'])
    assert np.array_equal(cov, cov.transpose(1, 0, 2))
    # the inverse property, independent of the reference: defect ~ eps * cond from rounding the result to double
    assert np.all(_inverse_defect(F, cov) <= 4 * EPS * z['cond_norm'] * F.shape[0])
    # the reported inversion error (un-normalised variables, like the reference's) is the rounding of Cov to float64 and nothing
    # else: bounded by eps * max_ij sum_k |C_ik||F_kj|; the reference's own eps obeys the same bound (its Cov is float64 as well)
    bound = EPS * np.max(np.einsum('ikn,kjn->ijn', np.abs(cov), np.abs(F)), axis=(0, 1))
    assert np.all(err <= bound) and np.all(z['eps_cho'] <= 4 * bound)
    cov_s, _, status_s = cov_fn(F, 'svd')
    assert np.all(status_s == 1)
    assert np.all(_rel(cov_s, z['cov_svd']) <= 2 * EPS * z['cond_norm'])
    assert np.all(_rel(cov_s, cov) <= 1e-15 * z['cond_norm'])       # the two routes agree far below the reference's accuracy
    ev, vec, cond = eig_fn(F)
    scale = np.abs(z['evals']).max(axis=1, keepdims=True)
    assert np.max(np.abs(ev - z['evals']) / scale) < 1e-14
    ok = z['cond'] < 1e13                                           # beyond that the reference's smallest eigenvalue is rounding noise
    assert np.all(np.abs(cond[ok] / z['cond'][ok] - 1) < 1e-15 * z['cond'][ok] * 10)
    rec = np.einsum('nik,nk,njk->ijn', vec, ev, vec)
    assert np.max(np.abs(rec - F).max(axis=(0, 1)) / np.abs(F).max(axis=(0, 1))) < 1e-14
    assert np.all(np.diff(ev, axis=1) >= 0)


def _special_cases(cov_fn):
    rng = np.random.default_rng(5)
    n, nP = 6, 7
    A = rng.normal(size=(nP, nP, n))
    F = np.einsum('ikn,jkn->ijn', A, A) + 0.1 * np.eye(nP)[:, :, None]
    F[:, :, 1] = np.nan                                              # all-NaN Fisher -> NaN covariance (fisherTools.py:64-67)
    # an indefinite matrix with a positive diagonal: Cholesky breaks down, the eigen route returns the exact inverse
    B = rng.normal(size=(nP, nP))
    Q, _ = np.linalg.qr(B)
    lam = np.array([3., 2., 1.5, 1., 0.5, 0.2, -0.05])
    F[:, :, 2] = (Q * lam) @ Q.T + 2.0 * np.eye(nP) * 0
    d = np.diag(F[:, :, 2]).copy()
    assert np.all(d > 0)
    F[:, :, 3] = np.diag(np.arange(1., nP + 1))                     # diagonal
    F[0, :, 4] = 0.
    F[:, 0, 4] = 0.                                                 # zero on the diagonal: normalisation skipped (fisherTools.py:97-99)
    F[0, 0, 4] = 0.
    cov, err, status = cov_fn(F, 'cho')
    assert status[0] == 0 and status[1] == 2 and status[2] == 1 and status[3] == 0 and status[5] == 0
    assert np.all(np.isnan(cov[:, :, 1])) and np.isnan(err[1])
    assert np.allclose(cov[:, :, 2], np.linalg.inv(F[:, :, 2]), rtol=1e-10, atol=1e-12)
    assert np.allclose(cov[:, :, 3], np.diag(1. / np.arange(1., nP + 1)), rtol=1e-15, atol=0)
    assert np.allclose(cov[:, :, 0] @ F[:, :, 0], np.eye(nP), atol=1e-10)
    assert status[4] in (3, 4)                                      # singular: reported, never an exception
    # svd_reg: singular values <= thresh are dropped -> pseudo-inverse
    Fs = np.zeros((3, 3, 1))
    Fs[:, :, 0] = np.diag([1., 1., 1e-20]) + 0.
    Fs[0, 1, 0] = Fs[1, 0, 0] = 0.5
    covr, _, _ = cov_fn(Fs, 'svd_reg', 1e-10)
    # normalised matrix: diag(1,1,1) with 0.5 coupling -- the third direction has unit normalised eigenvalue, nothing dropped
    assert np.allclose(covr[:, :, 0] @ Fs[:, :, 0], np.eye(3), atol=1e-12)


@pytest.mark.skipif(shutil.which('nvcc') is None, reason='nvcc needed to build the emulation harness')
@pytest.mark.parametrize('tag', ['c2', 'c1', 'c3'])
def test_emulated_covariance_matches_reference(tag):
    import emu_driver as E
    codes = {'cho': 0, 'svd': 1, 'svd_reg': 3}
    _check_against_golden(lambda F, m, t=1e-15: E.covariance(F, codes[m], t),
                          lambda F: (lambda ev, vec, cond: (ev.T.copy(), vec.transpose(2, 0, 1), cond))(*E.eigen(F)), tag)


@pytest.mark.skipif(shutil.which('nvcc') is None, reason='nvcc needed to build the emulation harness')
def test_emulated_covariance_special_cases():
    import emu_driver as E
    codes = {'cho': 0, 'svd': 1, 'svd_reg': 3}
    _special_cases(lambda F, m, t=1e-15: E.covariance(F, codes[m], t))


def test_host_bookkeeping_helpers():
    """fixParams / addPrior / log dL / localisation area only re-index and rescale: checked against direct numpy expressions."""
    from gwfast_b200 import fisherTools as T
    rng = np.random.default_rng(1)
    nP, n = 5, 4
    pn = {'Mc': 0, 'eta': 1, 'dL': 2, 'theta': 3, 'phi': 4}
    A = rng.normal(size=(nP, nP, n))
    F = np.einsum('ikn,jkn->ijn', A, A)
    G, pn2 = T.fixParams(F, pn, ['eta', 'theta'])
    assert pn2 == {'Mc': 0, 'dL': 1, 'phi': 2} and G.shape == (3, 3, n)
    assert np.array_equal(G, F[np.ix_([0, 2, 4], [0, 2, 4], range(n))])
    G2, _ = T.fixParams(F[:, :, 0], pn, ['dL'])
    assert np.array_equal(G2, np.delete(np.delete(F[:, :, 0], 2, 0), 2, 1))
    P = T.addPrior(F, [2., 3.], pn, ['phi', 'Mc'])                  # values go to the positions in increasing order (Mc, phi)
    D = P - F
    assert np.allclose(D[0, 0], 2.) and np.allclose(D[4, 4], 3.) and np.count_nonzero(D[:, :, 0]) == 2
    ev = {'dL': rng.uniform(1, 3, n), 'theta': rng.uniform(0.2, 2.5, n)}
    L = T.log_dL_to_dL_derivative_fish(F, pn, ev)
    J = np.ones((nP, n))
    J[2] = 1 / ev['dL']
    assert np.allclose(L, F * J[:, None, :] * J[None, :, :])
    Cc = T.log_dL_to_dL_derivative_cov(F, pn, ev)
    assert np.allclose(Cc, F / (J[:, None, :] * J[None, :, :]))
    area = T.compute_localization_region(F, pn, ev['theta'])
    want = (180 / np.pi) ** 2 * (-np.log(0.1)) * 2 * np.pi * np.sqrt(F[3, 3] * F[4, 4] - F[3, 4] ** 2) * np.abs(np.sin(ev['theta']))
    assert np.allclose(area, want)
    assert np.allclose(T.compute_localization_region(F, pn, ev['theta'], units='Sterad') * (180 / np.pi) ** 2, area)


@pytest.mark.gpu
@pytest.mark.parametrize('tag', ['c2', 'c1', 'c3'])
def test_gpu_covariance_matches_reference(tag):
    from gwfast_b200 import fisherTools as T
    _check_against_golden(lambda F, m, t=1e-15: T.CovMatr(F, invMethodIn=m, svals_thresh=t, return_status=True), T.CheckFisher, tag)


@pytest.mark.gpu
def test_gpu_covariance_special_cases_and_api():
    from gwfast_b200 import fisherTools as T
    _special_cases(lambda F, m, t=1e-15: T.CovMatr(F, invMethodIn=m, svals_thresh=t, return_status=True))
    z = _gold('c2')
    cov, eps = T.CovMatr(z['fisher'])                                # reference signature: (Cov, eps)
    assert cov.shape == z['fisher'].shape and eps.shape == (z['fisher'].shape[-1],)
    assert np.allclose(T.compute_inversion_error(z['fisher'], cov), eps, rtol=1e-12, atol=1e-15)
    with pytest.raises(ValueError):
        T.CovMatr(z['fisher'], invMethodIn='qr')
    # truncate=True only touches events whose normalised condition number exceeds condNumbMax
    c_tr, _ = T.CovMatr(z['fisher'], invMethodIn='svd', truncate=True, condNumbMax=1e50)
    c_sv, _ = T.CovMatr(z['fisher'], invMethodIn='svd')
    assert np.array_equal(c_tr, c_sv)
    c_tr2, _ = T.CovMatr(z['fisher'], invMethodIn='svd', truncate=True, condNumbMax=1.0, svals_thresh=1e-3)
    dg = np.sqrt(np.einsum('iin->in', z['fisher']))
    lam = np.linalg.eigvalsh((c_tr2 * (dg[:, None, :] * dg[None, :, :])).transpose(2, 0, 1))
    assert np.all(lam.max(axis=1) / lam.min(axis=1) < 1.0001e3)       # floor at 1e-3 of the largest singular value


@pytest.mark.gpu
def test_gpu_covariance_of_full_catalog():
    """BASELINE.json configs[1] at full size: Fisher (engine) -> covariance for 10^4 events; inverse property in the normalised
    variables and the sky-localisation areas are finite and positive."""
    from gwfast_b200 import fisherTools as T, synthetic, waveforms
    from conftest import make_network, copy_events
    cfg = dict(model=dict(cls='IMRPhenomD'), network='ET+2CE', rot=True, fmin=2.)
    net = make_network('engine', cfg)
    ev = synthetic.bbh_catalog(10000, synthetic.SEEDS['C2'])
    F = net.FisherMatr(copy_events(ev))
    cov, eps, status = T.CovMatr(F, return_status=True)
    assert np.all(np.isfinite(cov)) and np.all(status <= 1)
    _, _, cn = T.CheckFisher(F / (np.sqrt(np.einsum('iin->in', F))[:, None, :] * np.sqrt(np.einsum('iin->in', F))[None, :, :]))
    assert np.all(_inverse_defect(F, cov) <= 4 * EPS * cn * 11)
    area = T.compute_localization_region(cov, waveforms.IMRPhenomD().ParNums, ev['theta'])
    assert np.all(np.isfinite(area)) and np.all(area > 0)
