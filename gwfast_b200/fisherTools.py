"""Fisher-matrix post-processing with the reference's interface (gwfast/fisherTools.py), SURVEY.md 8(f) #1.

The per-event work -- inversion with diagonal normalisation (``CovMatr``), eigen-decomposition / condition number
(``CheckFisher``), inversion error -- runs on the GPU in double-double arithmetic, one thread per event
(``csrc/covariance.cuh`` behind ``gwf_covariance`` / ``gwf_eigen`` / ``gwf_inversion_error``); the reference does it with an
mpmath loop over events on the host.  There is no CPU fallback.  The remaining functions only re-index or rescale the
``(nP, nP, N)`` arrays (fix parameters, add priors, dL <-> log dL, sky area) and are plain host-side array bookkeeping.
Not provided: the m1m2 / chi_eff Jacobian rotations of single 2-D matrices and the matplotlib ellipse helpers.
"""
import ctypes as C

import numpy as np

from . import _capi as K
from . import _engine

_METHODS = {'cho': 0, 'inv': 0, 'lu': 0, 'svd': 1, 'svd_reg': 3}
COV_STATUS = {0: 'cholesky', 1: 'eigen', 2: 'nan_input', 3: 'failed', 4: 'zero_diagonal'}


def _as_stack(M):
    M = np.asarray(M, dtype=np.float64)
    if M.ndim == 2:
        M = M[:, :, None]
    if M.ndim != 3 or M.shape[0] != M.shape[1]:
        raise ValueError('expected an array of shape (nP, nP, N)')
    return np.ascontiguousarray(M)


def _to_device(st, a):
    host = st.torch.from_numpy(a)
    return host.to(st.device, non_blocking=False)


def CovMatr(FisherMatrix, invMethodIn='cho', condNumbMax=1e50, truncate=False, svals_thresh=1e-15, verbose=False, alt_method='svd',
            return_status=False):
    """Covariance matrices and inversion errors, ``(nP, nP, N)`` in -> ``((nP, nP, N), (N,))`` out (fisherTools.py:32-196).

    ``'cho'`` (default), ``'inv'`` and ``'lu'`` all denote the exact inverse of the diagonally normalised matrix and share the
    Cholesky kernel; a matrix that is not positive definite takes the symmetric eigen-decomposition route, the reference's
    ``alt_method='svd'`` (for a symmetric matrix the SVD inverse equals ``V diag(1/lambda) V^T``).  ``'svd'`` with
    ``truncate=True`` raises singular values below ``svals_thresh * max`` to that floor when the condition number of the
    normalised matrix exceeds ``condNumbMax``; ``'svd_reg'`` drops singular values ``<= svals_thresh``.  One deviation: with
    ``truncate=True`` the reference measures the inversion error against its truncated (normalised) Fisher; here it is
    always measured against the input matrix.
    """
    if invMethodIn not in _METHODS:
        raise ValueError("invMethodIn must be one of 'inv', 'cho', 'svd', 'svd_reg', 'lu'")
    F = _as_stack(FisherMatrix)
    nP, _, n = F.shape
    st = _engine.state()
    torch = st.torch
    method = _METHODS[invMethodIn]
    thresh = float(svals_thresh)
    dF = _to_device(st, F)
    sp = C.c_void_p(torch.cuda.current_stream(st.device).cuda_stream)
    if method == 1 and truncate:
        # truncation only applies to events whose normalised matrix is worse conditioned than condNumbMax (fisherTools.py:137)
        cond = _normalised_condition(st, dF, n, nP, sp)
        method = 2 if bool((cond > condNumbMax).any()) else 1
    cov = torch.empty_like(dF)
    err = torch.empty(n, dtype=torch.float64, device=st.device)
    status = torch.empty(n, dtype=torch.int32, device=st.device)
    K.check(st.lib.gwf_covariance(C.c_void_p(dF.data_ptr()), n, nP, method, thresh, C.c_void_p(cov.data_ptr()), C.c_void_p(err.data_ptr()),
                                  C.c_void_p(status.data_ptr()), sp), 'gwf_covariance')
    if method == 2:
        # events below the conditioning threshold keep the plain inverse
        cov1 = torch.empty_like(dF)
        err1 = torch.empty_like(err)
        K.check(st.lib.gwf_covariance(C.c_void_p(dF.data_ptr()), n, nP, 1, thresh, C.c_void_p(cov1.data_ptr()), C.c_void_p(err1.data_ptr()),
                                      None, sp), 'gwf_covariance')
        keep = ~(cond > condNumbMax)
        cov[:, :, keep] = cov1[:, :, keep]
        err[keep] = err1[keep]
    cov_h, err_h, st_h = cov.cpu().numpy(), err.cpu().numpy(), status.cpu().numpy()
    if verbose:
        names, counts = np.unique(st_h, return_counts=True)
        print('Inversion routes: ' + ', '.join('%s: %d' % (COV_STATUS[int(k)], c) for k, c in zip(names, counts)))
        print(' Inversion error with method %s: min=%s, max=%s, mean=%s, std=%s ' % (invMethodIn, np.nanmin(err_h), np.nanmax(err_h),
                                                                                    np.nanmean(err_h), np.nanstd(err_h)))
    if return_status:
        return cov_h, err_h, st_h
    return cov_h, err_h


def _normalised_condition(st, dF, n, nP, sp):
    torch = st.torch
    dg = torch.sqrt(torch.diagonal(dF, dim1=0, dim2=1).transpose(0, 1))          # (nP, N)
    Fn = (dF / (dg[:, None, :] * dg[None, :, :])).contiguous()
    ev = torch.empty((nP, n), dtype=torch.float64, device=st.device)
    cond = torch.empty(n, dtype=torch.float64, device=st.device)
    K.check(st.lib.gwf_eigen(C.c_void_p(Fn.data_ptr()), n, nP, C.c_void_p(ev.data_ptr()), None, C.c_void_p(cond.data_ptr()), sp), 'gwf_eigen')
    return cond


def compute_inversion_error(Fisher, Cov):
    """``max |Cov @ Fisher - 1|`` per event (fisherTools.py:199-211)."""
    F, Cv = _as_stack(Fisher), _as_stack(Cov)
    if F.shape != Cv.shape:
        raise ValueError('Fisher and Cov must have the same shape')
    nP, _, n = F.shape
    st = _engine.state()
    torch = st.torch
    dF, dC = _to_device(st, F), _to_device(st, Cv)
    err = torch.empty(n, dtype=torch.float64, device=st.device)
    sp = C.c_void_p(torch.cuda.current_stream(st.device).cuda_stream)
    K.check(st.lib.gwf_inversion_error(C.c_void_p(dF.data_ptr()), C.c_void_p(dC.data_ptr()), n, nP, C.c_void_p(err.data_ptr()), sp),
            'gwf_inversion_error')
    return err.cpu().numpy()


def CheckFisher(FisherM, condNumbMax=1.0e15, use_mpmath=True, verbose=False):
    """Eigenvalues ``(N, nP)`` (ascending), eigenvectors ``(N, nP, nP)`` (``evecs[k, :, i]`` belongs to ``evals[k, i]``) and condition
    numbers ``(N,)`` of the Fisher matrices (fisherTools.py:216-279).  ``use_mpmath`` is accepted for compatibility: both of the
    reference's branches are served by the same double-double Jacobi kernel."""
    F = _as_stack(FisherM)
    nP, _, n = F.shape
    st = _engine.state()
    torch = st.torch
    dF = _to_device(st, F)
    ev = torch.empty((nP, n), dtype=torch.float64, device=st.device)
    vec = torch.empty((nP, nP, n), dtype=torch.float64, device=st.device)
    cond = torch.empty(n, dtype=torch.float64, device=st.device)
    sp = C.c_void_p(torch.cuda.current_stream(st.device).cuda_stream)
    K.check(st.lib.gwf_eigen(C.c_void_p(dF.data_ptr()), n, nP, C.c_void_p(ev.data_ptr()), C.c_void_p(vec.data_ptr()), C.c_void_p(cond.data_ptr()), sp),
            'gwf_eigen')
    evals = ev.cpu().numpy().T.copy()
    evecs = np.ascontiguousarray(vec.cpu().numpy().transpose(2, 0, 1))
    condNumber = cond.cpu().numpy()
    if np.any(evals <= 0.):
        print('WARNING: one or more eigenvalues are negative at position(s) %s' % str(np.unique(np.where(evals < 0)[0])))
    if np.any(condNumber > condNumbMax) and verbose:
        print('WARNING: the condition number is too large (%s>%s)' % (condNumber, condNumbMax))
        print('Unreliable covariance at positions ' + str(condNumber > condNumbMax))
    elif verbose:
        print('Condition number= %s . Ok. ' % condNumber)
    return evals, evecs, condNumber


def perturb_Fisher(totF, eps=1e-10, **kwargs):
    """Print how much the covariance moves when the Fisher is perturbed at the ``eps`` level (fisherTools.py:282-299)."""
    base, _ = CovMatr(totF, **kwargs)
    pert, _ = CovMatr(np.asarray(totF) + np.random.rand(*np.shape(totF)) * eps, **kwargs)
    epsErr = [np.linalg.norm(base[i] / pert[i] - 1, ord=np.inf) for i in range(pert.shape[-1])]
    print('Relative errors when perturbing at the %s level: %s' % (eps, epsErr))
    return epsErr


def check_covariance(FisherM, Cov, tol=1e-10):
    """Products ``Cov @ Fisher`` per event, with the reference's printed diagnostics (fisherTools.py:302-334)."""
    F, Cv = _as_stack(FisherM), _as_stack(Cov)
    ids = np.einsum('ikn,kjn->ijn', Cv, F)
    print('Inversion errors: %s' % compute_inversion_error(F, Cv))
    nP = F.shape[0]
    print('diagonal-1 = %s' % str([np.diagonal(ids[:, :, i]) - 1 for i in range(ids.shape[-1])]))
    off = ids[~np.eye(nP, dtype=bool)]
    print('Max off diagonal: %s' % str(list(off.max(axis=0))))
    print('\nmask: where F*S(off-diagonal)>%s (--> problematic if True off diagonal)' % tol)
    print([ids[:, :, i] > tol for i in range(ids.shape[-1])])
    return ids


# ---------------------------------------------------------------------------------------------- array bookkeeping
def fixParams(MatrIn, ParNums_inp, ParMarg):
    """Drop the rows/columns of the parameters in ``ParMarg`` (fix them to their fiducial values); returns the reduced array and the
    re-numbered ParNums dict (fisherTools.py:340-376)."""
    M = np.asarray(MatrIn)
    drop = sorted(ParNums_inp[p] for p in ParMarg)
    keep = [i for i in range(M.shape[0]) if i not in drop]
    out = M[np.ix_(keep, keep)] if M.ndim == 2 else M[np.ix_(keep, keep, range(M.shape[2]))]
    new = {k: v - sum(1 for d in drop if d < v) for k, v in ParNums_inp.items() if k not in ParMarg}
    return np.array(out, dtype=float), new


def addPrior(Matr, vals, ParNums, ParAdd):
    """Add Gaussian priors ``vals`` (inverse variances) on the diagonal entries of the parameters in ``ParAdd`` (fisherTools.py:378-407);
    as in the reference the values are assigned in increasing order of the parameters' positions."""
    M = np.asarray(Matr, dtype=float)
    diag = np.zeros(M.shape[0])
    diag[np.sort(np.array([ParNums[p] for p in ParAdd]))] = vals
    P = np.diag(diag)
    return P + M if M.ndim == 2 else P[:, :, None] + M


def log_dL_to_dL_derivative_cov(or_matrix, ParNums, evParams):
    """Covariance in log(dL) -> covariance in dL (fisherTools.py:410-433)."""
    M = np.array(or_matrix, dtype=float)
    i = ParNums['dL']
    M[:, i, ...] = M[:, i, ...] * evParams['dL']
    M[i, :, ...] = M[i, :, ...] * evParams['dL']
    return M


def log_dL_to_dL_derivative_fish(or_matrix, ParNums, evParams):
    """Fisher in log(dL) -> Fisher in dL (fisherTools.py:435-458)."""
    M = np.array(or_matrix, dtype=float)
    i = ParNums['dL']
    M[:, i, ...] = M[:, i, ...] / evParams['dL']
    M[i, :, ...] = M[i, :, ...] / evParams['dL']
    return M


def compute_localization_region(Cov, parNum, thFid, perc_level=90, units='SqDeg'):
    """Sky area at ``perc_level`` per cent from the (theta, phi) block of the covariance, Barack & Cutler gr-qc/0310125
    (fisherTools.py:830-862)."""
    it, ip = parNum['theta'], parNum['phi']
    Cov = np.asarray(Cov)
    base = 2 * np.pi * np.sqrt(Cov[it, it] * Cov[ip, ip] - Cov[ip, it] ** 2) * np.abs(np.sin(thFid))
    area = -base * np.log(1 - perc_level / 100)
    if units == 'Sterad':
        return area
    if units == 'SqDeg':
        return (180 / np.pi) ** 2 * area
    raise ValueError("units must be 'SqDeg' or 'Sterad'")
