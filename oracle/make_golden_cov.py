"""Golden vectors for the covariance row: the UNMODIFIED reference's fisherTools.CovMatr / CheckFisher /
compute_localization_region (run under the oracle shim, mpmath as installed) on Fisher matrices taken from the committed
golden fixtures.  TEST INFRASTRUCTURE; runs only where /root/reference is mounted.  Writes tests/golden/cov_*.npz."""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import reference  # noqa: E402

reference.load()
with contextlib.redirect_stdout(io.StringIO()):
    from gwfast import fisherTools as RT  # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')


def one(src, n, tag):
    z = np.load(os.path.join(GOLD, src + '.npz'), allow_pickle=True)
    F = np.array(z['fisher'] if 'fisher' in z else z['out__fisher'])[..., :n]
    out = dict(fisher=F)
    with contextlib.redirect_stdout(io.StringIO()):
        cov, eps = RT.CovMatr(F)
        out['cov_cho'], out['eps_cho'] = np.array(cov, dtype=float), np.array(eps, dtype=float)
        cov, eps = RT.CovMatr(F, invMethodIn='svd')
        out['cov_svd'], out['eps_svd'] = np.array(cov, dtype=float), np.array(eps, dtype=float)
        ev, evec, cond = RT.CheckFisher(F)
        out['evals'], out['cond'] = np.array(ev, dtype=float), np.array(cond, dtype=float)
        # conditioning of the diagonally normalised matrices (what decides the attainable accuracy of the inverse)
        dg = np.sqrt(np.einsum('iin->in', F))
        _, _, out['cond_norm'] = RT.CheckFisher(F / (dg[:, None, :] * dg[None, :, :]))
        out['cond_norm'] = np.array(out['cond_norm'], dtype=float)
    np.savez_compressed(os.path.join(GOLD, 'cov_' + tag + '.npz'), **out)
    print(tag, F.shape, 'eps max %.2e' % np.nanmax(out['eps_cho']), 'cond_norm max %.2e' % np.nanmax(out['cond_norm']))


if __name__ == '__main__':
    one('c2_phenomd_et2ce', 24, 'c2')
    one('c1_tf2_bns_etsl', 16, 'c1')
    one('c3_nrtidal_et2ce', 12, 'c3')
