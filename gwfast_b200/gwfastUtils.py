"""Host-side helpers on the hot path (drop-in for the corresponding functions of gwfast/gwfastUtils.py).

These act on per-event scalars only (O(N) numpy on the host, the same work the reference does before it
enters its O(N*res) array code); everything O(N*res) runs in the CUDA library.
"""
import numpy as np


def ra_dec_from_th_phi_rad(theta, phi):
    """gwfastUtils.py:151-164."""
    return phi, 0.5 * np.pi - theta


def _seta(eta):
    return np.sqrt(np.where(eta < 0.25, 1.0 - 4.0 * eta, 0.))


def Lamt_delLam_from_Lam12(Lambda1, Lambda2, eta):
    """(Lambda1, Lambda2, eta) -> (LambdaTilde, deltaLambda); gwfastUtils.py:398-417."""
    eta2 = eta * eta
    Seta = _seta(eta)
    Lamt = (8. / 13.) * ((1. + 7. * eta - 31. * eta2) * (Lambda1 + Lambda2) + Seta * (1. + 9. * eta - 11. * eta2) * (Lambda1 - Lambda2))
    delLam = 0.5 * (Seta * (1. - 13272. / 1319. * eta + 8944. / 1319. * eta2) * (Lambda1 + Lambda2)
                    + (1. - 15910. / 1319. * eta + 32850. / 1319. * eta2 + 3380. / 1319. * eta2 * eta) * (Lambda1 - Lambda2))
    return Lamt, delLam


def Lam12_from_Lamt_delLam(Lamt, delLam, eta):
    """(LambdaTilde, deltaLambda, eta) -> (Lambda1, Lambda2); gwfastUtils.py:419-448."""
    eta2 = eta * eta
    Seta = _seta(eta)
    mLp = (8. / 13.) * (1. + 7. * eta - 31. * eta2)
    mLm = (8. / 13.) * Seta * (1. + 9. * eta - 11. * eta2)
    mdp = Seta * (1. - (13272. / 1319.) * eta + (8944. / 1319.) * eta2) * 0.5
    mdm = (1. - (15910. / 1319.) * eta + (32850. / 1319.) * eta2 + (3380. / 1319.) * (eta2 * eta)) * 0.5
    det = (306656. / 1319.) * (eta ** 5) - (5936. / 1319.) * (eta ** 4)
    return ((mdp - mdm) * Lamt + (mLm - mLp) * delLam) / det, ((-mdm - mdp) * Lamt + (mLm + mLp) * delLam) / det


def m1m2_from_Mceta(Mc, eta):
    """gwfastUtils.py:450-464."""
    Seta = _seta(eta)
    M = Mc / (eta ** (3. / 5.))
    return 0.5 * M * (1. + Seta), 0.5 * M * (1. - Seta)


def Mceta_from_m1m2(m1, m2):
    """gwfastUtils.py:466-479."""
    return ((m1 * m2) ** (3. / 5.)) / ((m1 + m2) ** (1. / 5.)), (m1 * m2) / ((m1 + m2) * (m1 + m2))


def GPSt_to_LMST(t_GPS, lat, long):
    """GPS time -> local mean sidereal time as a day fraction (gwfastUtils.py:735-774 uses astropy, absent here)."""
    raise NotImplementedError('tGPS conversion needs astropy, which is not part of this engine; provide tcoal (GMST day fraction)')


def check_evparams(evParams):
    """Fill derived keys of the events dict IN PLACE, as the reference does (gwfastUtils.py:965-1019)."""
    if 'tcoal' not in evParams:
        try:
            print('Adding tcoal from tGPS')
            evParams['tcoal'] = GPSt_to_LMST(evParams['tGPS'], lat=0., long=0.)
        except KeyError:
            raise ValueError('One among tGPS and tcoal has to be provided.')
    if 'iota' not in evParams:
        try:
            evParams['iota'] = evParams['thetaJN']
        except KeyError:
            raise ValueError('One among iota and thetaJN has to be provided.')
    if 'Mc' not in evParams:
        try:
            print('Adding Mc and eta from the individual detector-frame masses')
            evParams['Mc'], evParams['eta'] = Mceta_from_m1m2(evParams['m1'], evParams['m2'])
        except KeyError:
            raise ValueError('Two among (Mc, eta) and (m1, m2) have to be provided.')
    if 'theta' not in evParams:
        try:
            print('Adding (theta, phi) from (ra, dec)')
            evParams['theta'] = np.pi / 2 - evParams['dec']
            evParams['phi'] = evParams['ra']
        except KeyError:
            raise ValueError('Two among (theta, phi) and (ra, dec) have to be provided.')
    return evParams
