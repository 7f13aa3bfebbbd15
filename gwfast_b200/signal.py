"""``GWSignal`` with the reference's constructor, attributes, keyword arguments, side effects and return shapes
(gwfast/signal.py:37-1098).  The numerical work of ``SNRInteg`` / ``FisherMatr`` -- frequency grid, PSD
interpolation, waveform, detector projection, derivatives, inner products -- is one fused CUDA launch per call
(``gwf_snr`` / ``gwf_fisher``); this file only keeps the host-side contract: dict mutation, error messages,
duty-cycle masks drawn from numpy's global RNG, ``return_all`` shapes.
"""
import numpy as onp

from . import gwfastUtils as utils
from . import gwfastGlobals as glob
from . import _capi as K
from . import _engine


def _fill_spins(wf_model, evParams, verbose):
    """signal.py:694-704 / 824-834: add chi1z, chi2z from chiS, chiA (mutates the dict)."""
    if 'chi1z' not in evParams:
        try:
            if verbose:
                print('Adding chi1z, chi2z from chiS, chiA')
            evParams['chi1z'] = evParams['chiS'] + evParams['chiA']
            evParams['chi2z'] = evParams['chiS'] - evParams['chiA']
        except KeyError:
            raise ValueError('Two among chi1z, chi2z and chiS, chiA have to be provided.')


def _num_events(evParams):
    return len(onp.atleast_1d(evParams['Mc']))


def _engine_events(wf_model, evParams, lambdas=None, use_m1m2=False, exact_cut=False):
    """the arrays the engine consumes (never mutates the caller's dict).

    Besides the dict entries this passes two per-event scalars computed HERE with the reference's own numpy expressions:
    ``wf_model.fcut(**evParams)`` (signal.py:715, 884) and ``M*GMsun_over_c3`` (waveforms.py:1026).  With them the last
    grid sample lands on ``Mf = fcutPar`` with the reference's rounding, which decides whether it is inside the waveform
    cut -- a 1e-9 effect on the IMRPhenomHM SNR (the (2,2) mode switches off while the higher modes are still on).
    """
    ev = {k: evParams[k] for k in K.EVENT_KEYS[:11]}
    if wf_model.is_tidal:
        L1, L2 = lambdas if lambdas is not None else (evParams['Lambda1'], evParams['Lambda2'])
        ev['Lambda1'], ev['Lambda2'] = L1, L2
    if wf_model.is_eccentric:
        ev['ecc'] = evParams['ecc']              # signal.py:877-880
    if wf_model.is_HigherModes or wf_model.objType == 'NSBH' or (exact_cut and not wf_model.is_tidal and type(wf_model).__name__ != 'TaylorF2_RestrictedPN'):
        # only IMRPhenomHM needs it at the 1e-9 level (3e-12 for IMRPhenomD), and IMRPhenomNSBH, whose d ln A/d Lambda is ~1e2 at the
        # cut (1e-6 of F[Lambda, Lambda] from that one sample); two numpy pow() per event on the host.
        # exact_cut: the strain-derivative output exposes the last sample itself, so IMRPhenomD asks for it there too
        ev['_fcut'] = wf_model.fcut(**evParams)
        Mc, eta = evParams['Mc'], evParams['eta']
        if use_m1m2:
            Mc, eta = utils.Mceta_from_m1m2(*utils.m1m2_from_Mceta(Mc, eta))      # GWstrain's round trip, signal.py:522-524
        ev['_Mtot_sec'] = (Mc / (eta ** (3. / 5.))) * glob.GMsun_over_c3
    return ev


def hot_snr(signals, evParams, res):
    """per-arm SNR^2 for a list of GWSignal sharing one waveform model: ONE launch for the whole list.

    Returns (snr2_arm (n_arms, N), arm_slices) where arm_slices[i] is the slice of arms of signals[i].
    """
    wf = signals[0].wf_model
    n = _num_events(evParams)
    dets, handles, slices, a0 = [], [], [], 0
    for s in signals:
        dets.append(s._detector_struct(len(handles)))
        handles.append(s._psd_handle())
        na = 1 if s.detector_shape == 'L' else 3
        slices.append(slice(a0, a0 + na))
        a0 += na
    out, io = _engine.snr(wf._descriptor(evParams), dets, handles, _engine_events(wf, evParams), n, res)
    return out, slices, io


def _fisher_flags(spacing, use_m1m2, use_chi1chi2):
    flags = 0
    if use_m1m2:
        flags |= K.GWF_OPT_M1M2
    if not use_chi1chi2:
        flags |= K.GWF_OPT_CHIS_CHIA
    if spacing == 'lin':
        flags |= K.GWF_OPT_LIN_GRID
    elif spacing != 'geom':
        raise ValueError("spacing has to be 'geom' or 'lin'")
    return flags


def hot_strain_derivs(signals, evParams, lambdas, res, spacing, use_m1m2, use_chi1chi2):
    """d h / d p_i per arm, complex128 (n_arms, nP, N, res) (return_derivatives, signal.py:917-945)."""
    wf = signals[0].wf_model
    n = _num_events(evParams)
    dets, handles = [], []
    for s in signals:
        dets.append(s._detector_struct(len(handles)))
        handles.append(s._psd_handle())
    D = _engine.strain_derivs(wf._descriptor(evParams), dets, handles, _engine_events(wf, evParams, lambdas, use_m1m2, exact_cut=True), n, res,
                              _fisher_flags(spacing, use_m1m2, use_chi1chi2))
    rows = getattr(wf, '_engine_rows', None)
    return D if rows is None else onp.ascontiguousarray(D[:, rows])


def hot_fisher(signals, evParams, lambdas, res, spacing, use_m1m2, use_chi1chi2, per_arm, want_snr_derivs=False, want_snr_integ=False):
    """Fisher matrices for a list of GWSignal sharing one waveform model: ONE prologue + one launch per block.  With
    ``want_snr_integ`` the second return value is the integral SNRInteg forms for the same arms (the fused SNR + Fisher launch)."""
    wf = signals[0].wf_model
    n = _num_events(evParams)
    flags = _fisher_flags(spacing, use_m1m2, use_chi1chi2)
    dets, handles = [], []
    for s in signals:
        dets.append(s._detector_struct(len(handles)))
        handles.append(s._psd_handle())
    F, snr2, io = _engine.fisher(wf._descriptor(evParams), dets, handles, _engine_events(wf, evParams, lambdas, use_m1m2), n, res, flags, per_arm,
                                 want_snr_derivs=want_snr_derivs, want_snr_integ=want_snr_integ)
    rows = getattr(wf, '_engine_rows', None)
    if rows is not None:
        # NewtInspiral: the engine's eta / spin rows are identically zero and are not part of the model's 8 parameters
        F = onp.ascontiguousarray(F[:, rows][:, :, rows])
        if want_snr_derivs:
            snr2 = (snr2[0], onp.ascontiguousarray(snr2[1][:, :, rows]))
    return F, snr2, io


class GWSignal(object):
    """Single detector (L-shaped or triangular); see gwfast/signal.py:37-160 for the parameters."""

    def __init__(self, wf_model, psd_path=None, detector_shape='T', det_lat=40.44, det_long=9.45, det_xax=0., verbose=True,
                 is_ASD=True, useEarthMotion=False, noMotion=False, fmin=2., fmax=None, IntTablePath=None, DutyFactor=None,
                 compute2arms=True, jitCompileDerivs=False):
        if (detector_shape != 'L') and (detector_shape != 'T'):
            raise ValueError('Enter valid detector configuration')
        if psd_path is None:
            raise ValueError('Enter a valid PSD or ASD path')
        if verbose:
            print(('Using ASD from file %s ' if is_ASD else 'Using PSD from file %s ') % psd_path)
        if useEarthMotion and (wf_model.objType == 'BBH') and verbose:
            print('WARNING: the motion of Earth gives a negligible contribution for BBH signals, consider switching it off to make the code run faster')
        if (not useEarthMotion) and (wf_model.objType in ('BNS', 'NSBH')) and verbose:
            print('WARNING: the motion of Earth gives a relevant contribution for %s signals, consider switching it on' % wf_model.objType)
        self.wf_model = wf_model
        self.psd_base_path = ('/').join(psd_path.split('/')[:-1])
        self.psd_file_name = psd_path.split('/')[-1]
        self.verbose = verbose
        self.detector_shape = detector_shape
        self.det_lat_rad = det_lat * onp.pi / 180.
        self.det_long_rad = det_long * onp.pi / 180.
        self.det_xax_rad = det_xax * onp.pi / 180.
        self.IntTablePath = IntTablePath
        self.DutyFactor = DutyFactor
        noise = onp.loadtxt(psd_path, usecols=(0, 1))
        self.strainFreq = noise[:, 0]
        self.noiseCurve = noise[:, 1] ** 2 if is_ASD else noise[:, 1]
        self.useEarthMotion = useEarthMotion
        self.noMotion = noMotion
        if self.noMotion and self.useEarthMotion:
            print('noMotion and useEarthMotion are True. switching off useEarthMotion ')
            self.useEarthMotion = False
        self.fmin = fmin
        self.fmax = fmax
        self.angbtwArms = 0.5 * onp.pi if detector_shape == 'L' else onp.pi / 3.
        self.IntegInterpArr = None
        self.compute2arms = compute2arms
        onp.random.seed(None)
        self.seedUse = onp.random.randint(2 ** 32 - 1, size=1)
        self.jitCompileDerivs = jitCompileDerivs   # accepted for compatibility; there is nothing to jit
        self._psd = None
        self.last_status = None                    # per-event status words of the last FisherMatr call (_capi.GWF_EV_*)

    # ------------------------------------------------------------------ engine hooks
    def _psd_handle(self):
        if self._psd is None:
            self._psd = _engine.psd_handle(self.strainFreq, self.noiseCurve)
        return self._psd

    def _detector_struct(self, psd_index):
        return K.gwf_detector(self.det_lat_rad, self.det_long_rad, self.det_xax_rad, 0 if self.detector_shape == 'L' else 1,
                              int(bool(self.useEarthMotion)), int(bool(self.noMotion)), psd_index, float(self.fmin),
                              float(self.fmax) if self.fmax is not None else 0.)

    def _clear_cache(self):
        pass

    def _update_seed(self, seed=None):
        """signal.py:221-232."""
        onp.random.seed(None)
        self.seedUse = onp.random.randint(2 ** 32 - 1, size=1) if seed is None else seed

    def _narms(self):
        return 1 if self.detector_shape == 'L' else 3

    def _duty_masks(self, n):
        """one Bernoulli mask per arm from numpy's global RNG, drawn in the reference's order (signal.py:729-762)."""
        return [onp.random.choice([0, 1], n, p=[1. - self.DutyFactor, self.DutyFactor]) for _ in range(self._narms())]

    # ------------------------------------------------------------------ strain-level methods (signal.py:342-655)
    def _this_detector(self):
        return K.gwf_detector(self.det_lat_rad, self.det_long_rad, self.det_xax_rad, 0 if self.detector_shape == 'L' else 1,
                              int(bool(self.useEarthMotion)), int(bool(self.noMotion)), 0, float(self.fmin),
                              float(self.fmax) if self.fmax is not None else 0.)

    def _ra_dec_from_th_phi(self, theta, phi):
        """signal.py:389-399."""
        return utils.ra_dec_from_th_phi_rad(theta, phi)

    def _PatternFunction(self, theta, phi, t, psi, rot=0.):
        """Plus and cross pattern functions of the detector (arXiv:gr-qc/9804014 eq. 10-13) for sky positions, times (GMST, days) and
        polarisation angles of any mutually broadcastable shapes; ``rot`` in degrees (signal.py:342-387).  Evaluated by ``gwf_pattern``."""
        from . import _elementwise
        r = _elementwise.pattern(self._this_detector(), rot, theta, phi, t, psi, ('Fp', 'Fc'))
        return r['Fp'], r['Fc']

    def _DeltLoc(self, theta, phi, t):
        """Time (s) to go from the Earth's centre to the detector for sky positions and times (GMST, days) (signal.py:401-423)."""
        from . import _elementwise
        return _elementwise.pattern(self._this_detector(), 0., theta, phi, t, None, ('dt',))['dt']

    def GWAmplitudes(self, evParams, f, rot=0.):
        """Plus and cross amplitudes at the detector on the grid ``f`` (``(res,)`` or ``(res, N)``); signal.py:425-466."""
        from . import _elementwise
        r = _elementwise.signal_grid(self.wf_model, self._this_detector(), rot, f, evParams, ('Ap', 'Ac'))
        return r['Ap'], r['Ac']

    def GWPhase(self, evParams, f):
        """Complete signal phase 2 pi f tcoal 86400 - Phicoal - Phi(f) on the grid ``f``; signal.py:468-484."""
        if self.wf_model.is_HigherModes:
            raise TypeError('GWPhase is not defined for a higher-mode waveform: its Phi is a dictionary of mode phases (waveforms.py:2075)')
        from . import _elementwise
        return _elementwise.signal_grid(self.wf_model, self._this_detector(), 0., f, evParams, ('psi',))['psi']

    def GWstrain(self, f, Mc, eta, dL, theta, phi, iota, psi, tcoal, Phicoal, chiS, chiA, chi1x, chi2x, chi1y, chi2y, LambdaTilde, deltaLambda,
                 ecc, rot=0., is_m1m2=False, is_chi1chi2=False, is_prec_ang=False, is_Lam1Lam2=False, return_single_comp=None):
        """Full complex strain at the detector on the grid ``f`` as a function of the parameters (signal.py:486-655): the same
        re-parametrisation switches (``is_m1m2``, ``is_chi1chi2``, ``is_Lam1Lam2``) and ``return_single_comp`` values
        ('Ap', 'Ac', 'Psip', 'Psic', 'At', 'Psit')."""
        from . import _elementwise
        if is_m1m2:
            McUse, etaUse = utils.Mceta_from_m1m2(Mc, eta)
        else:
            McUse, etaUse = Mc, eta
        if is_chi1chi2:
            chi1z, chi2z = chiS, chiA
        else:
            chi1z, chi2z = chiS + chiA, chiS - chiA
        ev = {'Mc': McUse, 'dL': dL, 'theta': theta, 'phi': phi, 'iota': iota, 'psi': psi, 'tcoal': tcoal, 'eta': etaUse, 'Phicoal': Phicoal,
              'chi1z': chi1z, 'chi2z': chi2z}
        if self.wf_model.is_tidal:
            if not is_Lam1Lam2:
                ev['Lambda1'], ev['Lambda2'] = utils.Lam12_from_Lamt_delLam(LambdaTilde, deltaLambda, etaUse)
            else:
                ev['Lambda1'], ev['Lambda2'] = LambdaTilde, deltaLambda
        if self.wf_model.is_eccentric:
            ev['ecc'] = ecc
        det = self._this_detector()
        if return_single_comp is None:
            return _elementwise.signal_grid(self.wf_model, det, rot, f, ev, ('strain',))['strain']
        if return_single_comp not in ('Ap', 'Ac', 'Psip', 'Psic', 'At', 'Psit'):
            raise ValueError('Single component to return has to be among Ap, Ac, Psip, Psic')
        if self.wf_model.is_HigherModes:
            # signal.py:590-603: moduli and unwrapped phases of the two polarisation terms hp Fp e^{i...}, hc Fc e^{i...}
            r = _elementwise.signal_grid(self.wf_model, det, rot, f, ev, ('strain', 'Fp', 'Fc', 'dt'))
            hp, hc = self.wf_model.hphc(f, **ev)
            fa = onp.asarray(f, dtype=float)
            ph = onp.exp(1j * (2. * onp.pi * fa * r['dt'] + 2. * onp.pi * fa * (onp.asarray(tcoal) * 3600. * 24.) - onp.asarray(Phicoal)))
            hp, hc = hp * r['Fp'] * ph, hc * r['Fc'] * ph
            return {'Ap': lambda: onp.abs(hp), 'Ac': lambda: onp.abs(hc), 'Psip': lambda: onp.unwrap(onp.angle(hp)),
                    'Psic': lambda: onp.unwrap(onp.angle(hc)), 'At': lambda: onp.abs(hp + hc), 'Psit': lambda: onp.unwrap(onp.angle(hp + hc))}[return_single_comp]()
        r = _elementwise.signal_grid(self.wf_model, det, rot, f, ev, ('Ap', 'Ac', 'psi', 'dt'))
        Psi = r['psi'] + 2. * onp.pi * onp.asarray(f, dtype=float) * r['dt']
        if return_single_comp == 'Ap':
            return r['Ap']
        if return_single_comp == 'Ac':
            return r['Ac']
        if return_single_comp == 'Psip':
            return Psi
        if return_single_comp == 'Psic':
            return Psi + onp.pi * 0.5
        if return_single_comp == 'At':
            return onp.abs(r['Ap'] + 1j * r['Ac'])
        return Psi + onp.arctan2(r['Ac'], r['Ap'])

    def optimal_location(self, tcoal, is_tGPS=False):
        """Optimal (theta, phi) for a signal to be seen by the detector at a given GMST (psi = 0); signal.py:1586-1614."""
        from scipy.optimize import minimize
        if is_tGPS:
            tcoal = utils.GPSt_to_LMST(tcoal, lat=0., long=0.)

        def pattern_fixedtpsi(pars, tc=tcoal):
            Fp, Fc = self._PatternFunction(pars[0], pars[1], t=tc, psi=0)
            return -onp.sqrt(Fp ** 2 + Fc ** 2)
        return minimize(pattern_fixedtpsi, [1., 1.], bounds=((0., onp.pi), (0., 2. * onp.pi))).x

    # ------------------------------------------------------------------ SNR
    def _prepare_snr(self, evParams):
        utils.check_evparams(evParams)
        _fill_spins(self.wf_model, evParams, self.verbose)
        if self.wf_model.is_tidal and 'Lambda1' not in evParams:
            try:
                evParams['Lambda1'], evParams['Lambda2'] = utils.Lam12_from_Lamt_delLam(evParams['LambdaTilde'], evParams['deltaLambda'], evParams['eta'])
            except KeyError:
                raise ValueError('Two among Lambda1, Lambda2 and LambdaTilde and deltaLambda have to be provided.')

    def _snr_from_arms(self, s2, n, return_all):
        """s2: (narms, N) per-arm 4*int(...)  ->  the reference's return value (signal.py:767-777)."""
        if self.DutyFactor is not None:
            s2 = s2 * onp.array(self._duty_masks(n))
        if self.detector_shape == 'T':
            return onp.sqrt(s2) if return_all else onp.sqrt(s2.sum(axis=0))
        return onp.sqrt(s2[0])

    def SNRInteg(self, evParams, res=1000, return_all=False):
        """SNR of the event(s), shape (N,) ((3, N) for a triangle with ``return_all``); signal.py:658-777."""
        if self.DutyFactor is not None:
            onp.random.seed(self.seedUse)
        self._prepare_snr(evParams)
        s2, _, _ = hot_snr([self], evParams, res)
        return self._snr_from_arms(s2, _num_events(evParams), return_all)

    # ------------------------------------------------------------------ Fisher
    def _prepare_fisher(self, evParams, res, df, computeDerivFinDiff, return_derivatives, return_SNR_derivatives):
        """host-side part of signal.py:812-898; returns (Lambda1, Lambda2) locals (not written to the dict) and res."""
        utils.check_evparams(evParams)
        _fill_spins(self.wf_model, evParams, self.verbose)
        lambdas = None
        if self.wf_model.is_tidal:
            try:
                lambdas = (evParams['Lambda1'], evParams['Lambda2'])
            except KeyError:
                try:
                    lambdas = utils.Lam12_from_Lamt_delLam(evParams['LambdaTilde'], evParams['deltaLambda'], evParams['eta'])
                except KeyError:
                    raise ValueError('Two among Lambda1, Lambda2 and LambdaTilde and deltaLambda have to be provided.')
        if computeDerivFinDiff:
            raise NotImplementedError('finite-difference derivatives (numdifftools) are only used by the LAL/TEOBResumS wrappers of the reference; '
                                      'this engine always differentiates exactly')
        if res is None and df is not None:
            fcut = self.wf_model.fcut(**evParams)
            if self.fmax is not None:
                fcut = onp.where(fcut > self.fmax, self.fmax, fcut)
            res = onp.amax(onp.floor(onp.real(1 + (fcut - self.fmin) / df)))      # signal.py:890-892
        elif res is None and df is None:
            raise ValueError('Provide either resolution in frequency or step size.')
        return lambdas, int(res)

    def FisherMatr(self, evParams, res=1000, df=None, spacing='geom', use_m1m2=False, use_chi1chi2=True, use_prec_ang=True,
                   computeDerivFinDiff=False, computeAnalyticalDeriv=True, return_all=False, return_derivatives=False,
                   return_SNR_derivatives=False, return_SNR=False, **kwargs):
        """Fisher matrix, shape (nParams, nParams, N) (list per arm with ``return_all``); signal.py:782-1098.

        ``computeAnalyticalDeriv`` is accepted for compatibility: the engine's derivatives are exact either way.
        ``return_SNR=True`` (an addition to the reference's keywords) returns ``(F, SNR)`` with the SNR ``SNRInteg`` would give for
        the same events, from the same launch -- the Fisher kernel already integrates |h|^2/Sn, so ``SNRInteg`` followed by
        ``FisherMatr`` costs one launch instead of two.  After the call ``self.last_status`` holds the per-event status words
        (``_capi.GWF_EV_*``; 0 = clean).
        """
        if return_SNR:
            if return_all or return_derivatives or return_SNR_derivatives or self.DutyFactor is not None:
                snr = self.SNRInteg(evParams, res=res if res is not None else 1000)
                return self.FisherMatr(evParams, res=res, df=df, spacing=spacing, use_m1m2=use_m1m2, use_chi1chi2=use_chi1chi2,
                                       computeDerivFinDiff=computeDerivFinDiff, return_all=return_all, return_derivatives=return_derivatives,
                                       return_SNR_derivatives=return_SNR_derivatives, **kwargs), snr
            self._prepare_snr(evParams)                       # the dict bookkeeping SNRInteg would have done first
            lambdas, res = self._prepare_fisher(evParams, res, df, computeDerivFinDiff, False, False)
            F, s2, _ = hot_fisher([self], evParams, lambdas, res, spacing, use_m1m2, use_chi1chi2, False, want_snr_integ=True)
            self.last_status = _engine.state().last_status
            return F[0], onp.sqrt(s2[0])
        if self.DutyFactor is not None:
            onp.random.seed(self.seedUse)
        lambdas, res = self._prepare_fisher(evParams, res, df, computeDerivFinDiff, return_derivatives, return_SNR_derivatives)
        n = _num_events(evParams)
        if return_derivatives or return_SNR_derivatives:
            # the reference returns (allFishers, allDerivs) or (allFishers, allSNRDerivs), one entry per arm (signal.py:1091-1098)
            F, (_, sd), _ = hot_fisher([self], evParams, lambdas, res, spacing, use_m1m2, use_chi1chi2, True, want_snr_derivs=True)
            D = hot_strain_derivs([self], evParams, lambdas, res, spacing, use_m1m2, use_chi1chi2) if return_derivatives else None
            masks = onp.array(self._duty_masks(n)) if self.DutyFactor is not None else None
            if masks is not None:
                F = F * masks[:, None, None, :]
                sd = sd * masks[:, :, None]
                if D is not None:
                    D = D * masks[:, None, :, None]
            allF = [F[i] for i in range(F.shape[0])]
            if return_derivatives:
                return allF, [D[i] for i in range(D.shape[0])]
            return allF, [onp.ascontiguousarray(sd[i].T) for i in range(sd.shape[0])]
        per_arm = return_all or (self.DutyFactor is not None)
        F, _, _ = hot_fisher([self], evParams, lambdas, res, spacing, use_m1m2, use_chi1chi2, per_arm)
        self.last_status = _engine.state().last_status
        if self.DutyFactor is not None:
            masks = self._duty_masks(n)
            F = F * onp.array(masks)[:, None, None, :]
        if return_all:
            return [F[i] for i in range(F.shape[0])]
        return F.sum(axis=0) if F.shape[0] > 1 else F[0]

    # ------------------------------------------------------------------ overlap
    def _prepare_overlap(self, WF, evParams):
        """the dict bookkeeping of signal.py:1783-1849 for one of the two waveforms (mutates the dict like the reference);
        returns the arrays the engine consumes."""
        if WF.is_Precessing:
            raise NotImplementedError('precessing waveforms exist in the reference only through the LAL wrapper')
        zeros = onp.zeros_like(evParams['Mc'])
        if 'chi1z' in evParams:
            evParams['chi1x'], evParams['chi1y'], evParams['chi2x'], evParams['chi2y'] = zeros, zeros.copy(), zeros.copy(), zeros.copy()
        else:
            try:
                if self.verbose:
                    print('Adding chi1z, chi2z from chiS, chiA')
                evParams['chi1z'] = evParams['chiS'] + evParams['chiA']
                evParams['chi2z'] = evParams['chiS'] - evParams['chiA']
            except KeyError:
                raise ValueError('Two among chi1z, chi2z and chiS, chiA have to be provided.')
        ev = {k: evParams[k] for k in K.EVENT_KEYS[:11]}
        if WF.is_tidal:
            if 'LambdaTilde' not in evParams:
                try:
                    evParams['LambdaTilde'], evParams['deltaLambda'] = utils.Lamt_delLam_from_Lam12(evParams['Lambda1'], evParams['Lambda2'], evParams['eta'])
                except KeyError:
                    raise ValueError('Two among Lambda1, Lambda2 and LambdaTilde and deltaLambda have to be provided.')
            # GWstrain(is_chi1chi2=True) goes back to Lambda1, Lambda2 from LambdaTilde, deltaLambda (signal.py:554)
            ev['Lambda1'], ev['Lambda2'] = utils.Lam12_from_Lamt_delLam(evParams['LambdaTilde'], evParams['deltaLambda'], evParams['eta'])
        else:
            evParams['LambdaTilde'], evParams['deltaLambda'] = zeros.copy(), zeros.copy()
        if not WF.is_eccentric:
            evParams['ecc'] = zeros.copy()
        else:
            ev['ecc'] = evParams['ecc']
        return ev

    def WFOverlap(self, WF1, WF2, evParams1, evParams2, res=1000, return_separate=False, **kwargs):
        """Overlap of two waveforms in this detector, shape (N,) (or the tuple ``((h1|h2), SNR1, SNR2)`` with
        ``return_separate``); signal.py:1759-1930.  Both strains are evaluated on the grid geomspace(fmin, max(fcut1, fcut2), res)
        by ``gwf_strain`` and integrated by ``gwf_overlap``."""
        utils.check_evparams(evParams1)
        utils.check_evparams(evParams2)
        ev1 = self._prepare_overlap(WF1, evParams1)
        ev2 = self._prepare_overlap(WF2, evParams2)
        fcut1 = WF1.fcut(**evParams1)
        fcut2 = WF2.fcut(**evParams2)
        fcutUse = onp.where(fcut1 > fcut2, fcut1, fcut2)
        if self.fmax is not None:
            fcutUse = onp.where(fcutUse > self.fmax, self.fmax, fcut1)      # as written in the reference, signal.py:1857-1858
        n = _num_events(evParams1)
        fcutUse = onp.broadcast_to(onp.asarray(fcutUse, dtype=float), (n,))
        det = K.gwf_detector(self.det_lat_rad, self.det_long_rad, self.det_xax_rad, 0 if self.detector_shape == 'L' else 1,
                             int(bool(self.useEarthMotion)), int(bool(self.noMotion)), 0, float(self.fmin), 0.)
        ov, s1, s2 = _engine.overlap(WF1._descriptor(evParams1), WF2._descriptor(evParams2), det, self._psd_handle(), ev1, ev2, fcutUse,
                                     self.fmin, n, res)
        overlap_int, SNRh1, SNRh2 = ov.sum(axis=0), onp.sqrt(s1.sum(axis=0)), onp.sqrt(s2.sum(axis=0))
        if return_separate:
            return overlap_int, SNRh1, SNRh2
        return overlap_int / (SNRh1 * SNRh2)
