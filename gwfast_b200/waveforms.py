"""Waveform model classes with the reference's names, constructor arguments and attributes
(gwfast/waveforms.py: WaveFormModel :49-199, TaylorF2_RestrictedPN :697, IMRPhenomD :959,
IMRPhenomD_NRTidalv2 :1339, IMRPhenomHM :1838).

The classes carry no arithmetic of their own for the O(N*res) quantities: ``Phi``, ``Ampl``, ``tau_star`` (and
``hphc``) evaluate on the GPU through ``gwf_waveform``; ``fcut`` and the parameter bookkeeping are per-event
host work exactly as in the reference.  ``_descriptor()`` is what ``GWSignal``/``DetNet`` hand to the engine.
"""
from abc import ABC, abstractmethod

import numpy as np

from . import gwfastGlobals as glob
from . import _capi as K


class WaveFormModel(ABC):
    """Parameter-ordering contract of the reference base class (waveforms.py:68-147)."""

    def __init__(self, objType, fcutPar, is_newtonian=False, is_tidal=False, is_HigherModes=False, is_chi1chi2=True,
                 is_Precessing=False, is_LAL=False, is_prec_ang=False, is_eccentric=False, is_holomorphic=False, apply_fcut=True):
        if is_Precessing or is_LAL:
            raise NotImplementedError('gwfast_b200 builds the non-precessing native models only '
                                      '(NewtInspiral, TaylorF2_RestrictedPN, IMRPhenomD, IMRPhenomD_NRTidalv2, IMRPhenomHM, IMRPhenomNSBH)')
        self.objType = objType
        self.fcutPar = fcutPar
        self.is_newtonian = is_newtonian
        self.is_tidal = is_tidal
        self.is_HigherModes = is_HigherModes
        self.is_chi1chi2 = is_chi1chi2
        self.is_Precessing = is_Precessing
        self.is_LAL = is_LAL
        self.is_eccentric = is_eccentric
        self.is_holomorphic = is_holomorphic
        self.apply_fcut = apply_fcut
        names = ['Mc', 'eta', 'dL', 'theta', 'phi', 'iota', 'psi', 'tcoal', 'Phicoal']
        names += ['chi1z', 'chi2z'] if is_chi1chi2 else ['chiS', 'chiA']
        if is_tidal:
            # the Fisher is computed for LambdaTilde and deltaLambda although the waveforms take Lambda1, Lambda2
            names += ['LambdaTilde', 'deltaLambda']
        if is_eccentric:
            names += ['ecc']                    # waveforms.py:113-124
        # rows of the engine's 11/13-parameter layout that this model's Fisher keeps (all of them except for NewtInspiral)
        self._engine_rows = None
        if is_newtonian:
            if is_chi1chi2:
                # the reference renames ParNums['chiS'] -> 'chi1z' after replacing the dict by the 8-parameter one and fails
                # (waveforms.py:96, 130-131): NewtInspiral has to be built with is_chi1chi2=False there, and so here
                raise KeyError('chiS')
            # mass ratio and spins do not enter: 8 parameters (waveforms.py:94-96)
            self._engine_rows = [0, 2, 3, 4, 5, 6, 7, 8]
            names = ['Mc', 'dL', 'theta', 'phi', 'iota', 'psi', 'tcoal', 'Phicoal']
        self.ParNums = {k: i for i, k in enumerate(names)}
        self.nParams = len(names)

    # ---- engine descriptor
    _model_id = None

    def _flags(self):
        return 0 if self.apply_fcut else K.GWF_MODEL_NO_FCUT

    def _descriptor(self, evParams=None):
        fl = self._flags()
        fref = getattr(self, 'fRef', None)
        if fref is not None:
            fl |= K.GWF_MODEL_HAS_FREF
        if evParams is not None and 'Lambda1' in evParams:
            fl |= K.GWF_MODEL_LAMBDA_GIVEN
        return K.gwf_model(self._model_id, fl, float(self.fcutPar), float(fref) if fref is not None else 0.)

    # ---- elementwise API (GPU)
    def _waveform(self, f, what, **kwargs):
        from . import _elementwise
        return _elementwise.evaluate(self, f, what, kwargs)

    def Phi(self, f, **kwargs):
        """GW phase on the grid ``f`` (``(res,)`` or ``(res, N)``), as WaveFormModel.Phi (waveforms.py:149)."""
        return self._waveform(f, 'phi', **kwargs)

    def Ampl(self, f, **kwargs):
        """GW amplitude on the grid ``f``, as WaveFormModel.Ampl (waveforms.py:164)."""
        return self._waveform(f, 'ampl', **kwargs)

    def tau_star(self, f, **kwargs):
        """Time to coalescence in seconds (3.5PN, arXiv:0907.0700 eq. 3.8b), as in waveforms.py:878."""
        return self._waveform(f, 'tau', **kwargs)

    @abstractmethod
    def fcut(self, **kwargs):
        pass


class NewtInspiral(WaveFormModel):
    """Leading-order inspiral, waveforms.py:205-260.  Runs on the TaylorF2 kernels with every PN coefficient but the first
    switched off (GWF_MODEL_NEWTONIAN) and the base class's numerical ``tau_star``; 8 Fisher parameters."""
    _model_id = K.GWF_TAYLORF2

    def __init__(self, **kwargs):
        super().__init__('BBH', 1. / (6. * np.pi * np.sqrt(6.) * glob.GMsun_over_c3), is_newtonian=True, is_holomorphic=True, **kwargs)

    def _flags(self):
        return K.GWF_MODEL_NEWTONIAN

    def fcut(self, **kwargs):
        """WaveFormModel.fcut, waveforms.py:188-199."""
        return self.fcutPar / (kwargs['Mc'] / (kwargs['eta'] ** (3. / 5.)))


class TaylorF2_RestrictedPN(WaveFormModel):
    """waveforms.py:697-953, including the low-eccentricity extension (:814-845; parameter ``ecc``, ``fRef_ecc``)."""
    _model_id = K.GWF_TAYLORF2

    def __init__(self, fHigh=None, is_tidal=False, use_3p5PN_SpinHO=False, phiref_vlso=False, is_eccentric=False, fRef_ecc=None,
                 which_ISCO='Schw', use_QuadMonTid=False, **kwargs):
        if fHigh is None:
            fHigh = 1. / (6. * np.pi * np.sqrt(6.) * glob.GMsun_over_c3)   # Hz
        if which_ISCO not in ('Schw', 'Kerr'):
            raise ValueError("which_ISCO has to be 'Schw' or 'Kerr'")
        self.use_3p5PN_SpinHO = use_3p5PN_SpinHO
        self.phiref_vlso = phiref_vlso
        self.fRef_ecc = fRef_ecc
        self.which_ISCO = which_ISCO
        self.use_QuadMonTid = use_QuadMonTid
        super().__init__('BNS' if is_tidal else 'BBH', fHigh, is_tidal=is_tidal, is_eccentric=is_eccentric, is_holomorphic=True, **kwargs)

    def _flags(self):
        fl = 0
        if self.is_tidal:
            fl |= K.GWF_MODEL_TIDAL
        if self.use_3p5PN_SpinHO:
            fl |= K.GWF_MODEL_3P5PN_SPINHO
        if self.phiref_vlso:
            fl |= K.GWF_MODEL_PHIREF_VLSO
        if self.use_QuadMonTid:
            fl |= K.GWF_MODEL_QUADMON_TID
        if self.which_ISCO == 'Kerr':
            fl |= K.GWF_MODEL_KERR_ISCO
        if self.is_eccentric:
            fl |= K.GWF_MODEL_ECCENTRIC
        return fl

    def _descriptor(self, evParams=None):
        d = super()._descriptor(evParams)
        if self.is_eccentric and self.fRef_ecc is not None:
            # the eccentric reference frequency travels in gwf_model.fRef (TaylorF2 has no phase reference frequency of its own)
            d.flags |= K.GWF_MODEL_HAS_FREF
            d.fRef = float(self.fRef_ecc)
        return d

    def fcut(self, **kwargs):
        """waveforms.py:903-953."""
        if self.which_ISCO == 'Schw':
            return self.fcutPar / (kwargs['Mc'] / (kwargs['eta'] ** (3. / 5.)))
        return self._waveform(None, 'fcut', **kwargs)


class IMRPhenomD(WaveFormModel):
    """waveforms.py:959-1333."""
    _model_id = K.GWF_IMRPHENOMD

    def __init__(self, fRef=None, **kwargs):
        self.AMP_fJoin_INS = 0.014
        self.PHI_fJoin_INS = 0.018
        self.fRef = fRef
        super().__init__('BBH', 0.2, **kwargs)
        self.QNMgrid_a, self.QNMgrid_fring, self.QNMgrid_fdamp = _qnm_tables()

    def fcut(self, **kwargs):
        """waveforms.py:1324-1333."""
        return self.fcutPar / (kwargs['Mc'] * glob.GMsun_over_c3 / (kwargs['eta'] ** (3. / 5.)))


class IMRPhenomD_NRTidalv2(WaveFormModel):
    """waveforms.py:1339-1832."""
    _model_id = K.GWF_IMRPHENOMD_NRTIDALV2

    def __init__(self, fRef=None, **kwargs):
        self.AMP_fJoin_INS = 0.014
        self.PHI_fJoin_INS = 0.018
        self.fRef = fRef
        super().__init__('BNS', 0.2, is_tidal=True, **kwargs)
        self.QNMgrid_a, self.QNMgrid_fring, self.QNMgrid_fdamp = _qnm_tables()

    def fcut(self, **kwargs):
        """1.2 f_merger(kappa2T, q)/(M GMsun/c^3); waveforms.py:1794-1832 (Lambda -> 0 if the keys are absent)."""
        eta = kwargs['eta']
        M = kwargs['Mc'] / (eta ** (3. / 5.))
        Seta = np.sqrt(np.where(eta < 0.25, 1.0 - 4.0 * eta, 0.))
        q = 0.5 * (1.0 + Seta - 2.0 * eta) / eta
        if 'Lambda1' in kwargs:
            Lambda1, Lambda2 = kwargs['Lambda1'], kwargs['Lambda2']
        else:
            Lambda1, Lambda2 = np.zeros(M.shape), np.zeros(M.shape)
        Xa, Xb = 0.5 * (1.0 + Seta), 0.5 * (1.0 - Seta)
        kappa2T = (3.0 / 13.0) * ((1.0 + 12.0 * Xb / Xa) * (Xa ** 5) * Lambda1 + (1.0 + 12.0 * Xa / Xb) * (Xb ** 5) * Lambda2)
        numPT = 1.0 + 3.35411203e-2 * kappa2T + 4.31460284e-5 * kappa2T * kappa2T
        denPT = 1.0 + 7.54224145e-2 * kappa2T + 2.23626859e-4 * kappa2T * kappa2T
        f_merger = (0.3586 / np.sqrt(q)) * (numPT / denPT) / (2. * np.pi)
        return 1.2 * f_merger / (M * glob.GMsun_over_c3)


class IMRPhenomHM(WaveFormModel):
    """waveforms.py:1838-2749."""
    _model_id = K.GWF_IMRPHENOMHM

    def __init__(self, **kwargs):
        self.AMP_fJoin_INS = 0.014
        self.PHI_fJoin_INS = 0.018
        super().__init__('BBH', 0.2, is_HigherModes=True, **kwargs)

    def fcut(self, **kwargs):
        """waveforms.py:2737-2749."""
        return self.fcutPar / (kwargs['Mc'] * glob.GMsun_over_c3 / (kwargs['eta'] ** (3. / 5.)))

    def hphc(self, f, **kwargs):
        return self._waveform(f, 'hphc', **kwargs)


class IMRPhenomNSBH(WaveFormModel):
    """waveforms.py:2752-3374.  Index 1 is the black hole, 2 the neutron star; ``Lambda1`` is discarded as in the reference.  The
    xi_tide table the reference reads from ``WFfiles/xiTide_Table_200.h5`` (or tabulates with numpy.roots, :3286-3343) is computed
    on the device the first time the model runs there (csrc/model_nsbh.cuh)."""
    _model_id = K.GWF_IMRPHENOMNSBH

    def __init__(self, fRef=None, verbose=True, **kwargs):
        self.PHI_fJoin_INS = 0.018
        self.verbose = verbose
        self.fRef = fRef
        super().__init__('NSBH', 0.2, is_tidal=True, **kwargs)
        self.QNMgrid_a, self.QNMgrid_fring, self.QNMgrid_fdamp = _qnm_tables()

    def fcut(self, **kwargs):
        """waveforms.py:3273-3284."""
        return self.fcutPar / (kwargs['Mc'] * glob.GMsun_over_c3 / (kwargs['eta'] ** (3. / 5.)))


_QNM = None


def _qnm_tables():
    """QNM ringdown tables, waveforms.py:988-990 (also uploaded once to the device by the engine)."""
    global _QNM
    if _QNM is None:
        import os
        _QNM = tuple(np.loadtxt(os.path.join(glob.WFfilesPath, 'QNMData_%s.txt' % k)) for k in ('a', 'fring', 'fdamp'))
    return _QNM
