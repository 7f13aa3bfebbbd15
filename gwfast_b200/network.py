"""``DetNet`` with the reference's interface (gwfast/network.py:12-152).

Where the reference loops over detectors in Python (network.py:67, 103), the whole network goes to the GPU as ONE
fused launch: the waveform of an event is evaluated once per frequency sample and projected on every detector and
arm, and (without ``return_all``) all arms accumulate into a single packed Fisher matrix per event.  Detectors that
use different waveform objects fall back to one launch per detector (still on the GPU).
"""
import numpy as onp

from . import gwfastUtils as utils
from . import signal as _sig


class DetNet(object):
    def __init__(self, signals, verbose=True):
        # signals: {'detector_name': GWSignal}
        self.signals = signals
        self.verbose = verbose
        self.last_status = None

    def _clear_cache(self):
        for d in self.signals.keys():
            self.signals[d]._clear_cache()

    def _update_all_seeds(self, seeds=[], verbose=True):
        """network.py:39-51."""
        if seeds == []:
            seeds = [None for _ in range(len(list(self.signals.keys())))]
        for i, d in enumerate(list(self.signals.keys())):
            self.signals[d]._update_seed(seed=seeds[i])
            if verbose:
                print('\nSeed for detector %s is %s' % (d, self.signals[d].seedUse))

    def _fusable(self):
        """one fused launch needs a common waveform object and a network inside the C side's limits (8 detectors, 16 arms,
        4 distinct (fmin, fmax) pairs); anything else runs one launch per detector"""
        sigs = list(self.signals.values())
        groups = {(float(s.fmin), None if s.fmax is None else float(s.fmax)) for s in sigs}
        arms = sum(s._narms() for s in sigs)
        return all(s.wf_model is sigs[0].wf_model for s in sigs) and len(sigs) <= 8 and arms <= 16 and len(groups) <= 4

    # ------------------------------------------------------------------ SNR
    def SNR(self, evParams, res=1000, return_all=False):
        """Network SNR, shape (N,); with ``return_all`` a dict of per-detector (per-arm for triangles) SNRs plus 'net'."""
        utils.check_evparams(evParams)
        names = list(self.signals.keys())
        sigs = [self.signals[d] for d in names]
        snrs = {}
        if self._fusable():
            for s in sigs:
                s._prepare_snr(evParams)
            n = _sig._num_events(evParams)
            s2, slices, _ = _sig.hot_snr(sigs, evParams, res)
            for d, s, sl in zip(names, sigs, slices):
                if s.DutyFactor is not None:
                    onp.random.seed(s.seedUse)
                snr_ = s._snr_from_arms(s2[sl], n, return_all)
                if s.detector_shape == 'T' and return_all:
                    for i in range(3):
                        snrs[d + '_%s' % i] = snr_[i]
                else:
                    snrs[d] = snr_
        else:
            for d, s in zip(names, sigs):
                snr_ = s.SNRInteg(evParams, res=res, return_all=return_all)
                if s.detector_shape == 'T' and return_all:
                    for i in range(3):
                        snrs[d + '_%s' % i] = snr_[i]
                else:
                    snrs[d] = snr_
        net_snr = onp.sqrt(onp.array([snrs[k] ** 2 for k in snrs.keys()]).sum(axis=0))
        if return_all:
            snrs['net'] = net_snr
            return snrs
        return net_snr

    # ------------------------------------------------------------------ Fisher
    def FisherMatr(self, evParams, return_all=False, return_derivatives=False, return_SNR_derivatives=False, return_SNR=False, **kwargs):
        """Total Fisher matrix, shape (nParams, nParams, N); with ``return_all`` a dict per detector/arm plus 'net'.

        ``return_SNR=True`` (an addition to the reference's keywords): returns ``(F, SNR)`` where ``SNR`` is what ``self.SNR`` returns
        for the same events and ``res`` -- taken from the same fused launch when the network is fusable and no per-arm output or duty
        factor is involved (the Fisher kernel integrates |h|^2/Sn anyway), otherwise from a separate ``SNR`` call.  After the call
        ``self.last_status`` holds the per-event status words (``_capi.GWF_EV_*``; 0 = clean)."""
        utils.check_evparams(evParams)
        names = list(self.signals.keys())
        sigs = [self.signals[d] for d in names]
        if return_SNR:
            fused = self._fusable() and not (return_all or return_derivatives or return_SNR_derivatives) and \
                all(s.DutyFactor is None for s in sigs) and not (kwargs.get('df') is not None and kwargs.get('res', 1000) is None)
            if not fused:
                snr = self.SNR(evParams, res=kwargs.get('res', 1000) or 1000, return_all=return_all)
                return self.FisherMatr(evParams, return_all=return_all, return_derivatives=return_derivatives,
                                       return_SNR_derivatives=return_SNR_derivatives, **kwargs), snr
            for s in sigs:
                s._prepare_snr(evParams)                      # the dict bookkeeping DetNet.SNR would have done first
            kw = dict(kwargs)
            res = kw.pop('res', 1000)
            lambdas = None
            for s in sigs:
                lambdas, res_s = s._prepare_fisher(evParams, res, kw.get('df'), kw.get('computeDerivFinDiff', False), False, False)
            F, s2, _ = _sig.hot_fisher(sigs, evParams, lambdas, res_s, kw.get('spacing', 'geom'), kw.get('use_m1m2', False),
                                       kw.get('use_chi1chi2', True), False, want_snr_integ=True)
            self.last_status = _sig._engine.state().last_status
            return F[0], onp.sqrt(s2[0])
        if return_derivatives or return_SNR_derivatives:
            # network.py:124-152: per-detector calls, every arm separately; 'net' of the SNR derivatives = sum over arms / network SNR
            SNRs_net = kwargs.pop('SNRs', None)
            allF, allDer = {}, {}
            for d, s in zip(names, sigs):
                if self.verbose:
                    print('Computing Fisher for %s...' % d)
                F_, D_ = s.FisherMatr(evParams, return_all=True, return_derivatives=return_derivatives,
                                      return_SNR_derivatives=return_SNR_derivatives, **kwargs)
                if s.detector_shape == 'T':
                    for i in range(3):
                        allF[d + '_%s' % i] = F_[i]
                        allDer[d + '_%s' % i] = D_[i]
                else:
                    allF[d] = F_[0]
                    allDer[d] = D_[0]
            if return_SNR_derivatives and not return_derivatives:
                if SNRs_net is None:
                    SNRs_net = self.SNR(evParams, return_all=False)
                allDer['net'] = onp.array([allDer[k] for k in allDer.keys()]).sum(axis=0) / SNRs_net
            if self.verbose:
                print('Done.')
            allF['net'] = onp.array([allF[k] for k in allF.keys()]).sum(axis=0)
            return allF, allDer
        if not self._fusable():
            allF = {}
            for d, s in zip(names, sigs):
                if self.verbose:
                    print('Computing Fisher for %s...' % d)
                F_ = s.FisherMatr(evParams, return_all=return_all, **kwargs)
                if s.detector_shape == 'T' and return_all:
                    for i in range(3):
                        allF[d + '_%s' % i] = F_[i]
                elif return_all:
                    allF[d] = F_[0]
                else:
                    allF[d] = F_
            if self.verbose:
                print('Done.')
            totF = onp.array([allF[k] for k in allF.keys()]).sum(axis=0)
            if return_all:
                allF['net'] = totF
                return allF
            return totF
        res = kwargs.pop('res', 1000)
        df = kwargs.pop('df', None)
        spacing = kwargs.pop('spacing', 'geom')
        use_m1m2 = kwargs.pop('use_m1m2', False)
        use_chi1chi2 = kwargs.pop('use_chi1chi2', True)
        findiff = kwargs.pop('computeDerivFinDiff', False)
        kwargs.pop('computeAnalyticalDeriv', None)
        kwargs.pop('use_prec_ang', None)
        lambdas = None
        for s in sigs:
            lambdas, res_s = s._prepare_fisher(evParams, res, df, findiff, False, False)
        if df is not None and res is None:
            # each detector derives its own res from df (signal.py:890-892): keep per-detector launches
            self_fusable = False
        else:
            self_fusable = True
        duty = any(s.DutyFactor is not None for s in sigs)
        if not self_fusable:
            return self._per_detector(evParams, names, sigs, return_all, dict(res=res, df=df, spacing=spacing, use_m1m2=use_m1m2, use_chi1chi2=use_chi1chi2))
        if self.verbose:
            for d in names:
                print('Computing Fisher for %s...' % d)
        per_arm = return_all or duty
        F, _, _ = _sig.hot_fisher(sigs, evParams, lambdas, res_s, spacing, use_m1m2, use_chi1chi2, per_arm)
        self.last_status = _sig._engine.state().last_status
        if self.verbose:
            print('Done.')
        if not per_arm:
            return F[0]
        n = _sig._num_events(evParams)
        allF, a0 = {}, 0
        for d, s in zip(names, sigs):
            na = s._narms()
            Fd = F[a0:a0 + na]
            a0 += na
            if s.DutyFactor is not None:
                onp.random.seed(s.seedUse)
                Fd = Fd * onp.array(s._duty_masks(n))[:, None, None, :]
            if return_all and s.detector_shape == 'T':
                for i in range(3):
                    allF[d + '_%s' % i] = Fd[i]
            elif return_all:
                allF[d] = Fd[0]
            else:
                allF[d] = Fd.sum(axis=0)
        totF = onp.array([allF[k] for k in allF.keys()]).sum(axis=0)
        if return_all:
            allF['net'] = totF
            return allF
        return totF

    def _per_detector(self, evParams, names, sigs, return_all, kw):
        allF = {}
        for d, s in zip(names, sigs):
            if self.verbose:
                print('Computing Fisher for %s...' % d)
            F_ = s.FisherMatr(evParams, return_all=return_all, **kw)
            if s.detector_shape == 'T' and return_all:
                for i in range(3):
                    allF[d + '_%s' % i] = F_[i]
            elif return_all:
                allF[d] = F_[0]
            else:
                allF[d] = F_
        if self.verbose:
            print('Done.')
        totF = onp.array([allF[k] for k in allF.keys()]).sum(axis=0)
        if return_all:
            allF['net'] = totF
            return allF
        return totF

    def WFOverlap(self, WF1, WF2, evParams1, evParams2, res=1000, **kwargs):
        """Overlap of two waveforms in the network, shape (N,); network.py:198-225."""
        utils.check_evparams(evParams1)
        utils.check_evparams(evParams2)
        overlap_all = onp.zeros_like(evParams1['Mc'])
        SNR1_all = onp.zeros_like(evParams1['Mc'])
        SNR2_all = onp.zeros_like(evParams1['Mc'])
        for d in self.signals.keys():
            overlap_int, SNR1, SNR2 = self.signals[d].WFOverlap(WF1, WF2, evParams1, evParams2, res=res, return_separate=True)
            overlap_all += overlap_int
            SNR1_all += SNR1 ** 2
            SNR2_all += SNR2 ** 2
        return overlap_all / onp.sqrt(SNR1_all * SNR2_all)
