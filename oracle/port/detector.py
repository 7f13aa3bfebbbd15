"""numpy restatement of the reference's single-detector / network layer (TEST INFRASTRUCTURE).

Follows gwfast/signal.py (pattern functions :342-387, Earth-centre delay :401-423, amplitudes
:425-466, strain :485-655, SNR :658-777, Fisher :782-1098) and gwfast/network.py:53-123.
Derivatives of the strain are taken with forward-mode duals w.r.t. EVERY Fisher parameter
(the reference's ``computeAnalyticalDeriv=False`` formulation, signal.py:1161-1165, which the
reference states -- and SURVEY.md 8(c) verified -- agrees with its analytic+AD default), so this
checker shares no derivative formula with the CUDA path it checks.
"""
import numpy as np

from ..dual import Dual, seed
from . import waveforms as W
from .constants import R_EARTH_KM, C_KM_S, DAY_S

TWO_PI = 2. * np.pi


class Detector:
    def __init__(self, wf_model, psd_path=None, detector_shape='T', det_lat=40.44, det_long=9.45, det_xax=0.,
                 is_ASD=True, useEarthMotion=False, noMotion=False, fmin=2., fmax=None, compute2arms=True,
                 psd=None, **_ignored):
        if detector_shape not in ('L', 'T'):
            raise ValueError('Enter valid detector configuration')
        self.wf_model = wf_model
        self.detector_shape = detector_shape
        self.lat, self.long, self.xax = np.deg2rad(det_lat), np.deg2rad(det_long), np.deg2rad(det_xax)
        if psd is None:
            tab = np.loadtxt(psd_path, usecols=(0, 1))               # signal.py:122
            psd = (tab[:, 0], tab[:, 1] ** 2 if is_ASD else tab[:, 1])
        self.strainFreq, self.noiseCurve = psd
        self.useEarthMotion = useEarthMotion and not noMotion         # signal.py:138-140
        self.noMotion = noMotion
        self.fmin, self.fmax = fmin, fmax
        self.angbtwArms = 0.5 * np.pi if detector_shape == 'L' else np.pi / 3.
        self.compute2arms = compute2arms

    # ------------------------------------------------------------------ geometry
    def _ab(self, ra, dec, t, rot):
        """JKS a(t), b(t); signal.py:360-376."""
        x2 = 2. * (self.xax + rot)
        lat = self.lat
        ang = ra - self.long - TWO_PI * t
        a = (0.0625 * np.sin(x2) * (3. - np.cos(2. * lat)) * (3. - np.cos(2. * dec)) * np.cos(2. * ang)
             - 0.25 * np.cos(x2) * np.sin(lat) * (3. - np.cos(2. * dec)) * np.sin(2. * ang)
             + 0.25 * np.sin(x2) * np.sin(2. * lat) * np.sin(2. * dec) * np.cos(ang)
             - 0.5 * np.cos(x2) * np.cos(lat) * np.sin(2. * dec) * np.sin(ang)
             + 3. * 0.25 * np.sin(x2) * (np.cos(lat) * np.cos(dec)) ** 2.)
        b = (np.cos(x2) * np.sin(lat) * np.sin(dec) * np.cos(2. * ang)
             + 0.25 * np.sin(x2) * (3. - np.cos(2. * lat)) * np.sin(dec) * np.sin(2. * ang)
             + np.cos(x2) * np.cos(lat) * np.cos(dec) * np.cos(ang)
             + 0.5 * np.sin(x2) * np.sin(2. * lat) * np.cos(dec) * np.sin(ang))
        return a, b

    def pattern(self, theta, phi, t, psi, rot=0.):
        """signal.py:342-387 (rot in degrees)."""
        a, b = self._ab(phi, 0.5 * np.pi - theta, t, np.deg2rad(rot))
        s = np.sin(self.angbtwArms)
        return s * (a * np.cos(2. * psi) + b * np.sin(2 * psi)), s * (b * np.cos(2. * psi) - a * np.sin(2 * psi))

    def delt_loc(self, theta, phi, t):
        """signal.py:401-423; seconds."""
        ra, dec = phi, 0.5 * np.pi - theta
        n = (np.cos(dec) * np.cos(ra) * np.cos(self.lat) * np.cos(self.long + TWO_PI * t)
             + np.cos(dec) * np.sin(ra) * np.cos(self.lat) * np.sin(self.long + TWO_PI * t)
             + np.sin(dec) * np.sin(self.lat))
        return -R_EARTH_KM * n / C_KM_S

    def _times(self, ev, f):
        """detector time (days) and Earth-centre delay (s); signal.py:444-453, 564-580."""
        if self.noMotion:
            t = 0.
        elif self.useEarthMotion:
            t = ev['tcoal'] - self.wf_model.tau_star(f, **ev) / DAY_S
        else:
            t = ev['tcoal']
        dloc = self.delt_loc(ev['theta'], ev['phi'], t)
        return t + dloc / DAY_S, dloc

    # ------------------------------------------------------------------ signal
    def amplitudes(self, ev, f, rot=0.):
        """signal.py:425-466."""
        t, _ = self._times(ev, f)
        Fp, Fc = self.pattern(ev['theta'], ev['phi'], t, ev['psi'], rot)
        if self.wf_model.is_HigherModes:
            hp, hc = self.wf_model.hphc(f, **ev)
            return abs(hp) * Fp, abs(hc) * Fc
        A = self.wf_model.Ampl(f, **ev)
        ci = np.cos(ev['iota'])
        return A * Fp * 0.5 * (1. + ci ** 2), A * Fc * ci

    def strain(self, f, ev, rot=0.):
        """complex strain for an already-converted event dict; signal.py:564-654."""
        t, dloc = self._times(ev, f)
        extra = TWO_PI * f * dloc + TWO_PI * f * (ev['tcoal'] * DAY_S) - ev['Phicoal']
        Fp, Fc = self.pattern(ev['theta'], ev['phi'], t, ev['psi'], rot)
        if self.wf_model.is_HigherModes:
            hp, hc = self.wf_model.hphc(f, **ev)
            return (hp * Fp + hc * Fc) * np.exp(1j * extra)
        A = self.wf_model.Ampl(f, **ev)
        ci = np.cos(ev['iota'])
        Ap, Ac = A * Fp * 0.5 * (1. + ci ** 2), A * Fc * ci
        return (Ap + 1j * Ac) * np.exp(1j * (extra - self.wf_model.Phi(f, **ev)))

    # ------------------------------------------------------------------ grids
    def _fill(self, ev):
        """key filling done by SNRInteg / FisherMatr; signal.py:694-713 (mutates ev, like the reference)."""
        if 'chi1z' not in ev:
            try:
                ev['chi1z'] = ev['chiS'] + ev['chiA']
                ev['chi2z'] = ev['chiS'] - ev['chiA']
            except KeyError:
                raise ValueError('Two among chi1z, chi2z and chiS, chiA have to be provided.')

    def grids(self, ev, res=1000, spacing='geom'):
        """signal.py:715-723 / 884-901."""
        fcut = self.wf_model.fcut(**ev)
        if self.fmax is not None:
            fcut = np.where(fcut > self.fmax, self.fmax, fcut)
        fmin = np.full(fcut.shape, self.fmin)
        fg = np.geomspace(fmin, fcut, num=int(res)) if spacing == 'geom' else np.linspace(fmin, fcut, num=int(res))
        return fg, np.interp(fg, self.strainFreq, self.noiseCurve, left=1., right=1.)

    def _rots(self):
        if self.detector_shape == 'L':
            return [0.]
        return [0., 60.] if self.compute2arms else [0., 60., 120.]

    def snr2_arms(self, ev, res=1000):
        """per-arm SNR^2/4 integrals; signal.py:725-767."""
        self._fill(ev)
        if self.wf_model.is_tidal and 'Lambda1' not in ev:
            try:
                ev['Lambda1'], ev['Lambda2'] = W.lam12_from_lamt_dellam(ev['LambdaTilde'], ev['deltaLambda'], ev['eta'])
            except KeyError:
                raise ValueError('Two among Lambda1, Lambda2 and LambdaTilde and deltaLambda have to be provided.')
        fg, Sn = self.grids(ev, res)
        amps = [self.amplitudes(ev, fg, rot) for rot in self._rots()]
        if self.detector_shape == 'T' and self.compute2arms:
            amps.append((-(amps[0][0] + amps[1][0]), -(amps[0][1] + amps[1][1])))
        return np.array([np.trapezoid((Ap * Ap + Ac * Ac) / Sn, fg, axis=0) for Ap, Ac in amps])

    def SNRInteg(self, ev, res=1000, return_all=False):
        s2 = self.snr2_arms(ev, res)
        if self.detector_shape == 'T':
            return 2 * np.sqrt(s2) if return_all else 2 * np.sqrt(s2.sum(axis=0))
        return 2 * np.sqrt(s2[0])

    # ------------------------------------------------------------------ Fisher
    def _fisher_args(self, ev, use_m1m2, use_chi1chi2):
        """the differentiated argument list in ParNums order; signal.py:815-882."""
        Mc, eta = ev['Mc'], ev['eta']
        if use_m1m2:
            Mc, eta = W.m1m2_from_mceta(Mc, eta)
        c1, c2 = ev['chi1z'], ev['chi2z']
        if not use_chi1chi2:
            c1, c2 = 0.5 * (c1 + c2), 0.5 * (c1 - c2)
        args = [Mc, eta, ev['dL'], ev['theta'], ev['phi'], ev['iota'], ev['psi'], ev['tcoal'], ev['Phicoal'], c1, c2]
        if self.wf_model.is_tidal:
            if 'Lambda1' in ev:
                L1, L2 = ev['Lambda1'], ev['Lambda2']
            else:
                try:
                    L1, L2 = W.lam12_from_lamt_dellam(ev['LambdaTilde'], ev['deltaLambda'], ev['eta'])
                except KeyError:
                    raise ValueError('Two among Lambda1, Lambda2 and LambdaTilde and deltaLambda have to be provided.')
            args += list(W.lamt_dellam_from_lam12(L1, L2, ev['eta']))
        return [np.asarray(a, dtype=float) for a in args]

    def _strain_of_args(self, f, args, rot, use_m1m2, use_chi1chi2):
        """GWstrain's parameter re-mapping; signal.py:522-559."""
        Mc, eta = (W.mceta_from_m1m2(args[0], args[1]) if use_m1m2 else (args[0], args[1]))
        c1, c2 = (args[9], args[10]) if use_chi1chi2 else (args[9] + args[10], args[9] - args[10])
        ev = dict(Mc=Mc, eta=eta, dL=args[2], theta=args[3], phi=args[4], iota=args[5], psi=args[6], tcoal=args[7],
                  Phicoal=args[8], chi1z=c1, chi2z=c2)
        if self.wf_model.is_tidal:
            ev['Lambda1'], ev['Lambda2'] = W.lam12_from_lamt_dellam(args[11], args[12], eta)
        return self.strain(f, ev, rot)

    def derivatives(self, ev, res=1000, spacing='geom', use_m1m2=False, use_chi1chi2=True):
        """per evaluated arm: (nP, res, N) complex derivative rows, tcoal row already in 1/s; signal.py:917-920."""
        self._fill(ev)
        args = self._fisher_args(ev, use_m1m2, use_chi1chi2)
        fg, Sn = self.grids(ev, res, spacing)
        nP = len(args)
        tc = self.wf_model.ParNums['tcoal']
        out = []
        for rot in self._rots():
            h = self._strain_of_args(fg, seed(args, tuple(range(nP))), rot, use_m1m2, use_chi1chi2)
            D = np.moveaxis(h.d, -1, 0).copy()
            D[tc] /= DAY_S
            out.append(D)
        if self.detector_shape == 'T' and self.compute2arms:
            out.append(-(out[0] + out[1]))                            # signal.py:1057
        return out, fg, Sn

    def FisherMatr(self, ev, res=1000, spacing='geom', use_m1m2=False, use_chi1chi2=True, return_all=False, **_ignored):
        """signal.py:922-931, 1087-1093: F_ab = 4 trapz(Re(conj(D_a) D_b)/Sn)."""
        Ds, fg, Sn = self.derivatives(ev, res, spacing, use_m1m2, use_chi1chi2)
        Fs = []
        for D in Ds:
            nP = D.shape[0]
            F = np.zeros((nP, nP, D.shape[2]))
            for a in range(nP):
                for b in range(a, nP):
                    F[a, b] = F[b, a] = 4. * np.trapezoid((np.conj(D[a]) * D[b]).real / Sn, fg, axis=0)
            Fs.append(F)
        if return_all:
            return Fs
        return np.array(Fs).sum(axis=0) if self.detector_shape == 'T' else Fs[0]


class Network:
    """gwfast/network.py:53-123."""

    def __init__(self, signals):
        self.signals = signals

    def SNR(self, ev, res=1000, return_all=False):
        out = {}
        for d, s in self.signals.items():
            r = s.SNRInteg(ev, res=res, return_all=return_all)
            if s.detector_shape == 'T' and return_all:
                for i in range(3):
                    out['%s_%d' % (d, i)] = r[i]
            else:
                out[d] = r
        net = np.sqrt(np.array([v ** 2 for v in out.values()]).sum(axis=0))
        if return_all:
            out['net'] = net
            return out
        return net

    def FisherMatr(self, ev, return_all=False, **kw):
        out = {}
        for d, s in self.signals.items():
            F = s.FisherMatr(ev, return_all=return_all, **kw)
            if return_all and s.detector_shape == 'T':
                for i in range(3):
                    out['%s_%d' % (d, i)] = F[i]
            elif return_all:
                out[d] = F[0]
            else:
                out[d] = F
        tot = np.array(list(out.values())).sum(axis=0)
        if return_all:
            out['net'] = tot
            return out
        return tot
