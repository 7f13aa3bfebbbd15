"""return_SNR_derivatives / return_derivatives against the reference's own outputs
(tests/golden/deriv_*.npz, oracle/make_golden_derivs.py: DetNet.FisherMatr(return_SNR_derivatives=True / return_derivatives=True)
of the unmodified reference).

Tolerances (written here):
  * SNR derivatives (h | d_i h) per arm: |delta| <= 1e-6 * sqrt(F_ii) * SNR_arm -- the Fisher tolerance of the north star applied
    to the Cauchy-Schwarz scale of this inner product; the 'net' entry (sum over arms / network SNR) likewise.
  * derivative strain d h / d p_i: |delta| <= 2e-6 * max_f |d_i h| per (arm, parameter, event).  The total phase
    2 pi f tcoal 86400 reaches ~1e9 rad, where one float64 ulp is 1e-7 rad: reference and engine both evaluate exp(i Psi) from
    a float64 Psi, so agreement beyond ~1e-7 is not defined.
  * IMRPhenomD_NRTidalv2: the last grid sample is excluded (0 or 1 by last-bit rounding in the reference)."""
import shutil

import numpy as np
import pytest

from conftest import load_golden, make_network, copy_events

SD_TOL = 1e-6
D_TOL = 2e-6


def _arms(out):
    return [k[len('snrderiv__'):] for k in out if k.startswith('snrderiv__') and k != 'snrderiv__net']


def _sd_err(sd, F, ref_sd, ref_F):
    # scale sqrt(F_ii) * SNR with SNR^2 = F[dL,dL] dL^2 replaced by the Phicoal diagonal (= SNR^2 for the (2,2) models, ~SNR^2 for HM)
    scale = np.sqrt(np.abs(np.einsum('iin->in', ref_F)) * np.abs(ref_F[8, 8])[None, :]) + 1e-300
    return float(np.max(np.abs(sd - ref_sd) / scale))


@pytest.mark.skipif(shutil.which('nvcc') is None, reason='nvcc needed to build the emulation harness')
@pytest.mark.parametrize('name', ['deriv_c2', 'deriv_tf2tidal', 'deriv_m1m2_lvk', 'deriv_hm_lvk', 'deriv_nsbh'])
def test_emulated_snr_derivatives_match_reference(name):
    import emu_driver as E
    from test_device_math_emulated import _emu_inputs
    from gwfast_b200 import signal, _capi as K
    cfg, ev, out = load_golden(name)
    model, dets, psds = _emu_inputs(cfg)
    fkw = cfg.get('fisher_kw', {})
    flags = (K.GWF_OPT_M1M2 if fkw.get('use_m1m2') else 0) | (0 if fkw.get('use_chi1chi2', True) else K.GWF_OPT_CHIS_CHIA)
    e2 = signal._engine_events(model, ev, None, bool(fkw.get('use_m1m2')))
    packed, s2, sd = E.run(model._descriptor(ev), dets, psds, e2, res=cfg.get('res', 1000), flags=flags, per_arm=True, snr_derivs=True)
    F = E.unpack(packed, model.nParams)
    arms = _arms(out)
    assert len(arms) == F.shape[0]
    for i, k in enumerate(arms):
        assert _sd_err(sd[i].T, F[i], out['snrderiv__' + k], out['fisher__' + k]) < SD_TOL, k
        # d h/d dL = -h/dL  =>  (h | d_dL h) = -SNR^2/dL exactly (non-HM), a check that does not involve the reference
        if model.objType and not model.is_HigherModes:
            assert np.allclose(sd[i][:, 2], -s2[i] / ev['dL'], rtol=1e-12)
            assert np.allclose(sd[i][:, 8], 0.0, atol=1e-12 * np.abs(s2[i]).max())


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['deriv_c2', 'deriv_tf2tidal', 'deriv_m1m2_lvk', 'deriv_hm_lvk', 'deriv_nrtidal', 'deriv_nsbh'])
def test_engine_snr_derivatives_match_reference(name):
    cfg, ev, out = load_golden(name)
    net = make_network('engine', cfg)
    F, SD = net.FisherMatr(copy_events(ev), res=cfg.get('res', 1000), return_SNR_derivatives=True, **cfg.get('fisher_kw', {}))
    arms = _arms(out)
    assert set(F) == set(arms) | {'net'} and set(SD) == set(arms) | {'net'}
    masked = name == 'deriv_nrtidal'          # raw reference: last-sample artefact, compare at its size (App. A-3)
    for k in arms:
        assert SD[k].shape == out['snrderiv__' + k].shape
        assert _sd_err(SD[k], F[k], out['snrderiv__' + k], out['fisher__' + k]) < (5e-3 if masked else SD_TOL), k
    snr = net.SNR(copy_events(ev))            # the reference divides by self.SNR(evParams) at its default res=1000 (network.py:141)
    scale = np.sqrt(np.abs(np.einsum('iin->in', out['fisher__net'])))
    assert np.max(np.abs(SD['net'] - out['snrderiv__net']) / scale) < (5e-3 if masked else SD_TOL)
    # 'net' = d SNR_net / d p_i: sum over arms of (h | d_i h) divided by the network SNR (network.py:143)
    assert np.allclose(SD['net'], sum(SD[k] for k in arms) / snr, rtol=1e-12, atol=0)
    # the SNRs keyword short-cuts the extra SNR evaluation (network.py:138-141)
    _, SD2 = net.FisherMatr(copy_events(ev), res=cfg.get('res', 1000), return_SNR_derivatives=True, SNRs=snr, **cfg.get('fisher_kw', {}))
    assert np.array_equal(SD2['net'], SD['net'])


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['deriv_c2', 'deriv_tf2tidal', 'deriv_m1m2_lvk', 'deriv_nrtidal', 'deriv_hm_strain_lvk', 'deriv_hm_strain_et', 'deriv_nsbh'])
def test_engine_strain_derivatives_match_reference(name):
    cfg, ev, out = load_golden(name)
    net = make_network('engine', cfg)
    res = cfg.get('res', 1000)
    F, D = net.FisherMatr(copy_events(ev), res=res, return_derivatives=True, **cfg.get('fisher_kw', {}))
    arms = [k[len('deriv__'):] for k in out if k.startswith('deriv__')]
    assert set(D) == set(arms)
    last = -1 if name == 'deriv_nrtidal' else None
    for k in arms:
        ref = out['deriv__' + k]
        assert D[k].shape == ref.shape and D[k].dtype == np.complex128          # (nP, N, res)
        scale = np.max(np.abs(ref[..., :last]), axis=-1, keepdims=True) + 1e-300
        assert np.max(np.abs(D[k][..., :last] - ref[..., :last]) / scale) < D_TOL, k
    # the Fisher returned next to them is the per-arm Fisher, and it is the Gram of these very arrays (signal.py:922-931)
    from gwfast_b200 import synthetic
    if name == 'deriv_c2':
        k = 'CE1Id'
        s = net.signals[k]
        fcut = s.wf_model.fcut(**copy_events(ev))
        fgrid = np.geomspace(np.full(len(fcut), s.fmin), fcut, num=res)
        Sn = np.interp(fgrid, s.strainFreq, s.noiseCurve, left=1., right=1.)
        G = 4 * np.trapezoid(np.real(np.conj(D[k][:, None]) * D[k][None, :]) / Sn.T[None, None], fgrid.T[None, None], axis=-1)
        dg = np.sqrt(np.einsum('iin->in', F[k]))
        assert np.max(np.abs(G - F[k]) / (dg[:, None] * dg[None, :])) < 1e-9


@pytest.mark.gpu
def test_single_detector_derivative_api():
    """GWSignal.FisherMatr returns (allFishers, allDerivs) lists, one entry per arm (IMRPhenomHM included)."""
    from gwfast_b200 import waveforms, signal, synthetic
    ev = synthetic.bbh_catalog(5, 77)
    s = synthetic.build_network(signal.GWSignal, waveforms.IMRPhenomD(), 'ET')['ET']
    Fs, Ds = s.FisherMatr(copy_events(ev), res=64, return_derivatives=True)
    assert len(Fs) == 3 and len(Ds) == 3 and Ds[0].shape == (11, 5, 64)
    assert np.allclose(Ds[2], -(Ds[0] + Ds[1]), rtol=1e-9, atol=1e-12 * np.abs(Ds[0]).max())      # signal.py:1057
    Fs2, Ss = s.FisherMatr(copy_events(ev), res=64, return_SNR_derivatives=True)
    assert len(Ss) == 3 and Ss[0].shape == (11, 5)
    for a, b in zip(Fs, Fs2):
        assert np.array_equal(a, b)
    hm = synthetic.build_network(signal.GWSignal, waveforms.IMRPhenomHM(), 'ET')['ET']
    Fh, Dh = hm.FisherMatr(copy_events(ev), res=64, return_derivatives=True)
    assert len(Dh) == 3 and Dh[0].shape == (11, 5, 64) and np.allclose(Dh[2], -(Dh[0] + Dh[1]), rtol=1e-9, atol=1e-12 * np.abs(Dh[0]).max())
