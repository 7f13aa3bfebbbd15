"""SURVEY.md 8(d) parity subset: the first 1000 events of the C2 / C1 / C3 synthetic catalogs through the oracle port and through the
engine; prints the worst SNR and Fisher deviations (north-star tolerances: 1e-9, 1e-6)."""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np  # noqa: E402
from gwfast_b200 import waveforms, signal, network, synthetic  # noqa: E402
from oracle.port import waveforms as PW, detector as PD  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
CASES = [('C2', 'IMRPhenomD', 'ET+2CE', lambda n: synthetic.bbh_catalog(10000, synthetic.SEEDS['C2']), {}),
         ('C1', 'TaylorF2_RestrictedPN', 'ETSL', lambda n: synthetic.bns_catalog(10000, synthetic.SEEDS['C1']), {}),
         # NRTidalv2: the reference's last grid sample sits exactly on the end of the Planck taper (0 or 1 by last-bit rounding,
         # SURVEY.md App. A-3); the engine defines the taper as 0 there, and so does the port with taper_end_zero=True
         ('C3', 'IMRPhenomD_NRTidalv2', 'ET+2CE', lambda n: synthetic.bns_catalog(10000, synthetic.SEEDS['C3'], tidal=True), dict(taper_end_zero=True)),
         ('C3 raw taper end', 'IMRPhenomD_NRTidalv2', 'ET+2CE', lambda n: synthetic.bns_catalog(10000, synthetic.SEEDS['C3'], tidal=True), {})]
ONLY = sys.argv[2] if len(sys.argv) > 2 else ''
CASES = [c for c in CASES if c[0].startswith(ONLY)]
for tag, model, netname, cat, kw in CASES:
    ev = {k: v[:N] for k, v in cat(N).items()}
    eng = network.DetNet(synthetic.build_network(signal.GWSignal, getattr(waveforms, model)(), netname, useEarthMotion=True, fmin=2.), verbose=False)
    port = PD.Network(synthetic.build_network(PD.Detector, getattr(PW, model)(**kw), netname, useEarthMotion=True, fmin=2.))
    snr, F = eng.SNR(dict(ev)), eng.FisherMatr(dict(ev))
    t = time.time()
    es = ef = 0.0
    for lo in range(0, N, 100):
        sub = {k: v[lo:lo + 100] for k, v in ev.items()}
        so, Fo = port.SNR(dict(sub)), port.FisherMatr(dict(sub))
        dg = np.sqrt(np.einsum('iin->in', Fo))
        err = np.abs(F[..., lo:lo + 100] - Fo) / (dg[:, None, :] * dg[None, :, :])
        es = max(es, float(np.max(np.abs(snr[lo:lo + 100] / so - 1))))
        ef = max(ef, float(np.max(err)))
    print('%s %s %s: %d events, SNR max rel err %.2e, Fisher max err %.2e  (port %.0f s)' % (tag, model, netname, N, es, ef, time.time() - t), flush=True)
