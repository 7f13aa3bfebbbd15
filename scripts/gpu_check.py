"""Quick on-GPU sanity run: parity against the oracle port on a few configurations + a first timing."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from gwfast_b200 import waveforms, signal, network, synthetic, _engine
from oracle.port import waveforms as PW, detector as PD


def err(F, Fo):
    dg = np.sqrt(np.einsum('iin->in', Fo))
    return float(np.max(np.abs(F - Fo) / (dg[:, None, :] * dg[None, :, :])))


print(torch.cuda.get_device_name(0))
N = 24
for name, M, PM, cat in (('TF2', waveforms.TaylorF2_RestrictedPN, PW.TaylorF2_RestrictedPN, synthetic.bbh_catalog),
                         ('PhD', waveforms.IMRPhenomD, PW.IMRPhenomD, synthetic.bbh_catalog)):
    ev = cat(N, 7)
    for netname in ('ETSL', 'ET', 'ET+2CE'):
        for rot in (False, True):
            net = network.DetNet(synthetic.build_network(signal.GWSignal, M(), netname, useEarthMotion=rot), verbose=False)
            onet = PD.Network(synthetic.build_network(PD.Detector, PM(), netname, useEarthMotion=rot))
            snr, F = net.SNR(dict(ev)), net.FisherMatr(dict(ev))
            snr_o, F_o = onet.SNR(dict(ev)), onet.FisherMatr(dict(ev))
            print('%s %-7s rot=%d  snr %.2e  fisher %.2e' % (name, netname, rot, np.max(np.abs(snr / snr_o - 1)), err(F, F_o)))
    # return_all path
    net = network.DetNet(synthetic.build_network(signal.GWSignal, M(), 'ET+2CE'), verbose=False)
    onet = PD.Network(synthetic.build_network(PD.Detector, PM(), 'ET+2CE'))
    Fa, Fo = net.FisherMatr(dict(ev), return_all=True), onet.FisherMatr(dict(ev), return_all=True)
    Sa, So = net.SNR(dict(ev), return_all=True), onet.SNR(dict(ev), return_all=True)
    print(name, 'return_all', {k: '%.1e/%.1e' % (np.max(np.abs(Sa[k] / So[k] - 1)), err(Fa[k], Fo[k])) for k in Fo})

# timing
ev = synthetic.bbh_catalog(10000, synthetic.SEEDS['C2'])
net = network.DetNet(synthetic.build_network(signal.GWSignal, waveforms.IMRPhenomD(), 'ET+2CE'), verbose=False)
for it in range(4):
    torch.cuda.synchronize(); t = time.time()
    snr = net.SNR(dict(ev)); torch.cuda.synchronize(); t1 = time.time()
    F = net.FisherMatr(dict(ev)); torch.cuda.synchronize(); t2 = time.time()
    print('N=10000 ET+2CE PhenomD: SNR %.2f ms, Fisher %.2f ms -> %.3g events/s' % ((t1 - t) * 1e3, (t2 - t1) * 1e3, 10000 / (t2 - t)))
print('dL check', np.max(np.abs(F[2, 2] * ev['dL'] ** 2 / snr ** 2 - 1)))
