import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np
from gwfast_b200 import waveforms, signal, network, synthetic
ev = synthetic.bns_catalog(10000, synthetic.SEEDS['C3'], tidal=True)
net = network.DetNet(synthetic.build_network(signal.GWSignal, waveforms.IMRPhenomD_NRTidalv2(), 'ET+2CE'), verbose=False)
F = net.FisherMatr(dict(ev)); snr = net.SNR(dict(ev))
bad = np.where(~np.isfinite(F).all(axis=(0, 1)))[0]
print('non-finite Fisher events:', len(bad), bad[:10], 'non-finite SNR:', int((~np.isfinite(snr)).sum()))
for b in bad[:5]:
    print({k: float(v[b]) for k, v in ev.items()})
    print(np.argwhere(~np.isfinite(F[:, :, b]))[:6].tolist())
np.save('/root/repo/gpurun_out/nan_idx.npy', bad)
