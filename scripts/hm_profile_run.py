import os, sys, ctypes as C
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from gwfast_b200 import waveforms, signal, network, synthetic
n = 4000
ev = synthetic.bbh_catalog(n, synthetic.SEEDS['C4'])
wf = waveforms.IMRPhenomHM()
sigs = synthetic.build_network(signal.GWSignal, wf, 'LVK-O4', useEarthMotion=False, fmin=10.)
net = network.DetNet(sigs, verbose=False)
for _ in range(3):
    F = net.FisherMatr(dict(ev))
torch.cuda.synchronize()
print(F.shape)
