"""Host-side profile of the public-API step (DetNet.SNR + DetNet.FisherMatr, numpy in / numpy out) on one GPU."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch  # noqa: E402
from gwfast_b200 import waveforms, signal, network, synthetic  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
ev = synthetic.bbh_catalog(n, synthetic.SEEDS['C2'])
wf = waveforms.IMRPhenomD()
sigs = synthetic.build_network(signal.GWSignal, wf, 'ET+2CE', useEarthMotion=True, fmin=2.)
net = network.DetNet(sigs, verbose=False)
for _ in range(5):
    net.SNR(dict(ev), res=1000)
    net.FisherMatr(dict(ev), res=1000)
torch.cuda.synchronize()
for name, fn in (('SNR', lambda: net.SNR(dict(ev), res=1000)), ('FisherMatr', lambda: net.FisherMatr(dict(ev), res=1000))):
    t = time.perf_counter()
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    print('%s: %.3f ms per call' % (name, (time.perf_counter() - t) / 20 * 1e3))
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    net.SNR(dict(ev), res=1000)
    net.FisherMatr(dict(ev), res=1000)
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(35)
