"""Where the end-to-end step of bench.py spends its time: the public-API loop (DetNet.SNR + DetNet.FisherMatr, numpy in / numpy out)
with the bench's L2 flush and clock sampler switched on and off."""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import bench  # noqa: E402
from gwfast_b200 import waveforms, signal, network, synthetic  # noqa: E402

n = 10000
ev = synthetic.bbh_catalog(n, synthetic.SEEDS['C2'])
wf = waveforms.IMRPhenomD()
sigs = synthetic.build_network(signal.GWSignal, wf, 'ET+2CE', useEarthMotion=True, fmin=2.)
net = network.DetNet(sigs, verbose=False)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')
for _ in range(5):
    net.SNR(dict(ev), res=1000)
    net.FisherMatr(dict(ev), res=1000)
torch.cuda.synchronize()


def loop(do_flush, steps=20):
    ts, tf = 0.0, 0.0
    for _ in range(steps):
        if do_flush:
            flush.fill_(1)
        torch.cuda.synchronize()
        t = time.perf_counter()
        net.SNR(dict(ev), res=1000)
        t1 = time.perf_counter()
        net.FisherMatr(dict(ev), res=1000)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        ts += t1 - t
        tf += t2 - t1
    return 1e3 * ts / steps, 1e3 * tf / steps


for sampler_on in (False, True):
    s = None
    if sampler_on:
        s = bench.ClockSampler(0)
        s.start()
        time.sleep(0.2)
    for do_flush in (False, True):
        a, b = loop(do_flush)
        print('sampler=%d flush=%d: SNR %.3f ms  FisherMatr %.3f ms  step %.3f ms' % (sampler_on, do_flush, a, b, a + b))
    if s:
        s.stop_flag = True
