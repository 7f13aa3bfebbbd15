"""End-to-end step (DetNet.FisherMatr(return_SNR=True), numpy in / numpy out, 1e4 events) under different pipeline settings, with a
host-side profile of the default one."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np, torch
from gwfast_b200 import waveforms, signal, network, synthetic, _engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
ev = synthetic.bbh_catalog(n, synthetic.SEEDS['C2'])
wf = waveforms.IMRPhenomD()
net = network.DetNet(synthetic.build_network(signal.GWSignal, wf, 'ET+2CE', useEarthMotion=True, fmin=2.), verbose=False)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')
def loop(steps=30):
    tot = 0.0
    for _ in range(steps):
        flush.fill_(1); torch.cuda.synchronize()
        t = time.perf_counter()
        r = net.FisherMatr(dict(ev), res=1000, return_SNR=True)
        torch.cuda.synchronize()
        tot += time.perf_counter() - t
    return 1e3 * tot / steps
for g in (4, 1, 2, 3, 4, 6, 8):
    _engine.MAX_GROUPS = g
    for _ in range(5):
        net.FisherMatr(dict(ev), res=1000, return_SNR=True)
    print('MAX_GROUPS %d groups %s: %.3f ms per step' % (g, _engine._round_groups(n, 592), loop()), flush=True)
_engine.MAX_GROUPS = 4
for _ in range(5):
    net.FisherMatr(dict(ev), res=1000, return_SNR=True)
pr = cProfile.Profile(); pr.enable()
for _ in range(30):
    net.FisherMatr(dict(ev), res=1000, return_SNR=True)
pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(22)
