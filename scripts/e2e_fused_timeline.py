"""Where the end-to-end call spends its time: DetNet.FisherMatr(return_SNR=True), numpy in -> numpy out, 1e4 events (host profile +
wall-clock split around the engine call)."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
from gwfast_b200 import waveforms, signal, network, synthetic, _engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
ev = synthetic.bbh_catalog(n, synthetic.SEEDS['C2'])
net = network.DetNet(synthetic.build_network(signal.GWSignal, waveforms.IMRPhenomD(), 'ET+2CE', useEarthMotion=True, fmin=2.), verbose=False)
call = lambda: net.FisherMatr(dict(ev), res=1000, return_SNR=True)
for _ in range(10):
    call()
torch.cuda.synchronize()
ts = []
for _ in range(50):
    t = time.perf_counter(); call(); ts.append(time.perf_counter() - t)
ts.sort()
print('FisherMatr(return_SNR=True): median %.3f ms, min %.3f ms' % (ts[len(ts) // 2] * 1e3, ts[0] * 1e3))
# host time until the engine's final synchronize
orig = torch.cuda.Stream.synchronize
acc = {'t': 0.0, 'n': 0}
def timed_sync(self):
    t = time.perf_counter(); r = orig(self); acc['t'] += time.perf_counter() - t; acc['n'] += 1; return r
torch.cuda.Stream.synchronize = timed_sync
t = time.perf_counter()
for _ in range(50):
    call()
tot = time.perf_counter() - t
torch.cuda.Stream.synchronize = orig
print('per call: total %.3f ms, of which waiting in stream.synchronize %.3f ms (%d syncs/call) -> host work before the wait %.3f ms' %
      (tot / 50 * 1e3, acc['t'] / 50 * 1e3, acc['n'] // 50, (tot - acc['t']) / 50 * 1e3))
pr = cProfile.Profile(); pr.enable()
for _ in range(50):
    call()
pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(18)
