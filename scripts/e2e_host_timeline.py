"""Host timeline of one end-to-end call (DetNet.FisherMatr(return_SNR=True), 1e4 events): when, after entry, each C-ABI call is issued."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np, torch
from gwfast_b200 import waveforms, signal, network, synthetic, _engine, _capi as K
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
ev = synthetic.bbh_catalog(n, synthetic.SEEDS['C2'])
net = network.DetNet(synthetic.build_network(signal.GWSignal, waveforms.IMRPhenomD(), 'ET+2CE', useEarthMotion=True, fmin=2.), verbose=False)
call = lambda: net.FisherMatr(dict(ev), res=1000, return_SNR=True)
for _ in range(10):
    call()
lib = K.load()
log = []
class Wrap(object):
    def __init__(self, name, f):
        self.name, self.f = name, f
    def __call__(self, *a):
        t = time.perf_counter(); r = self.f(*a); log.append((self.name, t, time.perf_counter())); return r
for name in ('gwf_fisher_range', 'gwf_fisher_ex', 'gwf_unpack_fisher_ld', 'gwf_copy_2d'):
    setattr(lib, name, Wrap(name, getattr(lib, name)))
orig_stage, orig_fisher = _engine._stage, _engine.fisher
def stage(*a, **k):
    t = time.perf_counter(); r = orig_stage(*a, **k); log.append(('_stage', t, time.perf_counter())); return r
_engine._stage = stage
orig_sync = torch.cuda.Stream.synchronize
def sync(self):
    t = time.perf_counter(); r = orig_sync(self); log.append(('sync', t, time.perf_counter())); return r
torch.cuda.Stream.synchronize = sync
rows = {}
for rep in range(30):
    del log[:]
    t0 = time.perf_counter(); call(); t1 = time.perf_counter()
    for i, (nm, a, b) in enumerate(log):
        rows.setdefault((i, nm), []).append(((a - t0) * 1e6, (b - a) * 1e6))
    rows.setdefault((99, 'return'), []).append(((t1 - t0) * 1e6, 0.))
for (i, nm), v in sorted(rows.items()):
    a = np.median([x[0] for x in v]); d = np.median([x[1] for x in v])
    print('%2d %-22s issued at %7.1f us  took %6.1f us' % (i, nm, a, d))
