"""Host-side floor of the end-to-end call: DetNet.FisherMatr(return_SNR=True) on tiny batches (the GPU work is a few tens of us),
and a timeline of one 1e4-event call (time of each host stage, GPU idle gaps excluded)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np, torch
from gwfast_b200 import waveforms, signal, network, synthetic, _engine
wf = waveforms.IMRPhenomD()
net = network.DetNet(synthetic.build_network(signal.GWSignal, wf, 'ET+2CE', useEarthMotion=True, fmin=2.), verbose=False)
for n in (16, 592, 2368, 10000):
    ev = synthetic.bbh_catalog(n, synthetic.SEEDS['C2'])
    for _ in range(5):
        net.FisherMatr(dict(ev), res=1000, return_SNR=True)
    torch.cuda.synchronize()
    ts = []
    for _ in range(30):
        torch.cuda.synchronize()
        t = time.perf_counter()
        net.FisherMatr(dict(ev), res=1000, return_SNR=True)
        ts.append(time.perf_counter() - t)
    print('n = %5d: %.3f ms per call (min %.3f)' % (n, 1e3 * np.median(ts), 1e3 * min(ts)), flush=True)
# stage timings inside one call
import ctypes as C
ev = synthetic.bbh_catalog(10000, synthetic.SEEDS['C2'])
st = _engine.state()
marks = []
orig_stage = _engine._stage
def stage(*a, **k):
    marks.append(('stage in', time.perf_counter())); r = orig_stage(*a, **k); marks.append(('stage out', time.perf_counter())); return r
_engine._stage = stage
lib = st.lib
for name in ('gwf_fisher_range', 'gwf_fisher_ex', 'gwf_unpack_fisher_ld', 'gwf_copy_2d'):
    f = getattr(lib, name)
    def mk(f, name):
        def w(*a):
            marks.append((name + ' in', time.perf_counter())); r = f(*a); marks.append((name + ' out', time.perf_counter())); return r
        return w
    setattr(lib, name, mk(f, name))
for rep in range(3):
    marks.clear()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    net.FisherMatr(dict(ev), res=1000, return_SNR=True)
    t1 = time.perf_counter()
print('timeline of one call (ms from entry):')
for k, t in marks:
    print('   %-28s %.3f' % (k, 1e3 * (t - t0)))
print('   %-28s %.3f' % ('return', 1e3 * (t1 - t0)))
