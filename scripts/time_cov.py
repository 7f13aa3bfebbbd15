"""Time gwf_covariance / gwf_eigen on the C2 Fisher matrices (1e4 events, 11 parameters), inputs resident on the device."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gwfast_b200 import waveforms, signal, network, synthetic, _engine, _capi as K
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
ev = synthetic.bbh_catalog(n, synthetic.SEEDS['C2'])
net = network.DetNet(synthetic.build_network(signal.GWSignal, waveforms.IMRPhenomD(), 'ET+2CE'), verbose=False)
F = net.FisherMatr(dict(ev))
st = _engine.state(); lib = st.lib
dF = torch.from_numpy(np.ascontiguousarray(F)).to(st.device)
cov = torch.empty_like(dF); err = torch.empty(n, dtype=torch.float64, device=st.device); status = torch.empty(n, dtype=torch.int32, device=st.device)
evl = torch.empty((11, n), dtype=torch.float64, device=st.device); cond = torch.empty(n, dtype=torch.float64, device=st.device)
sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
def t(fn):
    for _ in range(2): fn()
    ts = []
    for _ in range(5):
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return np.median(ts)
for m, name in ((0, 'cholesky'), (1, 'eigen')):
    ms = t(lambda: K.check(lib.gwf_covariance(C.c_void_p(dF.data_ptr()), n, 11, m, 1e-15, C.c_void_p(cov.data_ptr()), C.c_void_p(err.data_ptr()), C.c_void_p(status.data_ptr()), sp), 'cov'))
    print('gwf_covariance [%s] %d events: %.3f ms -> %.3g events/s' % (name, n, ms, n / ms * 1e3))
ms = t(lambda: K.check(lib.gwf_eigen(C.c_void_p(dF.data_ptr()), n, 11, C.c_void_p(evl.data_ptr()), None, C.c_void_p(cond.data_ptr()), sp), 'eig'))
print('gwf_eigen %d events: %.3f ms -> %.3g events/s' % (n, ms, n / ms * 1e3))
