"""Time the dominant kernel (fisher_kernel, IMRPhenomD, ET+2CE, 1e4 events) of whatever library GWFAST_B200_LIB points to."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gwfast_b200 import waveforms, signal, synthetic, _engine, _capi as K
n, res = 10000, 1000
netname = sys.argv[1] if len(sys.argv) > 1 else 'ET+2CE'
mname = sys.argv[2] if len(sys.argv) > 2 else 'IMRPhenomD'
if mname == 'IMRPhenomD_NRTidalv2':
    ev = synthetic.bns_catalog(n, synthetic.SEEDS['C3'], tidal=True)
elif mname == 'IMRPhenomNSBH':
    ev = synthetic.nsbh_catalog(n, synthetic.SEEDS['NSBH'])
elif mname == 'TaylorF2_RestrictedPN':
    ev = synthetic.bns_catalog(n, synthetic.SEEDS['C1'])
else:
    ev = synthetic.bbh_catalog(n, synthetic.SEEDS['C2'])
wf = getattr(waveforms, mname)()
lvk = netname == 'LVK-O4'
sigs = synthetic.build_network(signal.GWSignal, wf, netname, useEarthMotion=not lvk, fmin=10. if lvk else 2.)
print('==', mname, netname)
st = _engine.state(); lib = st.lib
model = wf._descriptor(ev)
dets = [s._detector_struct(i) for i, s in enumerate(sigs.values())]; handles = [s._psd_handle() for s in sigs.values()]
darr, parr = _engine._call_arrays(dets, handles)
dev_ev, host_ev, evs, _ = _engine._upload(st, signal._engine_events(wf, ev), n, K.EVENT_KEYS)
ws = _engine._workspace(st, lib.gwf_workspace_bytes(C.byref(model), n))
nP = lib.gwf_num_params(C.byref(model)); packed = torch.empty((n, nP * (nP + 1) // 2), dtype=torch.float64, device=st.device); snr2 = torch.empty(n, dtype=torch.float64, device=st.device)
sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
def call(flags):
    o = K.gwf_opts(res, flags, 0, 0)
    K.check(lib.gwf_fisher(C.byref(model), darr, len(dets), parr, len(handles), C.byref(evs), n, C.byref(o), C.c_void_p(packed.data_ptr()), C.c_void_p(snr2.data_ptr()), C.c_void_p(ws.data_ptr()), ws.numel(), sp), 'f')
for extra, tag in ((0, 'fast+pair'), (16, 'generic+pair'), (32, 'fast, 1 warp/event'), (48, 'generic, 1 warp/event'), (64, 'HM: pairs take alternate blocks')):
    for _ in range(3): call(extra)
    ref = packed.clone()
    ts = []
    for _ in range(10):
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); call(extra | K.GWF_OPT_REUSE_WORKSPACE); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print('%s [%s]: fisher_kernel %.3f ms (min %.3f)  checksum %.10e' % (os.environ.get('GWFAST_B200_LIB', 'default'), tag, np.median(ts), min(ts), float(packed.sum())))
# SNR path (prologue + snr_kernel)
narm = lib.gwf_num_arms(darr, len(dets))
s2a = torch.empty((narm, n), dtype=torch.float64, device=st.device)
def call_snr(flags):
    o = K.gwf_opts(res, flags, 0, 0)
    K.check(lib.gwf_snr(C.byref(model), darr, len(dets), parr, len(handles), C.byref(evs), n, C.byref(o), C.c_void_p(s2a.data_ptr()), C.c_void_p(ws.data_ptr()), ws.numel(), sp), 's')
for extra, tag in ((0, 'fast+split'), (16, 'generic+split'), (32, 'fast, 1 warp/event'), (48, 'generic, 1 warp/event')):
    for _ in range(3): call_snr(extra)
    ts = []
    for _ in range(10):
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); call_snr(extra); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print('gwf_snr [%s]: prologue + snr_kernel %.3f ms (min %.3f)  checksum %.10e' % (tag, np.median(ts), min(ts), float(s2a.sum())))
