"""``WaveFormModel.Phi / Ampl / tau_star / fcut / hphc`` on a user frequency grid, evaluated on the GPU (``gwf_waveform``).

Shapes follow the reference's numpy broadcasting: ``f`` may be ``(res,)`` with scalar or ``(1,)`` parameters,
or ``(res, N)`` with ``(N,)`` parameters; ``IMRPhenomHM.Phi/Ampl`` return dicts keyed '21','22','32','33','43','44'.
"""
import ctypes as C

import numpy as np

from . import _capi as K
from . import _engine
from . import gwfastGlobals as glob

HM_KEYS = ('21', '22', '32', '33', '43', '44')


def evaluate(model, f, what, ev):
    st = _engine.state()
    torch = st.torch
    lib = st.lib
    Mc = np.atleast_1d(np.asarray(ev['Mc'], dtype=np.float64))
    n = Mc.shape[0]
    scalar_params = np.ndim(ev['Mc']) == 0
    events = {}
    for k in K.EVENT_KEYS[:13]:
        if k in ev:
            events[k] = np.broadcast_to(np.atleast_1d(np.real(np.asarray(ev[k]))).astype(np.float64), (n,))
        elif k in ('dL', 'theta', 'phi', 'iota', 'psi', 'tcoal', 'Phicoal'):
            events[k] = np.zeros(n) if k != 'dL' else np.ones(n)          # not needed by Phi / tau_star / fcut
    if model.is_tidal and 'Lambda1' not in events:
        events['Lambda1'] = np.zeros(n)                                    # waveforms.py:1397-1399
        events['Lambda2'] = np.zeros(n)
    if getattr(model, 'is_eccentric', False):
        events['ecc'] = np.broadcast_to(np.atleast_1d(np.real(np.asarray(ev['ecc']))).astype(np.float64), (n,))
    if model._model_id != K.GWF_TAYLORF2:
        events['_Mtot_sec'] = (events['Mc'] / (events['eta'] ** (3. / 5.))) * glob.GMsun_over_c3
    desc = model._descriptor(ev)
    dev_ev, host_ev, evs, _ = _engine._upload(st, events, n, K.EVENT_KEYS)
    stream = torch.cuda.current_stream(st.device)
    sp = C.c_void_p(stream.cuda_stream)
    ws = _engine._workspace(st, lib.gwf_workspace_bytes(C.byref(desc), n))

    def ptr(t):
        return C.c_void_p(t.data_ptr()) if t is not None else None

    if what == 'fcut':
        out = torch.empty(n, dtype=torch.float64, device=st.device)
        K.check(lib.gwf_waveform(C.byref(desc), C.byref(evs), n, None, 0, 0, None, None, None, None, ptr(out), ptr(ws), ws.numel(), sp), 'gwf_waveform')
        r = out.cpu().numpy()
        return r[0] if scalar_params else r
    fa = np.asarray(np.real(f), dtype=np.float64)
    if fa.ndim == 0:
        fa = fa.reshape(1)
    if fa.ndim == 2 and fa.shape[1] != n:
        if n == 1:
            raise ValueError('gwfast_b200: a (res, N) grid needs (N,) parameters')
        raise ValueError('frequency grid and parameters have incompatible shapes')
    res = fa.shape[0]
    f2d = fa.ndim == 2
    fd = torch.from_numpy(np.ascontiguousarray(fa)).to(st.device)
    nm = 6 if model._model_id == K.GWF_IMRPHENOMHM else 1
    phi = torch.empty((nm, res, n), dtype=torch.float64, device=st.device) if what == 'phi' else None
    amp = torch.empty((nm, res, n), dtype=torch.float64, device=st.device) if what == 'ampl' else None
    tau = torch.empty((res, n), dtype=torch.float64, device=st.device) if what == 'tau' else None
    hphc = torch.empty((4, res, n), dtype=torch.float64, device=st.device) if what == 'hphc' else None
    K.check(lib.gwf_waveform(C.byref(desc), C.byref(evs), n, ptr(fd), res, int(f2d), ptr(phi), ptr(amp), ptr(tau), ptr(hphc), None, ptr(ws), ws.numel(), sp),
            'gwf_waveform')
    _engine.launch_count += 3

    def shape(a):               # (res, n) -> what numpy broadcasting of f with the parameters gives
        if f2d:
            return a
        if n == 1:
            return a[:, 0]
        return a

    if what == 'hphc':
        h = hphc.cpu().numpy()
        return shape(h[0] + 1j * h[1]), shape(h[2] + 1j * h[3])
    t = {'phi': phi, 'ampl': amp, 'tau': tau}[what].cpu().numpy()
    if what == 'tau':
        return shape(t)
    if nm == 1:
        return shape(t[0])
    return {k: shape(t[i]) for i, k in enumerate(HM_KEYS)}


def _event_table(model, ev, n, need_angles=True):
    """the parameter arrays gwf_waveform / gwf_signal_grid consume, broadcast to (n,)"""
    events = {}
    for k in K.EVENT_KEYS[:13]:
        if k in ev:
            events[k] = np.broadcast_to(np.atleast_1d(np.real(np.asarray(ev[k]))).astype(np.float64), (n,))
        elif not need_angles and k in ('dL', 'theta', 'phi', 'iota', 'psi', 'tcoal', 'Phicoal'):
            events[k] = np.zeros(n) if k != 'dL' else np.ones(n)
    if model.is_tidal and 'Lambda1' not in events:
        events['Lambda1'] = np.zeros(n)                                    # waveforms.py:1397-1399
        events['Lambda2'] = np.zeros(n)
    if getattr(model, 'is_eccentric', False):
        events['ecc'] = np.broadcast_to(np.atleast_1d(np.real(np.asarray(ev['ecc']))).astype(np.float64), (n,))
    if model._model_id != K.GWF_TAYLORF2:
        events['_Mtot_sec'] = (events['Mc'] / (events['eta'] ** (3. / 5.))) * glob.GMsun_over_c3
    return events


def signal_grid(model, det, rot_deg, f, ev, want):
    """GWSignal.GWAmplitudes / GWPhase / GWstrain on a user grid (``gwf_signal_grid``).

    ``want``: names among Ap, Ac, psi, strain, Fp, Fc, dt; returns a dict of numpy arrays shaped like numpy broadcasting of ``f``
    (``(res,)`` or ``(res, N)``) with the parameters gives (``strain`` complex128)."""
    st = _engine.state()
    torch = st.torch
    lib = st.lib
    n = np.atleast_1d(np.asarray(ev['Mc'])).shape[0]
    events = _event_table(model, ev, n)
    desc = model._descriptor(ev)
    dev_ev, host_ev, evs, _ = _engine._upload(st, events, n, K.EVENT_KEYS)
    stream = torch.cuda.current_stream(st.device)
    sp = C.c_void_p(stream.cuda_stream)
    ws = _engine._workspace(st, lib.gwf_workspace_bytes(C.byref(desc), n))
    fa = np.asarray(np.real(f), dtype=np.float64)
    if fa.ndim == 0:
        fa = fa.reshape(1)
    if fa.ndim == 2 and fa.shape[1] != n:
        raise ValueError('frequency grid and parameters have incompatible shapes')
    res, f2d = fa.shape[0], fa.ndim == 2
    fd = torch.from_numpy(np.ascontiguousarray(fa)).to(st.device)
    bufs = {k: torch.empty((res, n, 2) if k == 'strain' else (res, n), dtype=torch.float64, device=st.device) for k in want}
    out = K.gwf_signal_out(*[bufs[k].data_ptr() if k in bufs else None for k in ('Ap', 'Ac', 'psi', 'strain', 'Fp', 'Fc', 'dt')])
    K.check(lib.gwf_signal_grid(C.byref(desc), C.byref(det), float(rot_deg), C.byref(evs), n, C.c_void_p(fd.data_ptr()), res, int(f2d), C.byref(out),
                                C.c_void_p(ws.data_ptr()), ws.numel(), sp), 'gwf_signal_grid')
    _engine.launch_count += 3
    r = {}
    for k, t in bufs.items():
        a = (torch.view_as_complex(t) if k == 'strain' else t).cpu().numpy()
        r[k] = a if (f2d or n > 1) else a[:, 0]
    return r


def pattern(det, rot_deg, theta, phi, t, psi, want):
    """GWSignal._PatternFunction / _DeltLoc (``gwf_pattern``): element-wise over the broadcast of the arguments; returns a dict."""
    st = _engine.state()
    torch = st.torch
    lib = st.lib
    args = [np.asarray(np.real(a), dtype=np.float64) for a in ((theta, phi, t) if psi is None else (theta, phi, t, psi))]
    shape = np.broadcast(*args).shape
    flat = [np.ascontiguousarray(np.broadcast_to(a, shape)).reshape(-1) for a in args]
    m = flat[0].shape[0]
    dev = torch.from_numpy(np.stack(flat)).to(st.device)
    bufs = {k: torch.empty(m, dtype=torch.float64, device=st.device) for k in want}
    sp = C.c_void_p(torch.cuda.current_stream(st.device).cuda_stream)

    def ptr(k):
        return C.c_void_p(bufs[k].data_ptr()) if k in bufs else None

    K.check(lib.gwf_pattern(C.byref(det), float(rot_deg), C.c_void_p(dev[0].data_ptr()), C.c_void_p(dev[1].data_ptr()), C.c_void_p(dev[2].data_ptr()),
                            C.c_void_p(dev[3].data_ptr()) if psi is not None else None, m, ptr('Fp'), ptr('Fc'), ptr('dt'), sp), 'gwf_pattern')
    _engine.launch_count += 1
    out = {}
    for k, b in bufs.items():
        a = b.cpu().numpy().reshape(shape)
        out[k] = a if shape else float(a)
    return out
