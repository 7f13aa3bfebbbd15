"""ctypes driver of tests/emu/libgwfast_emu.so (CPU emulation of the device math; TEST INFRASTRUCTURE)."""
import ctypes as C
import os
import subprocess
import numpy as np

from gwfast_b200 import _capi as K

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.environ.get("GWF_EMU_SO", os.path.join(_HERE, "emu", "libgwfast_emu.so"))
_lib = None


def build(force=False):
    src = os.path.join(_HERE, 'emu', 'emu.cu')
    csrc = os.path.join(os.path.dirname(_HERE), 'gwfast_b200', 'csrc')
    deps = [src] + [os.path.join(csrc, f) for f in os.listdir(csrc)]
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(['nvcc', '-O2', '-std=c++17', '-Wno-deprecated-gpu-targets', '-Xcompiler', '-fPIC', '-shared', '-o', _SO, src])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.emu_last_error.restype = C.c_char_p
        wf = os.path.join(os.path.dirname(_HERE), 'gwfast_b200', 'data', 'WFfiles')
        a, fr, fd = (np.ascontiguousarray(np.loadtxt(os.path.join(wf, 'QNMData_%s.txt' % k))) for k in ('a', 'fring', 'fdamp'))
        dp = C.POINTER(C.c_double)
        _lib.emu_set_qnm(a.ctypes.data_as(dp), fr.ctypes.data_as(dp), fd.ctypes.data_as(dp), len(a))
    return _lib


def run(model, dets, psds, ev, res=1000, flags=0, per_arm=False, snr_mode=False, snr_derivs=False):
    """model: gwf_model; dets: list of gwf_detector; psds: list of (f, S); ev: dict of arrays. Returns (packed, snr2)
    (or (packed, snr2, sd[npass, n, nP]) with snr_derivs)."""
    L = lib()
    n = len(ev['Mc'])
    dp = C.POINTER(C.c_double)
    arrs = [np.ascontiguousarray(ev[k], dtype=float) if k in ev else None for k in K.EVENT_KEYS]
    evp = (dp * K.GWF_NPARAM_IN)(*[a.ctypes.data_as(dp) if a is not None else None for a in arrs])
    pf = [np.ascontiguousarray(p[0], dtype=float) for p in psds]
    pS = [np.ascontiguousarray(p[1], dtype=float) for p in psds]
    pfp = (dp * len(psds))(*[a.ctypes.data_as(dp) for a in pf])
    pSp = (dp * len(psds))(*[a.ctypes.data_as(dp) for a in pS])
    pn = (C.c_int * len(psds))(*[len(a) for a in pf])
    darr = (K.gwf_detector * len(dets))(*dets)
    opts = K.gwf_opts(res, flags, int(per_arm), 0)
    nP = {0: (13 if model.flags & K.GWF_MODEL_TIDAL else 11) + (1 if model.flags & K.GWF_MODEL_ECCENTRIC else 0), 1: 11, 2: 13, 3: 11, 4: 13}[model.id]
    npack = nP * (nP + 1) // 2
    narms = sum(1 if d.shape == 0 else 3 for d in dets)
    if snr_mode:
        out = np.zeros((narms, n))
        s2 = None
    else:
        npass = narms if per_arm else 1
        out = np.zeros((npass, n, npack))
        s2 = np.zeros((npass, n))
    if snr_derivs:
        sd = np.zeros((npass, n, nP))
        rc = L.emu_fisher_sd(C.byref(model), darr, len(dets), pfp, pSp, pn, len(psds), evp, C.c_longlong(n), C.byref(opts),
                             out.ctypes.data_as(dp), s2.ctypes.data_as(dp), sd.ctypes.data_as(dp))
        if rc != 0:
            raise RuntimeError('emu failed %d: %s' % (rc, L.emu_last_error().decode()))
        return out, s2, sd
    rc = L.emu_fisher(C.byref(model), darr, len(dets), pfp, pSp, pn, len(psds), evp, C.c_longlong(n), C.byref(opts),
                      out.ctypes.data_as(dp), s2.ctypes.data_as(dp) if s2 is not None else None, int(snr_mode))
    if rc != 0:
        raise RuntimeError('emu failed %d: %s' % (rc, L.emu_last_error().decode()))
    return out, s2


def unpack(packed, nP):
    """[..., n, npack] -> (..., nP, nP, n)"""
    n = packed.shape[-2]
    F = np.zeros(packed.shape[:-2] + (nP, nP, n))
    for i in range(nP):
        for j in range(i + 1):
            F[..., i, j, :] = F[..., j, i, :] = packed[..., i * (i + 1) // 2 + j]
    return F


def waveform(model, ev, f, want=('phi', 'ampl', 'tau')):
    """CPU emulation of gwf_waveform: f (res,) or (res, n); returns dict of arrays [nm, res, n] / [res, n] / hphc complex."""
    L = lib()
    n = len(ev['Mc'])
    dp = C.POINTER(C.c_double)
    arrs = [np.ascontiguousarray(ev[k], dtype=float) if k in ev else None for k in K.EVENT_KEYS]
    evp = (dp * K.GWF_NPARAM_IN)(*[a.ctypes.data_as(dp) if a is not None else None for a in arrs])
    f = np.ascontiguousarray(f, dtype=float)
    res = f.shape[0]
    nm = 6 if model.id == 3 else 1
    out = {k: np.zeros((nm, res, n)) for k in ('phi', 'ampl') if k in want}
    if 'tau' in want:
        out['tau'] = np.zeros((res, n))
    if 'hphc' in want:
        out['hphc'] = np.zeros((4, res, n))
    out['fcut'] = np.zeros(n)
    ptr = lambda k: out[k].ctypes.data_as(dp) if k in out else None
    rc = L.emu_waveform(C.byref(model), evp, C.c_longlong(n), f.ctypes.data_as(dp), res, int(f.ndim == 2), ptr('phi'), ptr('ampl'), ptr('tau'),
                        ptr('hphc'), ptr('fcut'))
    if rc != 0:
        raise RuntimeError('emu_waveform failed %d' % rc)
    return out


def covariance(F, method=0, thresh=1e-15):
    """csrc/covariance.cuh:cov_one on the host: F (nP,nP,N) -> (cov, inv_err, status)."""
    L = lib()
    F = np.ascontiguousarray(F, dtype=float)
    nP, _, n = F.shape
    cov, err, st = np.zeros_like(F), np.zeros(n), np.zeros(n, dtype=np.int32)
    dp = C.POINTER(C.c_double)
    L.emu_covariance(F.ctypes.data_as(dp), C.c_longlong(n), nP, method, C.c_double(thresh), cov.ctypes.data_as(dp), err.ctypes.data_as(dp),
                     st.ctypes.data_as(C.POINTER(C.c_int)))
    return cov, err, st


def eigen(F):
    """csrc/covariance.cuh:eig_one on the host: F (nP,nP,N) -> (evals (nP,N), evecs (nP,nP,N), cond (N,))."""
    L = lib()
    F = np.ascontiguousarray(F, dtype=float)
    nP, _, n = F.shape
    ev, vec, cond = np.zeros((nP, n)), np.zeros_like(F), np.zeros(n)
    dp = C.POINTER(C.c_double)
    L.emu_eigen(F.ctypes.data_as(dp), C.c_longlong(n), nP, ev.ctypes.data_as(dp), vec.ctypes.data_as(dp), cond.ctypes.data_as(dp))
    return ev, vec, cond
