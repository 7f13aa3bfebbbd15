"""Device plumbing between the gwfast-style Python classes and ``libgwfast_b200.so``.

torch is used for what it is good at here -- device/pinned memory with caching allocators, streams, and (in
``parallel.py``) ``torch.distributed`` -- and for nothing numerical: every O(N*res) operation is a hand-written
CUDA kernel behind the C ABI (``include/gwfast_b200.h``).  There is no CPU path: without the library or without a
CUDA device every call raises :class:`EngineUnavailable`.
"""
import contextlib
import ctypes as C
import os

import numpy as np

from . import _capi as K
from ._capi import EngineUnavailable, EngineError  # noqa: F401
from . import gwfastGlobals as glob

CHUNK_EVENTS = 1 << 16     # events per launch group: bounds the coefficient-record workspace (2-10 KB/event) and sets the grain of the
                           # H2D / kernel / D2H pipeline of a large catalog
N_STREAMS = 2

_states = {}
_NULL_CTX = contextlib.nullcontext()
launch_count = 0           # kernels launched through this module (bench.py reports it as gpu_launches)


class _State:
    def __init__(self, torch, index):
        self.torch = torch
        self.lib = K.load()
        self.device = torch.device('cuda', index)
        with torch.cuda.device(self.device):
            torch.zeros(1, device=self.device)   # make sure the primary context exists before the library's first call
        self.workspace = None
        self.side_ws = {}
        self.streams = None
        self.last_fisher_device = None
        self.last_status = None
        self.round_events = {}


_qnm_set = False
_psds = {}


def state():
    """engine state of the CURRENT CUDA device (one per device: workspaces, side streams); PSD handles and the QNM tables are
    process-wide -- the library uploads them to each device on first use."""
    global _qnm_set
    try:
        import torch
    except ImportError as e:  # pragma: no cover
        raise EngineUnavailable('torch is required for device memory and streams') from e
    lib = K.load()
    if not torch.cuda.is_available():
        raise EngineUnavailable('no CUDA device visible: gwfast_b200 has no CPU fallback')
    idx = torch.cuda.current_device()
    st = _states.get(idx)
    if st is None:
        torch.cuda.init()
        st = _states[idx] = _State(torch, idx)
    if not _qnm_set:
        dp = C.POINTER(C.c_double)
        tabs = [np.ascontiguousarray(np.loadtxt(os.path.join(glob.WFfilesPath, 'QNMData_%s.txt' % k))) for k in ('a', 'fring', 'fdamp')]
        K.check(lib.gwf_set_qnm_tables(tabs[0].ctypes.data_as(dp), tabs[1].ctypes.data_as(dp), tabs[2].ctypes.data_as(dp), len(tabs[0])),
                'gwf_set_qnm_tables')
        _qnm_set = True
    return st


def xitide_table():
    """the (200, 200, 200) xi_tide table of IMRPhenomNSBH as the current device tabulated it (gwf_xitide_table; the reference reads it
    from WFfiles/xiTide_Table_200.h5 or tabulates it with numpy.roots, waveforms.py:3286-3373)."""
    state()
    tab = np.empty((200, 200, 200))
    K.check(K.load().gwf_xitide_table(tab.ctypes.data_as(C.c_void_p)), 'gwf_xitide_table')
    return tab


def psd_handle(freq, S):
    """PSD handle for (strainFreq, noiseCurve); cached on the arrays' content, valid on every device."""
    lib = K.load()
    freq = np.ascontiguousarray(freq, dtype=np.float64)
    S = np.ascontiguousarray(S, dtype=np.float64)
    key = (freq.shape[0], hash(freq.tobytes()), hash(S.tobytes()))
    h = _psds.get(key)
    if h is None:
        dp = C.POINTER(C.c_double)
        out = C.c_void_p()
        K.check(lib.gwf_psd_create(freq.ctypes.data_as(dp), S.ctypes.data_as(dp), freq.shape[0], C.byref(out)), 'gwf_psd_create')
        h = out.value
        _psds[key] = h
    return h


def _workspace(st, nbytes, slot=None):
    """device scratch for the coefficient records: one buffer for the current stream, one per side stream of the pipeline"""
    if slot is None:
        if st.workspace is None or st.workspace.numel() < nbytes:
            st.workspace = st.torch.empty(int(nbytes), dtype=st.torch.uint8, device=st.device)
        return st.workspace
    ws = st.side_ws.get(slot)
    if ws is None or ws.numel() < nbytes:
        ws = st.side_ws[slot] = st.torch.empty(int(nbytes), dtype=st.torch.uint8, device=st.device)
    return ws


def _side_streams(st):
    if st.streams is None:
        st.streams = [st.torch.cuda.Stream(device=st.device) for _ in range(N_STREAMS)]
    return st.streams


def _stage(st, ev, n, keys):
    """events dict -> ONE pinned staging table (present keys, n); returns (pinned tensor, present keys)."""
    torch = st.torch
    present = [k for k in keys if k in ev]
    host = torch.empty((len(present), n), dtype=torch.float64, pin_memory=True)
    hnp = host.numpy()
    for i, k in enumerate(present):
        a = np.asarray(ev[k])
        if a.dtype != np.float64 or a.shape != (n,):
            a = np.broadcast_to(np.real(a).astype(np.float64, copy=False), (n,))
        np.copyto(hnp[i], a)
    return host, present


_KEY_SLOT = {k: i for i, k in enumerate(K.EVENT_KEYS)}


def _events_struct(dev, present, m):
    """gwf_events over a device table (len(present), m)"""
    evs = K.gwf_events()
    base, stride = dev.data_ptr(), m * 8
    for row, k in enumerate(present):
        evs.p[_KEY_SLOT[k]] = base + row * stride
    return evs


def _upload(st, ev, n, keys):
    """events dict -> one pinned staging buffer -> one H2D copy; returns (device tensor, host tensor, gwf_events, bytes)."""
    host, present = _stage(st, ev, n, keys)
    dev = host.to(st.device, non_blocking=True)
    return dev, host, _events_struct(dev, present, n), host.numel() * 8


def _call_arrays(dets, psd_handles):
    darr = (K.gwf_detector * len(dets))(*dets)
    parr = (C.c_void_p * len(psd_handles))(*psd_handles)
    return darr, parr


# When set, fisher() leaves a reference to its device-resident result (npass, nP, nP, n) in state().last_fisher_device, so that a
# multi-GPU caller can all-gather it over NVLink without a host round trip (parallel.DistributedDetNet(gather='device')).
STASH_DEVICE = False

# When set (parallel.PeerGather), fisher() passes this rank's slots in the peers' gathered buffers to the Fisher kernels
# (gwf_fisher_out.peer_fisher): every finished packed row is stored there over NVLink while the rest of the batch is still computing.
PEER = None

# extra gwf_opts.flags OR-ed into every launch (tests set GWF_OPT_GENERIC_LOOP / GWF_OPT_ONE_WARP_PER_EVENT to compare the
# kernel variants; production leaves it 0 and the launcher picks the fastest applicable form)
KERNEL_FLAGS = 0


MAX_GROUPS = 4             # launch groups of a single-chunk batch (D2H of one group behind the kernels of the next)


def _round_groups(m, round_events, max_groups=None):
    """Cut [0, m) into at most `max_groups` launch groups whose boundaries are multiples of `round_events` (what one round of the
    persistent Fisher grid takes): every CTA stays equally loaded, so the cut costs no ragged extra round, and the device->host copy of a
    group hides behind the kernels of the next.  The last group (whose copy is exposed) is the smallest and holds the partial round."""
    rounds = -(-m // round_events)
    g = max(1, min(MAX_GROUPS if max_groups is None else max_groups, rounds // 2))
    if g == 1:
        return [(0, m)]
    q, r = divmod(rounds, g)
    out, lo = [], 0
    for k in range(g):
        hi = min(m, lo + (q + (1 if k < r else 0)) * round_events)
        out.append((lo, hi - lo))
        lo = hi
    return [t for t in out if t[1] > 0]


def fisher(model, dets, psd_handles, ev, n, res, flags, per_arm, want_snr2=True, keep_on_device=False, want_snr_derivs=False,
           want_snr_integ=False):
    """Run the Fisher kernels (+ unpack) on the current device, pipelined.

    Returns ``(F, snr2, io)`` with ``F`` of shape ``(npass, nP, nP, n)`` and ``snr2`` ``(npass, n)`` as numpy arrays
    (or device tensors if ``keep_on_device``); ``io`` = (h2d_bytes, d2h_bytes).  With ``want_snr_derivs`` the second element
    is the pair ``(snr2, snr_derivs)`` with ``snr_derivs`` of shape ``(npass, n, nP)``; with ``want_snr_integ`` it is the integral
    SNRInteg forms (gwf_fisher_out.snr2_integ) instead of 4 int |h|^2/Sn.  The per-event status words of the call are left in
    ``state().last_status`` (numpy int32, or None with ``keep_on_device``).

    The events are staged once in pinned memory.  A batch of up to CHUNK_EVENTS events is ONE H2D copy and ONE prologue, followed by a
    few launch groups cut at multiples of the persistent grid's round (``gwf_fisher_range``): while a group's kernels run, the Fisher
    matrices of the previous group are unpacked and copied to the host on a side stream.  A larger catalog is cut into chunks of
    CHUNK_EVENTS, each running H2D -> prologue -> kernels -> unpack -> D2H on one of two side streams.  The host waits once, at the end.
    """
    global launch_count
    st = state()
    torch = st.torch
    lib = st.lib
    nP = lib.gwf_num_params(C.byref(model))
    if nP < 0:
        raise EngineError('unknown model')
    npack = nP * (nP + 1) // 2
    darr, parr = _call_arrays(dets, psd_handles)
    ndet, npsd = len(dets), len(psd_handles)
    npass = lib.gwf_num_arms(darr, ndet) if per_arm else 1
    cur = torch.cuda.current_stream(st.device)
    f64 = torch.float64
    # the input copy goes first: it is in flight while the host allocates the outputs
    host_ev, present = _stage(st, ev, n, K.EVENT_KEYS)
    nk = len(present)
    chunks = [(lo, min(CHUNK_EVENTS, n - lo)) for lo in range(0, n, CHUNK_EVENTS)]
    multi = len(chunks) > 1
    dev_all = None
    if not multi:
        dev_all = torch.empty((nk, n), dtype=f64, device=st.device)
        dev_all.copy_(host_ev, non_blocking=True)
    full = torch.empty((npass, nP, nP, n), dtype=f64, device=st.device)
    snr2 = torch.empty((npass, n), dtype=f64, device=st.device)
    sder = torch.empty((npass, n, nP), dtype=f64, device=st.device) if want_snr_derivs else None
    status = torch.empty((n,), dtype=torch.int32, device=st.device)
    opts = K.gwf_opts(int(res), int(flags) | KERNEL_FLAGS, int(bool(per_arm)), 0)
    to_host = not keep_on_device
    host_out = []

    def pinned_outputs():
        # allocated when the first copy is about to be queued: by then the kernels are already running
        if not host_out:
            host_out.extend([torch.empty(full.shape, dtype=f64, pin_memory=True), torch.empty(snr2.shape, dtype=f64, pin_memory=True),
                             torch.empty((n,), dtype=torch.int32, pin_memory=True),
                             torch.empty(sder.shape, dtype=f64, pin_memory=True) if want_snr_derivs else None])
        return host_out
    side = _side_streams(st)
    copy_stream = side[-1]
    used = side if multi else [copy_stream]         # a single-chunk batch computes on the current stream and only copies on a side stream
    for s_ in used:
        s_.wait_stream(cur)
    round_events = st.round_events.get(model.id)
    if round_events is None:
        round_events = st.round_events[model.id] = max(1, int(lib.gwf_round_events(C.byref(model))))
    keep = []

    def outputs(packed, s2, sd, lo):
        fo = K.gwf_fisher_out(packed.data_ptr(), None if want_snr_integ else s2, s2 if want_snr_integ else None, sd, status.data_ptr() + 4 * lo)
        if PEER is not None and npass == 1:
            # multi-GPU: the Fisher kernel itself stores every finished row into this rank's slot of the peers' gathered buffers
            fo.peer_fisher = C.cast(PEER.slots(lo), C.c_void_p)
            fo.npeers = PEER.world
        return fo

    def unpack(packed, lo, m, sp):
        global launch_count
        for p in range(npass):
            K.check(lib.gwf_unpack_fisher_ld(C.c_void_p(packed[p].data_ptr()), m, nP, C.c_void_p(full[p].data_ptr() + 8 * lo), n, sp), 'gwf_unpack_fisher')
            launch_count += 1

    def to_host_copy(lo, m, sp):
        out_f = pinned_outputs()[0]
        if m == n:
            out_f.copy_(full, non_blocking=True)
        else:
            K.check(lib.gwf_copy_2d(C.c_void_p(out_f.data_ptr() + 8 * lo), n * 8, C.c_void_p(full.data_ptr() + 8 * lo), n * 8, m * 8,
                                    npass * nP * nP, sp), 'gwf_copy_2d')

    for ci, (clo, cm) in enumerate(chunks):
        stream = side[ci % len(side)] if multi else cur
        sp = C.c_void_p(stream.cuda_stream)
        with (torch.cuda.stream(stream) if multi else _NULL_CTX):
            dev_ev = dev_all if dev_all is not None else torch.empty((nk, cm), dtype=f64, device=st.device)
            if dev_all is not None:
                pass
            else:
                K.check(lib.gwf_copy_2d(C.c_void_p(dev_ev.data_ptr()), cm * 8, C.c_void_p(host_ev.data_ptr() + clo * 8), n * 8, cm * 8, nk, sp), 'gwf_copy_2d')
            evs = _events_struct(dev_ev, present, cm)
            ws = _workspace(st, lib.gwf_workspace_bytes(C.byref(model), cm), (ci % len(side)) if multi else None)
            wsp, wsn = C.c_void_p(ws.data_ptr()), ws.numel()
            groups = [(0, cm)] if multi else _round_groups(cm, round_events)
            if len(groups) > 1:
                # one prologue for the chunk, then the groups
                fo = K.gwf_fisher_out(None, None, None, None, status.data_ptr() + 4 * clo)
                K.check(lib.gwf_fisher_range(C.byref(model), darr, ndet, parr, npsd, C.byref(evs), cm, 0, cm, 1, C.byref(opts), C.byref(fo), wsp, wsn, sp),
                        'gwf_fisher_range')
                launch_count += 1
            for (glo, gm) in groups:
                lo = clo + glo
                packed = torch.empty((npass, gm, npack), dtype=f64, device=st.device)
                direct = npass == 1
                s2 = None if direct else torch.empty((npass, gm), dtype=f64, device=st.device)
                sd = None if (direct or not want_snr_derivs) else torch.empty((npass, gm, nP), dtype=f64, device=st.device)
                s2p = snr2.data_ptr() + 8 * lo if direct else s2.data_ptr()
                sdp = (sder.data_ptr() + 8 * nP * lo if direct else sd.data_ptr()) if want_snr_derivs else None
                fo = outputs(packed, s2p, sdp, lo)
                if len(groups) > 1:
                    K.check(lib.gwf_fisher_range(C.byref(model), darr, ndet, parr, npsd, C.byref(evs), cm, glo, gm, 2, C.byref(opts), C.byref(fo), wsp, wsn, sp),
                            'gwf_fisher_range')
                    launch_count += npass
                else:
                    K.check(lib.gwf_fisher_ex(C.byref(model), darr, ndet, parr, npsd, C.byref(evs), cm, C.byref(opts), C.byref(fo), wsp, wsn, sp), 'gwf_fisher')
                    launch_count += 1 + npass
                unpack(packed, lo, gm, sp)
                if not direct:
                    snr2[:, lo:lo + gm] = s2
                    if want_snr_derivs:
                        sder[:, lo:lo + gm] = sd
                if to_host:
                    if multi or len(groups) == 1:
                        to_host_copy(lo, gm, sp)                     # same stream: the chunk pipeline overlaps it with the other stream's kernels
                    else:
                        done = torch.cuda.Event()
                        done.record(stream)
                        copy_stream.wait_event(done)
                        to_host_copy(lo, gm, C.c_void_p(copy_stream.cuda_stream))
                keep.append((packed, s2, sd))
            keep.append((dev_ev,))
    for s_ in used:
        cur.wait_stream(s_)
    for t in ((full, snr2, sder, status) if multi else (full,)):     # allocated on the current stream, touched on the side streams
        if t is not None:
            for s_ in used:
                t.record_stream(s_)
    st.last_fisher_device = full if STASH_DEVICE else None
    h2d = host_ev.numel() * 8
    if keep_on_device:
        st.last_status = None
        return full, ((snr2, sder) if want_snr_derivs else snr2), (h2d, 0)
    out_f, out_s, out_st, out_d = pinned_outputs()
    out_s.copy_(snr2, non_blocking=True)
    out_st.copy_(status, non_blocking=True)
    if want_snr_derivs:
        out_d.copy_(sder, non_blocking=True)
    cur.synchronize()
    st.last_status = out_st.numpy()
    d2h = (out_f.numel() + out_s.numel()) * 8 + out_st.numel() * 4
    if want_snr_derivs:
        return out_f.numpy(), (out_s.numpy(), out_d.numpy()), (h2d, d2h + out_d.numel() * 8)
    return out_f.numpy(), out_s.numpy(), (h2d, d2h)


def strain_derivs(model, dets, psd_handles, ev, n, res, flags):
    """Run gwf_strain_derivs: complex128 array (n_arms, nP, n, res)."""
    global launch_count
    st = state()
    torch = st.torch
    lib = st.lib
    nP = lib.gwf_num_params(C.byref(model))
    darr, parr = _call_arrays(dets, psd_handles)
    narms = lib.gwf_num_arms(darr, len(dets))
    stream = torch.cuda.current_stream(st.device)
    sp = C.c_void_p(stream.cuda_stream)
    out = np.empty((narms, nP, n, int(res)), dtype=np.complex128)
    opts = K.gwf_opts(int(res), int(flags) | KERNEL_FLAGS, 1, 0)
    # chunks of events sized to ~256 MB of output per launch group
    chunk = max(1, min(n, int(2 ** 28 // (narms * nP * int(res) * 16)) or 1))
    for lo in range(0, n, chunk):
        m = min(chunk, n - lo)
        sub = ev if (lo == 0 and m == n) else {k: np.asarray(v)[lo:lo + m] for k, v in ev.items() if k in K.EVENT_KEYS}
        dev_ev, host_ev, evs, _ = _upload(st, sub, m, K.EVENT_KEYS)
        ws = _workspace(st, lib.gwf_workspace_bytes(C.byref(model), m))
        d = torch.empty((narms, nP, m, int(res), 2), dtype=torch.float64, device=st.device)
        K.check(lib.gwf_strain_derivs(C.byref(model), darr, len(dets), parr, len(psd_handles), C.byref(evs), m, C.byref(opts),
                                      C.c_void_p(d.data_ptr()), C.c_void_p(ws.data_ptr()), ws.numel(), sp), 'gwf_strain_derivs')
        launch_count += 1 + narms
        out[:, :, lo:lo + m, :] = torch.view_as_complex(d).cpu().numpy()
    return out


def overlap(model1, model2, det, psd_handle_, ev1, ev2, fcut, fmin, n, res):
    """GWSignal.WFOverlap on one detector: gwf_strain for each waveform on the common grid, then gwf_overlap per arm.

    Returns (overlap, snr2_1, snr2_2), each (n_arms, n): the per-arm integrals of signal.py:1876-1925.
    """
    global launch_count
    st = state()
    torch = st.torch
    lib = st.lib
    darr, parr = _call_arrays([det], [psd_handle_])
    narms = lib.gwf_num_arms(darr, 1)
    stream = torch.cuda.current_stream(st.device)
    sp = C.c_void_p(stream.cuda_stream)
    opts = K.gwf_opts(int(res), KERNEL_FLAGS, 1, 0)
    out = np.empty((3, narms, n))
    chunk = max(1, min(n, int(2 ** 27 // (narms * int(res) * 16)) or 1))       # ~128 MB of strain per waveform and launch group
    for lo in range(0, n, chunk):
        m = min(chunk, n - lo)
        hs, keep = [], []
        for model, ev in ((model1, ev1), (model2, ev2)):
            sub = {k: np.asarray(v)[lo:lo + m] for k, v in ev.items() if k in K.EVENT_KEYS}
            sub['_fcut'] = np.asarray(fcut)[lo:lo + m]
            dev_ev, host_ev, evs, _ = _upload(st, sub, m, K.EVENT_KEYS)
            ws = torch.empty(int(lib.gwf_workspace_bytes(C.byref(model), m)), dtype=torch.uint8, device=st.device)
            h = torch.empty((narms, m, int(res), 2), dtype=torch.float64, device=st.device)
            K.check(lib.gwf_strain(C.byref(model), darr, 1, parr, 1, C.byref(evs), m, C.byref(opts), C.c_void_p(h.data_ptr()),
                                   C.c_void_p(ws.data_ptr()), ws.numel(), sp), 'gwf_strain')
            launch_count += 1 + narms
            hs.append(h)
            keep.append((dev_ev, host_ev, ws, evs))
        fc = torch.from_numpy(np.array(np.asarray(fcut, dtype=np.float64)[lo:lo + m])).to(st.device)
        res3 = torch.empty((3, narms, m), dtype=torch.float64, device=st.device)
        for a in range(narms):
            K.check(lib.gwf_overlap(C.c_void_p(hs[0][a].data_ptr()), C.c_void_p(hs[1][a].data_ptr()), C.c_void_p(fc.data_ptr()), m, int(res), float(fmin),
                                    C.c_void_p(psd_handle_), C.c_void_p(res3[0, a].data_ptr()), C.c_void_p(res3[1, a].data_ptr()),
                                    C.c_void_p(res3[2, a].data_ptr()), sp), 'gwf_overlap')
            launch_count += 1
        out[:, :, lo:lo + m] = res3.cpu().numpy()
    return out[0], out[1], out[2]


def snr(model, dets, psd_handles, ev, n, res, flags=0, keep_on_device=False):
    """Run gwf_snr: per-arm integrals 4*int (Ap^2+Ac^2)/Sn df, shape (n_arms, n)."""
    global launch_count
    st = state()
    torch = st.torch
    lib = st.lib
    darr, parr = _call_arrays(dets, psd_handles)
    narms = lib.gwf_num_arms(darr, len(dets))
    stream = torch.cuda.current_stream(st.device)
    sp = C.c_void_p(stream.cuda_stream)
    out = torch.empty((narms, n), dtype=torch.float64, device=st.device)
    opts = K.gwf_opts(int(res), int(flags) | KERNEL_FLAGS, 0, 0)
    h2d = 0
    keep = []
    for lo in range(0, n, CHUNK_EVENTS):
        m = min(CHUNK_EVENTS, n - lo)
        sub = ev if (lo == 0 and m == n) else {k: np.asarray(v)[lo:lo + m] for k, v in ev.items() if k in K.EVENT_KEYS}
        dev_ev, host_ev, evs, nb = _upload(st, sub, m, K.EVENT_KEYS)
        h2d += nb
        ws = _workspace(st, lib.gwf_workspace_bytes(C.byref(model), m))
        part = out if m == n else torch.empty((narms, m), dtype=torch.float64, device=st.device)
        K.check(lib.gwf_snr(C.byref(model), darr, len(dets), parr, len(psd_handles), C.byref(evs), m, C.byref(opts),
                            C.c_void_p(part.data_ptr()), C.c_void_p(ws.data_ptr()), ws.numel(), sp), 'gwf_snr')
        launch_count += 2
        if m != n:
            out[:, lo:lo + m] = part
        keep.append((dev_ev, host_ev, part))
    if keep_on_device:
        return out, (h2d, 0)
    host = torch.empty(out.shape, dtype=torch.float64, pin_memory=True)
    host.copy_(out, non_blocking=True)
    stream.synchronize()
    return host.numpy(), (h2d, host.numel() * 8)
