"""Multi-GPU sharding of a catalog: one process per GPU (torchrun), contiguous event shards, ONE all-gather at the end.

Events are independent (no cross-event term anywhere in gwfast/signal.py or network.py), so the only exchange step of the
path is the final gather of the results -- the same partition as the reference's batch pool
(run/calculate_forecasts_from_catalog.py:900-1007, 1016-1024), with NCCL over NVLink instead of result files.  With the
``nccl`` backend the gather runs on device tensors; with ``gloo`` (CPU tests) on host tensors.
"""
import numpy as np


def shard_bounds(n, world, rank):
    """contiguous slice [lo, hi) of rank `rank`: first (n % world) ranks get one extra event"""
    base, extra = divmod(int(n), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_events(evParams, world, rank):
    n = len(np.atleast_1d(evParams['Mc']))
    lo, hi = shard_bounds(n, world, rank)
    return {k: np.ascontiguousarray(np.asarray(v)[lo:hi]) for k, v in evParams.items()}, lo, hi


def all_gather_event_axis(local, n_total, dist=None, device=None):
    """All-gather arrays whose LAST axis is this rank's shard of the event axis into the full (..., n_total) array.

    `local`: numpy array (host) or torch tensor (device).  Shards may be uneven: every rank pads to the largest shard,
    one all_gather_into_tensor moves the padded blocks, and the padding is dropped when the result is assembled.
    """
    import torch
    if dist is None:
        import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    was_numpy = isinstance(local, np.ndarray)
    t = torch.from_numpy(np.ascontiguousarray(local)) if was_numpy else local
    if device is not None:
        t = t.to(device)
    lead = tuple(t.shape[:-1])
    nmax = max(shard_bounds(n_total, world, r)[1] - shard_bounds(n_total, world, r)[0] for r in range(world))
    pad = torch.zeros(lead + (nmax,), dtype=t.dtype, device=t.device)
    pad[..., :t.shape[-1]] = t
    out = torch.empty((world,) + lead + (nmax,), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out.view(-1), pad.contiguous().view(-1))
    parts = []
    for r in range(world):
        lo, hi = shard_bounds(n_total, world, r)
        parts.append(out[r][..., :hi - lo])
    full = torch.cat(parts, dim=-1)
    return full.cpu().numpy() if was_numpy or device is not None else full


class DistributedDetNet(object):
    """A DetNet whose SNR / FisherMatr shard the catalog over the ranks of the default process group.

    Every rank passes the FULL events dict (as every worker of the reference's pool loads the catalog) and receives the
    FULL result; the work done by a rank is its contiguous shard.
    """

    def __init__(self, net, dist=None, device=None):
        self.net = net
        self.device = device
        if dist is None:
            import torch.distributed as dist
        self.dist = dist

    def _device(self):
        if self.device is not None:
            return self.device
        if self.dist.get_backend() == 'nccl':
            import torch
            return torch.device('cuda', torch.cuda.current_device())
        return None

    def SNR(self, evParams, res=1000):
        n = len(np.atleast_1d(evParams['Mc']))
        local, lo, hi = shard_events(evParams, self.dist.get_world_size(), self.dist.get_rank())
        snr = self.net.SNR(local, res=res) if hi > lo else np.zeros(0)
        return all_gather_event_axis(np.asarray(snr, dtype=np.float64), n, self.dist, self._device())

    def FisherMatr(self, evParams, gather='host', **kwargs):
        """``gather='host'`` (default): the full (nP, nP, N) numpy array on every rank.
        ``gather='device'`` (NCCL): returns ``(F_local, F_all)`` -- this rank's shard as a numpy array, and the full matrix as a
        device tensor gathered over NVLink straight from the engine's device-resident result (no host round trip); hosts only read
        their own shard, so the step does not move world x the result through every rank's PCIe link."""
        n = len(np.atleast_1d(evParams['Mc']))
        local, lo, hi = shard_events(evParams, self.dist.get_world_size(), self.dist.get_rank())
        nP = next(iter(self.net.signals.values())).wf_model.nParams
        if gather == 'device' and self._device() is not None:
            import torch
            out, buf = fisher_with_device_gather(self.net, local, n, self.dist, **kwargs)
            world = self.dist.get_world_size()
            sizes = [shard_bounds(n, world, r)[1] - shard_bounds(n, world, r)[0] for r in range(world)]
            full = buf[0] if world == 1 else torch.cat([buf[r][..., :sizes[r]] for r in range(world)], dim=-1)
            return out, full
        F = self.net.FisherMatr(local, **kwargs) if hi > lo else np.zeros((nP, nP, 0))
        full = all_gather_event_axis(np.asarray(F, dtype=np.float64), n, self.dist, self._device())
        return (F, full) if gather == 'device' else full


_gather_buffers = {}


def fisher_with_device_gather(net, local_events, n_total, dist=None, peer=None, **kwargs):
    """``net.FisherMatr(local_events, **kwargs)`` on this rank's shard plus the gather of the device-resident result.

    Returns ``(result, gathered)``: ``result`` is whatever ``FisherMatr`` returns for the shard (host numpy; a tuple with
    ``return_SNR=True``), ``gathered`` the Fisher matrices of ALL ranks as a device tensor that never crossed PCIe:

    * ``peer`` = a :class:`PeerGather`: packed rows ``(world, n_max, nP(nP+1)/2)``, stored into every rank's buffer by the engine's
      own unpack kernels over NVLink (``gwf_unpack_gather``) while the shard is still being computed group by group;
    * otherwise one ``all_gather_into_tensor`` (NCCL) of the engine's device-resident ``(nP, nP, m)`` result into a preallocated
      ``(world, nP, nP, m_max)`` buffer (uneven shards: rank r's events are ``gathered[r, :, :, :m_r]``).
    """
    import torch
    from . import _engine
    if dist is None:
        import torch.distributed as dist
    if peer is not None:
        _engine.PEER = peer
        try:
            out = net.FisherMatr(local_events, **kwargs)
        finally:
            _engine.PEER = None
        return out, peer.finish()
    _engine.STASH_DEVICE = True
    try:
        out = net.FisherMatr(local_events, **kwargs)
        dev = _engine.state().last_fisher_device
    finally:
        _engine.STASH_DEVICE = False
        _engine.state().last_fisher_device = None
    F = np.asarray(out[0] if isinstance(out, tuple) else out, dtype=np.float64)
    if dev is not None and dev.shape[0] == 1 and tuple(dev.shape[1:]) == F.shape:
        local_dev = dev[0]
    else:       # per-arm / re-indexed results (return_all, duty factors, NewtInspiral): upload the host result
        local_dev = torch.from_numpy(np.ascontiguousarray(F)).to(_engine.state().device)
    world, rank = dist.get_world_size(), dist.get_rank()
    m_max = max(shard_bounds(n_total, world, r)[1] - shard_bounds(n_total, world, r)[0] for r in range(world))
    lead = tuple(local_dev.shape[:-1])
    key = (world, lead, m_max, local_dev.device)
    buf = _gather_buffers.get(key)
    if buf is None:
        _gather_buffers.clear()
        buf = _gather_buffers[key] = torch.zeros((world,) + lead + (m_max,), dtype=torch.float64, device=local_dev.device)
    if local_dev.shape[-1] == m_max:
        src = local_dev.contiguous()
    else:
        src = torch.zeros(lead + (m_max,), dtype=torch.float64, device=local_dev.device)
        src[..., :local_dev.shape[-1]] = local_dev
    dist.all_gather_into_tensor(buf.view(-1), src.view(-1))
    return out, buf


class PeerGather(object):
    """Write-based all-gather of the packed Fisher matrices over NVLink peer memory (``gwf_unpack_gather``).

    Every rank owns a ``gathered`` buffer ``(world, n_max, npack)`` in its HBM and maps the buffers of the other ranks into its
    address space through CUDA IPC (one exchange of handles at construction, over the process group).  A rank's unpack kernel then
    stores its packed rows into its own slot of every peer's buffer while it transposes them for the caller: the gather costs no
    extra kernel, no NCCL launch and no staging copy, and the remote stores ride behind the transposition.  ``finish()`` is the
    rendezvous (a barrier on the process group) after which ``gathered`` holds every rank's rows.

    Only for the one-process-per-GPU layout of a single NVSwitch box (``world <= 8``, NCCL backend); ``available()`` says whether
    the peers could be mapped -- callers fall back to ``dist.all_gather_into_tensor`` otherwise.
    """

    def __init__(self, n_max, npack, dist=None):
        import ctypes as C
        import torch
        from . import _capi as K
        if dist is None:
            import torch.distributed as dist
        self.dist, self.torch = dist, torch
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.n_max, self.npack = int(n_max), int(npack)
        if self.world > 8:
            raise ValueError('PeerGather is for the GPUs of one box (world <= 8)')
        dev = torch.device('cuda', torch.cuda.current_device())
        self._lib = lib = K.load()
        self._base = self._opened = None
        self._slot_cache = {}
        nbytes = self.world * self.n_max * self.npack * 8
        handle = C.create_string_buffer(64)
        base = C.c_void_p()
        ok_local = lib.gwf_peer_alloc(nbytes, C.byref(base), handle) == 0
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle.raw) if ok_local else None)
        ptrs = None
        if ok_local and all(h is not None for h in handles):
            self._base = base.value
            ptrs, opened = [], []
            for r, h in enumerate(handles):
                if r == self.rank:
                    ptrs.append(self._base)
                    continue
                p = C.c_void_p()
                if lib.gwf_peer_open(h, C.byref(p)) != 0:
                    ptrs = None
                    break
                ptrs.append(p.value)
                opened.append(p.value)
            self._opened = opened
        ok = torch.tensor([1 if ptrs is not None else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        self._ptrs = ptrs if int(ok.item()) == 1 else None
        self.gathered = None
        if self._ptrs is not None:
            # this rank's buffer as a tensor (zero-copy view of the cudaMalloc'ed block through __cuda_array_interface__)
            class _View(object):
                pass
            v = _View()
            v.__cuda_array_interface__ = dict(shape=(self.world, self.n_max, self.npack), typestr='<f8', data=(self._base, False), version=2, strides=None)
            self.gathered = torch.as_tensor(v, device=dev)
            self._view = v

    def close(self):
        """unmap the peers' buffers and free this rank's (collective: every rank must be done reading)"""
        if self._base is None:
            return
        self.torch.cuda.synchronize()
        self.dist.barrier()
        for p in (self._opened or []):
            self._lib.gwf_peer_close(p)
        self.gathered = None
        self.dist.barrier()
        self._lib.gwf_peer_free(self._base)
        self._base = self._ptrs = self._opened = None

    def available(self):
        return self._ptrs is not None

    def slots(self, lo=0):
        """ctypes array of this rank's slot in every rank's buffer, advanced to event `lo` of the shard"""
        import ctypes as C
        key = int(lo)
        arr = self._slot_cache.get(key)
        if arr is None:
            slot_bytes = self.n_max * self.npack * 8
            arr = self._slot_cache[key] = (C.c_void_p * self.world)(*[p + self.rank * slot_bytes + key * self.npack * 8 for p in self._ptrs])
        return arr

    def unpack_and_scatter(self, packed, n, nP, full, ld, stream):
        """gwf_unpack_gather on `stream`: `packed` (n, npack) device tensor of this rank; `full` (nP, nP, ld) device tensor or None"""
        import ctypes as C
        from . import _capi as K
        lib = K.load()
        K.check(lib.gwf_unpack_gather(C.c_void_p(packed.data_ptr()), int(n), int(nP), C.c_void_p(full.data_ptr()) if full is not None else None,
                                      int(ld), self.slots(0), self.world, C.c_void_p(stream.cuda_stream)), 'gwf_unpack_gather')

    def finish(self):
        """rendezvous: after it returns (stream-ordered on the current stream) every rank's rows are in ``gathered``"""
        self.torch.cuda.current_stream().synchronize()
        self.dist.barrier()
        return self.gathered
